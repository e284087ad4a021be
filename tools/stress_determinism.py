"""Run-to-run determinism stress for the recognizer: 64-crop batches repeated under host-side jitter (random sleeps, a
background torch CPU load, or arena re-allocation) and compared bitwise with a baseline.  This is what exposed the
two-epilogue-group CTC race; expected output: 0 mismatches.
Usage: python tools/stress_determinism.py [sleep|torch|big] [iterations]   (GPU box only)"""
import sys, numpy as np, os, time, random, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oar_ocr_b200 import ffi, models, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "sleep"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = ffi.Context(0)
rec = ffi.Model(ctx, models.get_blob("rec"))
crops = [synth.crop(j, 48, 320) for j in range(512)]
base = [rec.rec_run(crops[s:s + 64], 18385)["scores"] for s in range(0, 512, 64)]
base2 = [rec.rec_run(crops[s:s + 64], 18385)["scores"] for s in range(0, 512, 64)]
print("baseline stable", all(np.array_equal(a, b) for a, b in zip(base, base2)), flush=True)
stop = False
def burn():
    import torch
    torch.set_num_threads(8)
    a = torch.randn(1, 64, 256, 256)
    w = torch.randn(64, 64, 3, 3)
    while not stop:
        torch.nn.functional.conv2d(a, w, padding=1)
if mode == "torch":
    th = threading.Thread(target=burn); th.start()
random.seed(1)
bad = 0
for i in range(n_iter):
    if mode in ("sleep", "torch"):
        time.sleep(random.random() * 0.03)
    if mode == "big" and i % 8 == 0:
        rec.rec_run(crops, 18385)   # forces arena growth + coalesce on the next call
    k = i % 8
    sc = rec.rec_run(crops[64 * k:64 * k + 64], 18385)["scores"]
    if not np.array_equal(sc, base[k]):
        bad += 1
        d = np.abs(sc - base[k])
        print("  iter", i, "part", k, "max diff", d.max(), "idx", np.nonzero(d)[0][:6], flush=True)
stop = True
print("mode", mode, "mismatches", bad, "of", n_iter, flush=True)
