"""Race hunting: run one 64-crop recogniser batch many times with OAR_DBG_DUMP_CTC set and report WHERE the CTC-head
partials (per row, N tile and column half: max and sum of exponentials) differ from the first run.
Usage: OAR_CTC_GROUPS=2 python tools/ctc_dump_diff.py [iterations]   (GPU box only)"""
import os, sys, struct, time, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = "/tmp/ctc_dump.bin"
os.environ["OAR_DBG_DUMP_CTC"] = path
from oar_ocr_b200 import ffi, models, synth
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = ffi.Context(0)
rec = ffi.Model(ctx, models.get_blob("rec"))
crops = [synth.crop(j, 48, 320) for j in range(64)]
def run():
    if os.path.exists(path): os.remove(path)
    rec.rec_run(crops, 18385)
    raw = open(path, "rb").read()
    rows, nt = struct.unpack("qq", raw[:16])
    a = np.frombuffer(raw[16:], np.float32).reshape(2, rows, nt)
    return a.copy()
base = run()
rows, nt = base.shape[1], base.shape[2]
print("rows", rows, "partials per row", nt, flush=True)
random.seed(2)
bad = 0
for i in range(n_iter):
    time.sleep(random.random() * 0.02)
    a = run()
    if not np.array_equal(a, base):
        bad += 1
        for which, name in ((0, "max"), (1, "sum")):
            d = np.argwhere(a[which] != base[which])
            if len(d) == 0: continue
            rws, cols = d[:, 0], d[:, 1]
            tiles = cols // 2
            rt = rws // 128
            # work item = (row tile, N tile); with resident weights CTA = (N tile) + 72 * (row tile parity), ti = row tile // 2
            items = sorted(set(zip(rt.tolist(), tiles.tolist())))
            rel = np.abs(a[which][rws, cols] - base[which][rws, cols]) / np.maximum(np.abs(base[which][rws, cols]), 1e-30)
            print(f"iter {i} {name}: {len(d)} entries, items (row tile, N tile) {items[:6]}{'...' if len(items) > 6 else ''}, "
                  f"rows-in-tile {sorted(set((rws % 128).tolist()))[:12]}, halves {sorted(set((cols % 2).tolist()))}, "
                  f"rel diff {rel.min():.2e}..{rel.max():.2e}", flush=True)
print("mismatching iterations", bad, "of", n_iter)
