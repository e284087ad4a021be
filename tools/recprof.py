"""Recognition-only workload for ncu captures: one chunk of 256 synthetic 48x320 crops (BASELINE configs[2] shape).
Usage: python tools/recprof.py [--engine E] [--n 256] [--reps 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", type=int, default=None)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    from oar_ocr_b200 import ffi, models, synth
    ctx = ffi.Context(0)
    rec = ffi.Model(ctx, models.get_blob("rec"))
    if args.engine is not None:
        rec.set_engine(args.engine)
    crops = [synth.crop(j, 48, 320) for j in range(args.n)]
    for _ in range(args.reps):
        r = rec.rec_run(crops, 18385)
    print("T", r["T"], "labels", sum(len(l) for l in r["labels"]))


if __name__ == "__main__":
    main()
