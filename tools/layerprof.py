"""Per-launch kernel times of one bench step (32 synthetic 960x960 pages), grouped by (kernel, algorithmic bytes).
Usage: python tools/layerprof.py [--engine E] [--out file.json]   (GPU box only)"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", type=int, default=None)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench
    from oar_ocr_b200 import models
    from oar_ocr_b200.ocr import OAROCRBuilder
    ocr = (OAROCRBuilder(models.get_blob("det"), models.get_blob("rec"))
           .character_dict_content("\n".join(models.synthetic_dict())).image_batch_size(32).region_batch_size(256).build())
    if args.engine is not None:
        ocr.det.set_engine(args.engine)
        ocr.rec.set_engine(args.engine)
    ctx = ocr.ctx
    B = args.batch
    pages = bench.make_pages(0, B)
    hs = np.full(B, 960, np.int32)
    ws = np.full(B, 960, np.int32)
    pb = 960 * 960 * 3
    d_base = ctx.device_alloc(B * pb)
    for i, p in enumerate(pages):
        ctx.memcpy_h2d(d_base + i * pb, p)
    ptrs = (C.c_void_p * B)(*[d_base + i * pb for i in range(B)])
    for _ in range(2):
        ctx.l2_flush()
        ocr.predict_raw(ptrs, hs, ws, True)
    ctx.profile(True)
    ctx.l2_flush()
    ocr.predict_raw(ptrs, hs, ws, True)
    recs = ctx.profile_read()
    ctx.profile(False)
    agg = {}
    order = []
    for r in recs:
        k = (r["name"], r["bytes"], r["flops"])
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += r["ms"]
    tot = sum(v[1] for v in agg.values())
    rows = []
    for k in order:
        n, ms = agg[k]
        gbs = k[1] * n / (ms * 1e-3) / 1e9 if ms > 0 else 0
        rows.append(dict(name=k[0], n=n, ms=round(ms, 4), mb=round(k[1] / 1e6, 2), gflop=round(k[2] / 1e9, 3),
                         gbs=round(gbs, 1)))
        print(f"{k[0]:22s} n={n:3d} ms={ms:8.3f} MB/launch={k[1]/1e6:9.2f} GB/s={gbs:8.1f}")
    print("total kernel ms", round(tot, 3), "stage", ocr.last_timing)
    if args.out:
        json.dump(rows, open(args.out, "w"))


if __name__ == "__main__":
    main()
