"""One bench-configuration step (32 synthetic 960x960 pages, image/region batch 32/256, device-resident pages) for
ncu captures.  Under `ncu --kernel-id ::regex:.:N --set full` this yields one full capture of the N-th launch of every
distinct kernel function, at the sizes the bench line is quoted on.
Usage: python tools/ncu_step.py [--steps 1] [--batch 32] [--rec512]   (GPU box only)"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--rec512", action="store_true", help="configs[2]: one oar_rec_run of 512 synthetic 48x320 crops")
    args = ap.parse_args()
    import bench
    from oar_ocr_b200 import ffi, models, synth
    from oar_ocr_b200.ocr import OAROCRBuilder
    if args.rec512:
        ctx = ffi.Context(0)
        rec = ffi.Model(ctx, models.get_blob("rec"))
        crops = [synth.crop(j, 48, 320) for j in range(512)]
        for _ in range(args.steps):
            r = rec.rec_run(crops, 18385)
        print("rec512 T", r["T"])
        return
    ocr = (OAROCRBuilder(models.get_blob("det"), models.get_blob("rec"))
           .character_dict_content("\n".join(models.synthetic_dict())).image_batch_size(32).region_batch_size(256).build())
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
    for _ in range(args.steps):
        ctx.l2_flush()
        b = ocr.predict_raw(ptrs, hs, ws, True)
    print("regions", int(b.region_off[B]))


if __name__ == "__main__":
    main()
