"""One layout-detection step (8 synthetic 1024x1024 pages -> 640x640, BASELINE.json configs[4] per GPU) for ncu captures.
Usage: python tools/ncu_layout_step.py [--batch 8]   (GPU box only)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    import bench
    from oar_ocr_b200 import ffi, models
    ctx = ffi.Context(0)
    w = models.layout_weights(42)
    shapes = [(bench.LAYOUT_IN[0] // s, bench.LAYOUT_IN[1] // s) for s in (8, 16, 32)]
    enc = ffi.Model(ctx, models.build_layout_encoder(w, seed=42, shapes_hw=shapes))
    head = ffi.Model(ctx, models.build_layout_head(w))
    pages = bench.layout_pages(0, args.batch)
    rows = ffi.layout_rows(enc, head, pages, bench.LAYOUT_IN)
    print("rows", rows.shape, float(rows[0, 0, 1]))


if __name__ == "__main__":
    main()
