"""Run the 32-page bench configuration several times and report which regions' scores / labels differ between runs
(race hunting: a protocol slip shows as a few regions with large differences, a summation-order effect as many regions
with 1-ulp differences).  Usage: python tools/det_diff.py [runs]   (GPU box only)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    from oar_ocr_b200 import ffi, models, synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    ctx = ffi.Context(0)
    det, rec = ffi.Model(ctx, models.get_blob("det")), ffi.Model(ctx, models.get_blob("rec"))
    ocr = OAROCR(ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 32, 256)
    pages = [synth.page(300 + i, 960) for i in range(32)]
    ref = None
    for run in range(runs):
        res = ocr.predict(pages)
        flat = [(pi, ri, r.confidence, r.label_indices.tobytes(), r.bounding_box.points.tobytes(),
                 float(r.bounding_box.points[:, 0].max() - r.bounding_box.points[:, 0].min()) /
                 max(1.0, float(r.bounding_box.points[:, 1].max() - r.bounding_box.points[:, 1].min())))
                for pi, p in enumerate(res) for ri, r in enumerate(p.text_regions)]
        if ref is None:
            ref = flat
            print("regions", len(flat))
            continue
        nd = 0
        for a, b in zip(ref, flat):
            if a[2] != b[2] or a[3] != b[3] or a[4] != b[4]:
                nd += 1
                if nd <= 12:
                    print(f"  run {run}: page {a[0]} region {a[1]} ratio {a[5]:.2f} score {a[2]:.7f} vs {b[2]:.7f} "
                          f"labels_equal {a[3] == b[3]} box_equal {a[4] == b[4]}")
        print(f"run {run}: {nd} of {len(flat)} regions differ")


if __name__ == "__main__":
    main()
