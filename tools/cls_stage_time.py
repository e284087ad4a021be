"""Device time of the text-line orientation stage inside OAROCR::predict.

Runs the pipeline on N synthetic 960x960 pages with and without the classifier attached and prints one JSON line with
the stage's share (oar_ocr_result.ms_cls, CUDA events on the library's stream).  Usage (GPU box):
    python tools/cls_stage_time.py [pages=8] [reps=3]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from oar_ocr_b200 import models, synth
    from oar_ocr_b200.ocr import OAROCRBuilder
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pages = [synth.page(i, 960) for i in range(n)]
    out = {}
    for name, with_cls in (("plain", False), ("with_cls", True)):
        b = (OAROCRBuilder(models.get_blob("det"), models.get_blob("rec"))
             .character_dict_content("\n".join(models.synthetic_dict())).image_batch_size(n).region_batch_size(256))
        if with_cls:
            b = b.with_text_line_orientation_classification(models.get_blob("cls"))
        ocr = b.build()
        best = None
        for _ in range(reps):
            res = ocr.predict(pages)
            t = dict(ocr.last_timing)
            if best is None or t["ms_total"] < best["ms_total"]:
                best = t
        regions = sum(len(r.text_regions) for r in res)
        n180 = sum(1 for r in res for t in r.text_regions if t.orientation_angle == 180.0)
        out[name] = dict(ms_total=round(best["ms_total"], 3), ms_cls=round(best["ms_cls"], 3),
                         ms_rec=round(best["ms_rec"], 3), regions=regions, rotated=n180)
    out["pages"] = n
    print(json.dumps(out))


if __name__ == "__main__":
    main()
