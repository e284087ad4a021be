"""Generates tests/golden/hotpath_v1.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference cannot be built or imported in this environment (no cargo/rustc, no onnxruntime, no .onnx
files), so these vectors come from oracle/ -- which is itself pinned against the reference's unit-test vectors
by tests/test_oracle_kat.py.  The GPU parity tests compare the CUDA path with these committed vectors and, at
larger sizes, with the oracle run live on the same seeded inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oar_ocr_b200 import models, synth  # noqa: E402
from oracle import cpu, pipeline  # noqa: E402
from oracle.net import OracleNet  # noqa: E402


def smooth_pred(seed, h, w, n):
    """probability-map-like field: soft rotated rectangles, quantised to 1/1024 so it stores compactly"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    p = np.full((h, w), 0.04, np.float32)
    for _ in range(n):
        cx, cy = rng.uniform(30, w - 30), rng.uniform(20, h - 20)
        bw, bh = rng.uniform(20, 70), rng.uniform(6, 14)
        a = np.deg2rad(rng.uniform(-8, 8))
        u = (xx - cx) * np.cos(a) + (yy - cy) * np.sin(a)
        v = -(xx - cx) * np.sin(a) + (yy - cy) * np.cos(a)
        d = np.maximum(np.abs(u) - bw / 2, np.abs(v) - bh / 2)
        p = np.maximum(p, 1.0 / (1.0 + np.exp(d * 3.0)) * rng.uniform(0.8, 0.99))
    return (np.round(p * 1024) / 1024).astype(np.float32)


def main():
    out = {}
    rng = np.random.default_rng(7)
    # row 2: normalize (DB constants)
    imgs = rng.integers(0, 256, (2, 12, 16, 3), dtype=np.uint8)
    out["norm_in"] = imgs
    out["norm_out"] = np.stack([cpu.det_normalize(i) for i in imgs])
    # rows 4-9: DB post-process
    pred = np.stack([smooth_pred(11, 160, 192, 9), smooth_pred(12, 160, 192, 14)])
    pred[1, 40:43, 100:103] = 0.9  # tiny blob: min_side < 3 path
    pred[1, 80, 20:60] = 0.95      # 1-pixel-high line: degenerate hull
    out["db_pred"] = pred
    for i in range(2):
        b, s = cpu.db_postprocess(pred[i], 192, 160)
        out[f"db_boxes{i}"], out[f"db_scores{i}"] = b, s
    b, s = cpu.db_postprocess(pred[0], 384, 240, unclip_ratio=1.5)  # rescale to a different source size
    out["db_boxes_scaled"], out["db_scores_scaled"] = b, s
    # row 11: crops
    page = synth.page(3, 320)
    quads = np.array([
        [[40, 30], [200, 30], [200, 70], [40, 70]],            # exact axis aligned -> copy
        [[42.5, 95.2], [260.1, 101.7], [259.0, 140.3], [41.4, 133.8]],  # homography + bicubic
        [[100, 150], [130, 150], [130, 300], [100, 300]],      # tall -> rotate270 (aligned)
        [[210.3, 160.1], [240.9, 162.0], [236.2, 290.5], [205.6, 288.6]],  # tall, warped, rotate270
        [[-20, -10], [90, -12], [92, 25], [-18, 27]],          # clipped by the image border
        [[400, 400], [500, 400], [500, 450], [400, 450]],      # outside -> Err
        [[10, 10], [20, 20], [30, 30], [40, 40]],              # collinear -> singular -> Err
    ], np.float32)
    out["crop_page"] = page
    out["crop_quads"] = quads
    for i, q in enumerate(quads):
        c = cpu.rotate_crop(page, q)
        out[f"crop{i}"] = c if c is not None else np.zeros((0, 0, 3), np.uint8)
    # row 13: CRNN preprocess
    crops = [synth.crop(0, 48, 320)[:, :200], synth.crop(1, 31, 333), synth.crop(2, 60, 90)]
    for i, c in enumerate(crops):
        out[f"crnn_in{i}"] = c
    out["crnn_out"] = cpu.crnn_preprocess(crops)
    # rows 15-16: CTC
    logits = rng.random((3, 9, 37), dtype=np.float32)
    logits[0, 2, 5] = logits[0, 2, 30] = 2.0  # tie -> last index wins
    logits[1, :, 0] = 3.0                     # all blank
    out["ctc_pred"] = logits
    idx, prob = cpu.ctc_argmax(logits)
    labels, scores, cols, _ = cpu.ctc_decode(idx, prob, 36)  # class 36 is out of dictionary
    out["ctc_idx"], out["ctc_prob"], out["ctc_scores"] = idx, prob, scores
    for i in range(3):
        out[f"ctc_labels{i}"], out[f"ctc_cols{i}"] = labels[i], cols[i]
    # rows 3 / 14: the networks on tiny inputs (fp32 torch-CPU stands in for ORT)
    det = OracleNet(models.get_blob("det"))
    rec = OracleNet(models.get_blob("rec"))
    small = synth.page(5, 96)[:64]
    out["det_in"] = small
    out["det_pred"] = det.forward(cpu.det_normalize(small)[None])[0, 0]
    rc = [synth.crop(3, 48, 96), synth.crop(4, 48, 64)]
    out["rec_in0"], out["rec_in1"] = rc
    r = pipeline.rec_forward(rec, rc, 18385, return_probs=True)
    out["rec_idx"], out["rec_prob"] = r["idx"], r["prob"]
    out["rec_top_probs"] = np.take_along_axis(r["probs"], r["idx"][..., None].astype(np.int64), -1)[..., 0]
    # the whole path on one small page
    pg = synth.page(9, 480)
    out["pipe_page"] = pg
    res = pipeline.predict(det, rec, [pg], 18385, image_batch_size=8, region_batch_size=4)[0]
    out["pipe_boxes"] = np.stack([r["box"] for r in res]) if res else np.zeros((0, 4, 2), np.float32)
    out["pipe_scores"] = np.array([r["score"] for r in res], np.float32)
    out["pipe_label_off"] = np.cumsum([0] + [len(r["labels"]) for r in res]).astype(np.int32)
    out["pipe_labels"] = np.concatenate([r["labels"] for r in res]).astype(np.int32) if res else np.zeros(0, np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.startswith(("db_b", "pipe"))})


if __name__ == "__main__":
    main()
