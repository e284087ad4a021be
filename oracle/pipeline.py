"""CPU restatement of the pipeline composition -- TEST INFRASTRUCTURE ONLY.

Follows OAROCR::predict (src/oarocr/ocr.rs:518-659), TextDetectionAdapter::execute ->
DBModel::forward (oar-ocr-core/src/models/detection/db.rs:281-335), crop_text_regions
(ocr.rs:718-753), recognize_global (ocr.rs:802-897) and CRNNModel::forward_refs
(oar-ocr-core/src/models/recognition/crnn.rs:247-293) on top of oracle/cpu.py (pre/post
arithmetic) and oracle/net.py (the networks, fp32 torch-CPU standing in for ONNX Runtime).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.
"""
from __future__ import annotations

import numpy as np

from . import cpu
from .net import OracleNet

MAX_POOLED_CROPS = 4096  # ocr.rs:603
BASE_REC_RATIO = np.float32(320.0) / np.float32(48.0)  # DEFAULT_REC_IMAGE_SHAPE, ocr.rs:817


def is_cjk(ch):
    """OAROCR::is_cjk, ocr.rs:1065-1084"""
    u = ord(ch)
    for lo, hi in ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F), (0x2B740, 0x2B81F)):
        if lo <= u <= hi:
            return True
    return False


def ctc_word_boxes(box, text, col_indices, seq_len, wh_ratio, max_wh_ratio):
    """OAROCR::ctc_word_boxes, ocr.rs:949-1022, f32 step by step.  box: [4,2] quad; returns [n,4] (x0,y0,x1,y1)."""
    F = np.float32
    if len(col_indices) == 0 or seq_len == 0 or len(text) == 0:
        return np.zeros((0, 4), F)
    eps = F(np.finfo(np.float32).eps)
    effective = F(F(seq_len) * F(F(wh_ratio) / F(max_wh_ratio)))
    if effective <= eps:
        return np.zeros((0, 4), F)
    box = np.asarray(box, F).reshape(-1, 2)
    x_min, x_max, y_min, y_max = box[:, 0].min(), box[:, 0].max(), box[:, 1].min(), box[:, 1].max()
    width = F(x_max - x_min)
    cell_width = F(width / (effective if effective > eps else eps))
    avg_char_width = F(width / F(max(len(text), 1)))
    centers = []
    for idx in col_indices:
        centers.append(F(x_min + F(F(F(int(idx)) + F(0.5)) * cell_width)))
    out = np.zeros((len(col_indices), 4), F)
    for i in range(len(col_indices)):
        ch = text[i] if i < len(text) else "?"
        c = centers[i]
        if is_cjk(ch):
            half = F(avg_char_width / F(2.0))
            a, b = F(c - half), F(c + half)
            a = a if a > x_min else x_min
            b = b if b < x_max else x_max
        else:
            a = x_min if i == 0 else F(F(centers[i - 1] + c) / F(2.0))
            a = a if a > x_min else x_min
            b = x_max if i == len(col_indices) - 1 else F(F(c + centers[i + 1]) / F(2.0))
            b = b if b < x_max else x_max
        out[i] = (a, y_min, b, y_max)
    return out


def det_forward(net: OracleNet, images, thresh=0.3, box_thresh=0.6, unclip_ratio=2.0, max_candidates=1000,
                limit=960, limit_type=0, max_side=4000, return_pred=False):
    """DBModel::forward: resize -> same-shape groups -> normalize -> net -> DBPostProcess.
    Returns per image (boxes [n,4,2], scores [n]) in discovery order."""
    n = len(images)
    resized = []
    for img in images:
        h, w, _ = img.shape
        src = img
        if h + w < 64:  # image_padding, resize_detection.rs:204-220
            pad = np.zeros((max(h, 32), max(w, 32), 3), np.uint8)
            pad[:h, :w] = img
            src = pad
        rh, rw = cpu.det_resize_dims(src.shape[0], src.shape[1], limit, limit_type, max_side)
        if (rh, rw) != src.shape[:2]:
            src = cpu.resize_triangle(src, rw, rh)
        resized.append(src)
    groups = []
    for i, r in enumerate(resized):
        for g in groups:
            if resized[g[0]].shape == r.shape:
                g.append(i)
                break
        else:
            groups.append([i])
    out = [None] * n
    preds = [None] * n
    for g in groups:
        x = np.stack([cpu.det_normalize(resized[i]) for i in g])
        pred = net.forward(x)  # [B,1,H,W]
        for k, i in enumerate(g):
            sh, sw = images[i].shape[:2]
            out[i] = cpu.db_postprocess(pred[k, 0], sw, sh, thresh, box_thresh, unclip_ratio, max_candidates)
            preds[i] = pred[k, 0]
    return (out, preds) if return_pred else out


def rec_forward(net: OracleNet, crops, n_chars, return_probs=False, with_margin=False):
    """CRNNModel::forward_refs on one batch: preprocess -> net -> argmax -> CTC decode."""
    x = cpu.crnn_preprocess(crops)
    probs = net.forward(x)  # [B,T,V]
    idx, prob = cpu.ctc_argmax(probs)
    labels, scores, cols, T = cpu.ctc_decode(idx, prob, n_chars)
    r = dict(labels=labels, scores=scores, cols=cols, T=T, idx=idx, prob=prob)
    # smallest top-1 / top-2 probability gap over the timesteps of each crop: how far the crop's arg-max decisions are
    # from a tie.  A parity check may excuse a label difference only where this is below the float tolerance.
    # (test-only: the timed CPU baseline never asks for it)
    if with_margin:
        import torch
        t2 = torch.topk(torch.from_numpy(np.ascontiguousarray(probs)).reshape(len(crops), -1, probs.shape[-1]), 2, dim=-1).values
        r["margin"] = (t2[..., 0] - t2[..., 1]).min(dim=1).values.numpy()
    if return_probs:
        r["probs"] = probs
    return r


def cls_forward(net: OracleNet, crops, topk=1, input_shape=cpu.CLS_INPUT_SHAPE):
    """TextLineOrientationAdapter::execute -> PPLCNetModel::forward_refs (text_line_orientation_adapter.rs:63-121,
    pp_lcnet.rs:139-196, 255-300): preprocess -> net -> Topk.  Returns per crop (class_ids, scores) of the top-k and
    the full probability rows."""
    x = cpu.cls_preprocess(crops, input_shape)
    if len(x) == 0:
        return [], np.zeros((0, 0), np.float32)
    probs = net.forward(x)
    probs = probs.reshape(probs.shape[0], -1)
    return [cpu.topk(p, topk) for p in probs], probs


def predict(det_net: OracleNet, rec_net: OracleNet, images, n_chars, image_batch_size=8, region_batch_size=64,
            rec_score_thresh=0.0, det_kwargs=None, chars=None, cls_net: OracleNet | None = None, with_margin=False):
    """OAROCR::predict.  Returns per image a list of dicts
    {box [4,2], det_index, labels (int array), score} in detection-index (reading) order.  With `cls_net` the crops of
    each image go through classify_line_orientations (ocr.rs:615, 755-792) before they are pooled: `angle` = 0 / 180
    and crops of class 1 are rotated by 180 degrees."""
    det_kwargs = det_kwargs or {}
    n = len(images)
    all_boxes = [None] * n
    for s in range(0, n, image_batch_size):
        chunk = images[s:s + image_batch_size]
        for k, (boxes, _scores) in enumerate(det_forward(det_net, chunk, **det_kwargs)):
            all_boxes[s + k] = cpu.sort_quad_boxes(boxes)[0] if len(boxes) else boxes
    results = [[None] * len(b) for b in all_boxes]
    pool = []

    def recognize_global(pool):
        order = sorted(range(len(pool)), key=lambda i: pool[i][3])  # stable sort by wh_ratio
        for s in range(0, len(order), region_batch_size):
            chunk = [pool[i] for i in order[s:s + region_batch_size]]
            r = rec_forward(rec_net, [c[2] for c in chunk], n_chars, with_margin=with_margin)
            chunk_max = BASE_REC_RATIO
            for c in chunk:
                chunk_max = max(chunk_max, np.float32(c[3]))  # ocr.rs:828-831
            for k, (img_idx, det_idx, _crop, ratio) in enumerate(chunk):
                score = float(r["scores"][k])
                keep = score >= rec_score_thresh
                labels = r["labels"][k] if keep else r["labels"][k][:0]
                res = dict(box=all_boxes[img_idx][det_idx], det_index=det_idx, labels=labels, score=score,
                           cols=r["cols"][k], T=r["T"], angle=angles.get((img_idx, det_idx)),
                           margin=float(r["margin"][k]) if with_margin else None)
                if chars is not None:  # return_word_box (ocr.rs:860-868); chars = index -> character
                    cols = r["cols"][k] if keep else r["cols"][k][:0]
                    text = "".join(chars[i] for i in labels if 0 < i < len(chars))
                    res["word_boxes"] = ctc_word_boxes(res["box"], text, cols, r["T"], ratio, chunk_max)
                results[img_idx][det_idx] = res

    angles = {}
    for i, img in enumerate(images):
        crops = []
        for k, box in enumerate(all_boxes[i]):
            crop = cpu.rotate_crop(img, box)
            if crop is None:
                continue
            ratio = np.float32(crop.shape[1]) / np.float32(max(crop.shape[0], 1))  # before any rotation, ocr.rs:735
            crops.append([i, k, crop, float(ratio)])
        if cls_net is not None and crops:
            tops, _ = cls_forward(cls_net, [c[2] for c in crops], topk=1)
            for c, (ids, _sc) in zip(crops, tops):
                if len(ids) == 0:
                    continue
                angles[(c[0], c[1])] = float(ids[0]) * 180.0
                if ids[0] == 1:
                    c[2] = cpu.rotate180(c[2])
        for c in crops:
            pool.append(tuple(c))
            if len(pool) >= MAX_POOLED_CROPS:
                recognize_global(pool)
                pool = []
    if pool:
        recognize_global(pool)
    return [[r for r in res if r is not None] for res in results]
