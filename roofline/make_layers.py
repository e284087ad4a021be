"""Algorithmic FLOPs and bytes per layer of the OARG graphs (SURVEY.md §8d asks for this table to be recomputed
from the builder's own layer list).  Bytes are fp32 NHWC input (read once) + output per op; for the fused engine the
depthwise output of a [depthwise -> 1x1] block is dropped (it never reaches HBM).
Usage: python roofline/make_layers.py > roofline/layers.json"""
import json
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oar_ocr_b200 import models  # noqa: E402

NAMES = {v: k for k, v in vars(models).items() if k.startswith("OP_")}


def parse(blob):
    _, _, n_ops, n_t, _ = struct.unpack_from("<4IQ", blob, 4)
    ops = []
    for i in range(n_ops):
        rec = struct.unpack_from("<4i12i4f4q4q", blob, 28 + 144 * i)
        ops.append(dict(type=rec[0], in0=rec[1], in1=rec[2], out=rec[3], p=rec[4:16]))
    return ops, n_t


def layers(blob, B, H, W):
    ops, n_t = parse(blob)
    shape = {0: (H, W, 3)}
    rows = []
    for i, op in enumerate(ops):
        h, w, c = shape[op["in0"]]
        p = op["p"]
        t = op["type"]
        flops = 0
        oh, ow, oc = h, w, c
        if t == models.OP_CONV:
            kh, kw, sh, sw, ph, pw, cin, cout = p[:8]
            oh, ow = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
            oc = p[11] or cout
            flops = 2 * oh * ow * cout * kh * kw * cin
            out_c = cout
        elif t == models.OP_DWCONV:
            k, _, sh, sw = p[:4]
            oh, ow = (h + 2 * (k // 2) - k) // sh + 1, (w + 2 * (k // 2) - k) // sw + 1
            flops = 2 * oh * ow * c * k * k
            out_c = c
        elif t == models.OP_DECONV2:
            oh, ow, oc = 2 * h, 2 * w, p[1]
            flops = 2 * h * w * 4 * p[1] * p[0]
            out_c = oc
        elif t == models.OP_UPSAMPLE:
            oh, ow = h * p[0], w * p[0]
            oc = p[11] or c
            out_c = c
        elif t == models.OP_AVGPOOL:
            if p[0] == 0 and p[1] == 0:  # global pool
                oh, ow = 1, 1
            else:
                oh, ow = (h - p[0]) // p[2] + 1, (w - p[1]) // p[3] + 1
            out_c = c
        elif t == models.OP_PAD:
            oh, ow = h + p[0] + p[2], w + p[1] + p[3]
            out_c = c
        elif t == models.OP_MAXPOOL:
            oh, ow = (h - p[0]) // p[2] + 1, (w - p[1]) // p[3] + 1
            out_c = c
        elif t == models.OP_TOKENS:
            oh, ow = 1, p[1]
            out_c = c
            shape[op["out"]] = (oh, ow, c)
            rows.append(dict(op=i, type=NAMES.get(t, str(t)), in_hwc=[h, w, c], out_hwc=[oh, ow, c], gflop=0.0,
                             mb_in=B * h * w * c * 4 / 1e6, mb_out=B * h * w * c * 4 / 1e6))
            continue
        elif t == models.OP_ATTN:
            T = h * w
            flops = 2 * T * c * 3 * c + 2 * T * c * c + 4 * T * T * c
            out_c = c
        elif t == models.OP_CTC_HEAD:
            oc = p[1]
            flops = 2 * h * w * c * oc
            out_c = 0  # the fused head writes (index, prob) per timestep only
        else:
            out_c = c
        shape[op["out"]] = (oh, ow, oc)
        rows.append(dict(op=i, type=NAMES.get(t, str(t)), in_hwc=[h, w, c], out_hwc=[oh, ow, oc],
                         gflop=B * flops / 1e9, mb_in=B * h * w * c * 4 / 1e6, mb_out=B * oh * ow * out_c * 4 / 1e6))
    # fused engine: drop the depthwise round trip of [DWCONV -> 1x1 CONV] pairs
    fused_saved = 0.0
    for a, b in zip(rows, rows[1:]):
        if a["type"] == "OP_DWCONV" and b["type"] == "OP_CONV" and ops[b["op"]]["p"][0] == 1 and \
                ops[b["op"]]["in0"] == ops[a["op"]]["out"]:
            fused_saved += a["mb_out"] + b["mb_in"]
    return rows, fused_saved


def main():
    out = {}
    # cls = the optional text-line orientation classifier: one chunk of 256 crops at 80 x 160
    # rec_server = PP-OCRv5_server_rec shaped graph, spec + oracle only so far
    # hgnetv2_l = the backbone of PP-DocLayout-L (BASELINE.json configs[4]: batch 64; the reference resizes every page
    # to 640 x 640 before the network) -- spec + oracle only so far
    for kind, (B, H, W) in (("det", (32, 960, 960)), ("rec", (256, 48, 320)), ("cls", (256, 80, 160)),
                            ("hgnetv2_l", (64, 640, 640)), ("rec_server", (256, 48, 320)),
                            ("layout_encoder", (64, 640, 640))):
        if kind == "hgnetv2_l":
            blob = models.build_hgnetv2_l()
        elif kind == "rec_server":
            blob = models.build_rec_server()
        elif kind == "layout_encoder":
            # backbone + hybrid encoder + decoder-input projections of RT-DETR-L
            blob = models.build_layout_encoder(models.layout_weights(),
                                               shapes_hw=[(H // 8, W // 8), (H // 16, W // 16), (H // 32, W // 32)])
        else:
            blob = models.get_blob(kind)
        rows, saved = layers(blob, B, H, W)
        tot_f = sum(r["gflop"] for r in rows)
        tot_b = sum(r["mb_in"] + r["mb_out"] for r in rows)
        out[kind] = dict(batch=B, input_hw=[H, W], gflop_total=round(tot_f, 2), gflop_per_item=round(tot_f / B, 3),
                         mb_total_per_layer_kernels=round(tot_b, 1), mb_total_fused_engine=round(tot_b - saved, 1),
                         hbm_ms_at_6650_gbs_fused=round((tot_b - saved) / 6650.0, 3),
                         tensor_ms_at_1590_tflops_x3=round(3 * tot_f / 1590.0, 3), layers=rows)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
