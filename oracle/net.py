"""CPU fp32 executor for OARG layer-list blobs -- TEST INFRASTRUCTURE ONLY.

Stands in for ONNX Runtime's CPU execution provider, which the reference uses
for all network arithmetic (oar-ocr-core/src/core/inference/
ort_infer_execution.rs:178,281; `ort` =2.0.0-rc.13 -> libonnxruntime, absent
here).  Each op is evaluated with the ONNX operator definition of the same
name (Conv, ConvTranspose, GlobalAveragePool, HardSigmoid, HardSwish,
LayerNormalization, MatMul, Softmax, Resize-nearest, AveragePool) in fp32 on
torch-CPU; fp32 implementations differ from ORT/MLAS only in summation order.
PARITY UNPINNED (no model-loading test exists in the reference).

The blob parser here is deliberately independent of oar_ocr_b200.models.
"""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.nn.functional as F

OP_CONV, OP_DWCONV, OP_SE, OP_ADD, OP_UPADD, OP_UPSAMPLE, OP_DECONV2, OP_AVGPOOL, OP_LAYERNORM, OP_ATTN, \
    OP_CTC_HEAD = range(1, 12)
OP_PAD, OP_MAXPOOL = 12, 13  # HGNetV2 stem (models.py): zero padding, max pool
OP_TOKENS = 14  # a map's pixels as token rows of a sequence (decoder memory of the layout detector)


def parse(blob: bytes):
    assert blob[:4] == b"OARG"
    version, kind, n_ops, n_tensors, n_w = struct.unpack_from("<4IQ", blob, 4)
    assert version == 1
    off = 4 + 16 + 8
    ops = []
    for _ in range(n_ops):
        rec = struct.unpack_from("<4i12i4f4q4q", blob, off)
        off += 144
        ops.append(dict(type=rec[0], in0=rec[1], in1=rec[2], out=rec[3], p=rec[4:16], f=rec[16:20],
                        w_off=rec[20:24], w_len=rec[24:28]))
    weights = np.frombuffer(blob, np.float32, n_w, off)
    return kind, n_tensors, ops, weights


def _act(x, a):
    if a == 0:
        return x
    if a == 1:
        return F.relu(x)
    if a == 2:
        return F.hardswish(x)
    if a == 3:
        return F.silu(x)
    if a == 4:
        return torch.sigmoid(x)
    if a == 5:
        return F.hardsigmoid(x)
    if a == 6:  # exact GELU (AIFI feed-forward of the layout detector)
        return F.gelu(x)
    raise ValueError(a)


class OracleNet:
    def __init__(self, blob: bytes):
        self.kind, self.n_tensors, self.ops, self.weights = parse(blob)

    def _w(self, op, i, shape):
        o, n = op["w_off"][i], op["w_len"][i]
        return torch.from_numpy(self.weights[o:o + n].copy()).reshape(shape)

    @torch.no_grad()
    def forward(self, x: np.ndarray, capture=None) -> np.ndarray:
        """x: f32 [B,3,H,W] (the tensor the reference hands to ORT as input "x").
        Returns det: [B,1,H,W] probabilities; rec: [B,T,V] softmax probabilities."""
        t = {0: torch.from_numpy(np.ascontiguousarray(x, np.float32))}
        for op in self.ops:
            ty, p, f = op["type"], op["p"], op["f"]
            a = t[op["in0"]]
            if ty == OP_CONV:
                kh, kw, sh, sw, ph, pw, cin, cout, act = p[:9]
                w = self._w(op, 0, (cout, kh, kw, cin)).permute(0, 3, 1, 2).contiguous()
                b = self._w(op, 1, (cout,))
                y = F.conv2d(a, w, b, (sh, sw), (ph, pw))
                y = _act(y, act)
                if f[0] != 1.0 or f[1] != 0.0:
                    y = y * f[0] + f[1]
                self._store(t, op, y, p[10], p[11])
            elif ty == OP_DWCONV:
                kh, kw, sh, sw, ph, pw, c, act = p[:8]
                w = self._w(op, 0, (kh, kw, c)).permute(2, 0, 1).reshape(c, 1, kh, kw).contiguous()
                b = self._w(op, 1, (c,))
                y = F.conv2d(a, w, b, (sh, sw), (ph, pw), groups=c)
                y = _act(y, act)
                if f[0] != 1.0 or f[1] != 0.0:
                    y = y * f[0] + f[1]
                t[op["out"]] = y
            elif ty == OP_SE:
                c, cm, residual = p[:3]
                m = a.mean(dim=(2, 3))
                h = F.relu(F.linear(m, self._w(op, 0, (cm, c)), self._w(op, 1, (cm,))))
                s = F.linear(h, self._w(op, 2, (c, cm)), self._w(op, 3, (c,)))
                s = torch.clamp(s * f[0] + f[1], 0.0, 1.0)[:, :, None, None]
                t[op["out"]] = a + a * s if residual else a * s
            elif ty == OP_ADD:
                t[op["out"]] = a + t[op["in1"]]
            elif ty == OP_UPADD:
                b = t[op["in1"]]
                t[op["out"]] = a + F.interpolate(b, scale_factor=p[0], mode="nearest")
            elif ty == OP_UPSAMPLE:
                y = a if p[0] == 1 else F.interpolate(a, scale_factor=p[0], mode="nearest")
                self._store(t, op, y, p[10], p[11])
            elif ty == OP_DECONV2:
                cin, cout, act = p[:3]
                w = self._w(op, 0, (2, 2, cout, cin)).permute(3, 2, 0, 1).contiguous()  # [cin][cout][kh][kw]
                b = self._w(op, 1, (cout,))
                y = _act(F.conv_transpose2d(a, w, b, stride=2), act)
                t[op["out"]] = y
            elif ty == OP_AVGPOOL:
                if p[0] == 0 and p[1] == 0:  # global average pool (classifier trunk)
                    t[op["out"]] = a.mean(dim=(2, 3), keepdim=True)
                else:
                    t[op["out"]] = F.avg_pool2d(a, (p[0], p[1]), (p[2], p[3]))
            elif ty == OP_PAD:
                t[op["out"]] = F.pad(a, (p[1], p[3], p[0], p[2]))  # (left, right, top, bottom)
            elif ty == OP_MAXPOOL:
                t[op["out"]] = F.max_pool2d(a, (p[0], p[1]), (p[2], p[3]))
            elif ty == OP_TOKENS:
                B, C_, H, W = a.shape
                o = op["out"]
                if o not in t:
                    t[o] = torch.zeros((B, C_, 1, p[1]), dtype=a.dtype)
                t[o][:, :, 0, p[0]:p[0] + H * W] = a.reshape(B, C_, H * W)
            elif ty == OP_LAYERNORM:
                c = p[0]
                y = F.layer_norm(a.permute(0, 2, 3, 1), (c,), self._w(op, 0, (c,)), self._w(op, 1, (c,)), f[0])
                t[op["out"]] = y.permute(0, 3, 1, 2).contiguous()
            elif ty == OP_ATTN:
                c, heads = p[:2]
                B, _, H, W = a.shape
                x2 = a.permute(0, 2, 3, 1).reshape(B, H * W, c)
                if p[2] == 1:  # positions on the q / k inputs only (encoder.rs:34-79, 179-216)
                    from .rtdetr import sine_position_embedding
                    wq, bq = self._w(op, 0, (3 * c, c)), self._w(op, 1, (3 * c,))
                    xp = x2 + sine_position_embedding(H, W, c)
                    qkv = torch.cat([F.linear(xp, wq[:2 * c], bq[:2 * c]), F.linear(x2, wq[2 * c:], bq[2 * c:])], -1)
                else:
                    qkv = F.linear(x2, self._w(op, 0, (3 * c, c)), self._w(op, 1, (3 * c,)))
                qkv = qkv.reshape(B, H * W, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
                q, k, v = qkv[0] * f[0], qkv[1], qkv[2]
                att = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
                y = (att @ v).permute(0, 2, 1, 3).reshape(B, H * W, c)
                y = F.linear(y, self._w(op, 2, (c, c)), self._w(op, 3, (c,)))
                t[op["out"]] = y.reshape(B, H, W, c).permute(0, 3, 1, 2).contiguous()
            elif ty == OP_CTC_HEAD:
                c, v = p[:2]
                B, _, H, W = a.shape
                x2 = a.permute(0, 2, 3, 1).reshape(B, H * W, c)
                logits = F.linear(x2, self._w(op, 0, (v, c)), self._w(op, 1, (v,)))
                t[op["out"]] = torch.softmax(logits, dim=-1)
            else:
                raise ValueError(ty)
            if capture is not None:
                capture[op["out"]] = t[op["out"]]
        return t[self.ops[-1]["out"]].numpy()

    @staticmethod
    def _store(t, op, y, c_off, c_total):
        if c_total == 0:
            t[op["out"]] = y
            return
        o = op["out"]
        if o not in t:
            t[o] = torch.zeros((y.shape[0], c_total, y.shape[2], y.shape[3]), dtype=y.dtype)
        t[o][:, c_off:c_off + y.shape[1]] = y
