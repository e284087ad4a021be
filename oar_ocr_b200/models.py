"""Layer-list ("OARG") model graphs for the two networks on the hot path.

The reference treats both networks as opaque ONNX files executed by ONNX
Runtime (oar-ocr-core/src/core/inference/ort_infer_execution.rs:121-306); the
files themselves (pp-ocrv5_mobile_det.onnx / pp-ocrv5_mobile_rec.onnx,
oar-ocr-core/src/core/download/registry.rs:75-76) are not available offline.
This module therefore builds the PP-OCRv5-mobile *architectures* (PP-LCNetV3
backbone + RSEFPN + DBHead for detection; PP-LCNetV3 + SVTR neck + CTC head for
recognition, re-parameterised inference form) as a flat list of ops with
deterministic synthetic weights, and serialises them into one little-endian
blob that both the CUDA engine (csrc/engine.cu, via oar_model_load_blob) and
the CPU oracle (oracle/net.py) execute.  One blob = one source of truth for
layer shapes and weights.

Blob layout (all little endian):
  magic "OARG" | u32 version=1 | u32 kind (0 det, 1 rec, 2 cls) | u32 n_ops |
  u32 n_tensors | u64 n_weight_floats |
  n_ops x OpRec{ i32 type, i32 in0, i32 in1, i32 out, i32 p[12], f32 f[4],
                 i64 w_off[4], i64 w_len[4] }   (144 bytes)
  f32 weights[n_weight_floats]
Tensor 0 is the network input (NHWC f32, 3 channels, BGR-normalised).
Activations are NHWC.  `out` may name a tensor that several ops fill by channel
slice (concat without a copy): p[10] = channel offset, p[11] = total channels
(0 = dense).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC = b"OARG"
VERSION = 1
KIND_DET, KIND_REC, KIND_CLS = 0, 1, 2
KIND_FEAT = 3  # feature extractor (HGNetV2 backbone): spec + oracle only, the CUDA engine does not load it yet

# op types
OP_CONV, OP_DWCONV, OP_SE, OP_ADD, OP_UPADD, OP_UPSAMPLE, OP_DECONV2, OP_AVGPOOL, OP_LAYERNORM, OP_ATTN, \
    OP_CTC_HEAD = range(1, 12)
# Ops the HGNetV2 stem needs.  Executed by the oracle (oracle/net.py) and exported / imported by
# onnx_io; csrc/engine.cu rejects them ("unknown op type") until their kernels exist -- no silent fallback.
OP_PAD, OP_MAXPOOL = 12, 13
# OP_TOKENS: copies a [B,H,W,C] map as H*W token rows into rows [p[0], p[0] + H*W) of a [B,1,p[1],C] sequence (the
# decoder memory of the layout detector is the concatenation of three flattened maps).  Spec + oracle only so far.
OP_TOKENS = 14
# activations
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_SWISH, ACT_SIGMOID, ACT_HSIGMOID = range(6)
ACT_GELU = 6  # exact (erf) GELU: the AIFI feed-forward of the layout detector; oracle / spec only so far


@dataclass
class Op:
    type: int
    in0: int
    in1: int
    out: int
    p: list
    f: list = field(default_factory=lambda: [0.0] * 4)
    w: list = field(default_factory=list)  # up to 4 float32 arrays


class GraphBuilder:
    def __init__(self, kind: int, seed: int):
        self.kind = kind
        self.rng = np.random.default_rng(seed)
        self.ops: list[Op] = []
        self.n_tensors = 1
        self.channels = {0: 3}
        self.base_gain = 1.0

    def new_tensor(self, c: int) -> int:
        t = self.n_tensors
        self.n_tensors += 1
        self.channels[t] = c
        return t

    def he(self, shape, fan_in, gain=1.0):
        # base_gain < 1 keeps hardswish/swish stacks from growing layer over layer
        return (self.rng.standard_normal(shape) * (gain * self.base_gain * np.sqrt(2.0 / fan_in))).astype(np.float32)

    def small(self, shape, s=0.05):
        return (self.rng.standard_normal(shape) * s).astype(np.float32)

    # -- ops -------------------------------------------------------------
    def conv(self, x, cout, k=(1, 1), s=(1, 1), pad=None, act=ACT_NONE, bias=True, post=(1.0, 0.0), w=None, b=None,
             out=None, c_off=0, c_total=0, gain=1.0):
        cin = self.channels[x]
        kh, kw = k
        ph, pw = pad if pad is not None else (kh // 2, kw // 2)
        if w is None:
            w = self.he((cout, kh, kw, cin), kh * kw * cin, gain)
        if b is None:
            b = self.small((cout,)) if bias else np.zeros((cout,), np.float32)
        if out is None:
            out = self.new_tensor(cout)
        self.ops.append(Op(OP_CONV, x, -1, out, [kh, kw, s[0], s[1], ph, pw, cin, cout, act, 1, c_off, c_total],
                           [post[0], post[1], 0.0, 0.0], [w.astype(np.float32), b.astype(np.float32)]))
        return out

    def dwconv(self, x, k, s=(1, 1), act=ACT_NONE, post=(1.0, 0.0), w=None, b=None):
        c = self.channels[x]
        if w is None:
            w = self.he((k, k, c), k * k)
        if b is None:
            b = self.small((c,))
        out = self.new_tensor(c)
        self.ops.append(Op(OP_DWCONV, x, -1, out, [k, k, s[0], s[1], k // 2, k // 2, c, act, 1, 0, 0, 0],
                           [post[0], post[1], 0.0, 0.0], [w, b]))
        return out

    def se(self, x, cmid, residual=False, slope=1.0 / 6.0, offset=0.5, keep0=False):
        c = self.channels[x]
        w1 = self.he((cmid, c), c)
        b1 = self.small((cmid,))
        w2 = self.he((c, cmid), cmid)
        b2 = self.small((c,))
        out = self.new_tensor(c)
        self.ops.append(Op(OP_SE, x, -1, out, [c, cmid, 1 if residual else 0] + [0] * 9, [slope, offset, 0.0, 0.0],
                           [w1, b1, w2, b2]))
        return out

    def add(self, a, b):
        out = self.new_tensor(self.channels[a])
        self.ops.append(Op(OP_ADD, a, b, out, [0] * 12))
        return out

    def upadd(self, a, b, scale=2):
        """out = a + nearest_upsample(b, scale)"""
        out = self.new_tensor(self.channels[a])
        self.ops.append(Op(OP_UPADD, a, b, out, [scale] + [0] * 11))
        return out

    def upsample_into(self, x, scale, out, c_off, c_total):
        self.ops.append(Op(OP_UPSAMPLE, x, -1, out, [scale] + [0] * 9 + [c_off, c_total]))
        return out

    def deconv2(self, x, cout, act=ACT_NONE, w=None, b=None):
        cin = self.channels[x]
        if w is None:
            w = self.he((2, 2, cout, cin), cin)  # [dy][dx][cout][cin]
        if b is None:
            b = self.small((cout,))
        out = self.new_tensor(cout)
        self.ops.append(Op(OP_DECONV2, x, -1, out, [cin, cout, act] + [0] * 9, [1.0, 0.0, 0.0, 0.0], [w, b]))
        return out

    def avgpool(self, x, k, s):
        """k = (0, 0): global average pool"""
        out = self.new_tensor(self.channels[x])
        self.ops.append(Op(OP_AVGPOOL, x, -1, out, [k[0], k[1], s[0], s[1]] + [0] * 8))
        return out

    def pad(self, x, top, left, bottom, right):
        """zero padding of the spatial dims (ONNX Pad, constant 0)"""
        out = self.new_tensor(self.channels[x])
        self.ops.append(Op(OP_PAD, x, -1, out, [top, left, bottom, right] + [0] * 8))
        return out

    def maxpool(self, x, k, s):
        """max pool, no padding, floor mode (ONNX MaxPool)"""
        out = self.new_tensor(self.channels[x])
        self.ops.append(Op(OP_MAXPOOL, x, -1, out, [k[0], k[1], s[0], s[1]] + [0] * 8))
        return out

    def layernorm(self, x, eps, g=None, b=None):
        c = self.channels[x]
        g = (1.0 + self.small((c,), 0.02)).astype(np.float32) if g is None else g
        b = self.small((c,), 0.02) if b is None else b
        out = self.new_tensor(c)
        self.ops.append(Op(OP_LAYERNORM, x, -1, out, [c] + [0] * 11, [eps, 0.0, 0.0, 0.0], [g, b]))
        return out

    def attn(self, x, heads, wqkv=None, bqkv=None, wp=None, bp=None, positions=False):
        """positions=True (p[2] = 1): the 2-D sin/cos position embedding of the map (temperature 10000, layout
        [sin y, cos y, sin x, cos x], oar-ocr-vl/src/models/pp_doclayout/encoder.rs:179-216) is added to the inputs of
        the q and k projections, not to v -- the AIFI layer of the layout detector.  Spec + oracle only so far."""
        c = self.channels[x]
        wqkv = self.he((3 * c, c), c, 0.7) if wqkv is None else wqkv
        bqkv = self.small((3 * c,)) if bqkv is None else bqkv
        wp = self.he((c, c), c, 0.7) if wp is None else wp
        bp = self.small((c,)) if bp is None else bp
        out = self.new_tensor(c)
        self.ops.append(Op(OP_ATTN, x, -1, out, [c, heads, 1 if positions else 0] + [0] * 9,
                           [float((c // heads) ** -0.5), 0, 0, 0], [wqkv, bqkv, wp, bp]))
        return out

    def tokens(self, x, out, row_off, rows_total):
        self.ops.append(Op(OP_TOKENS, x, -1, out, [row_off, rows_total] + [0] * 10))
        return out

    def ctc_head(self, x, vocab, w, b):
        c = self.channels[x]
        out = self.new_tensor(vocab)
        self.ops.append(Op(OP_CTC_HEAD, x, -1, out, [c, vocab] + [0] * 10, [0.0] * 4, [w, b]))
        return out

    # -- serialisation ----------------------------------------------------
    def serialize(self) -> bytes:
        recs = []
        weights = []
        off = 0
        for op in self.ops:
            w_off = [0] * 4
            w_len = [0] * 4
            for i, w in enumerate(op.w):
                w = np.ascontiguousarray(w, np.float32).ravel()
                w_off[i] = off
                w_len[i] = w.size
                weights.append(w)
                off += w.size
                pad = (-off) % 4  # keep every array 16-byte aligned
                if pad:
                    weights.append(np.zeros(pad, np.float32))
                    off += pad
            p = list(op.p) + [0] * (12 - len(op.p))
            f = list(op.f) + [0.0] * (4 - len(op.f))
            recs.append(struct.pack("<4i12i4f4q4q", op.type, op.in0, op.in1, op.out, *p, *f, *w_off, *w_len))
        wcat = np.concatenate(weights) if weights else np.zeros(0, np.float32)
        head = MAGIC + struct.pack("<4IQ", VERSION, self.kind, len(self.ops), self.n_tensors, wcat.size)
        return head + b"".join(recs) + wcat.tobytes()


def make_divisible(v, divisor=16):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# PP-LCNetV3 stage tables: (k, in, out, stride, use_se)
_NET_DET = {
    2: [(3, 16, 32, (1, 1), False)],
    3: [(3, 32, 64, (2, 2), False), (3, 64, 64, (1, 1), False)],
    4: [(3, 64, 128, (2, 2), False), (3, 128, 128, (1, 1), False)],
    5: [(3, 128, 256, (2, 2), False)] + [(5, 256, 256, (1, 1), False)] * 4,
    6: [(5, 256, 512, (2, 2), True), (5, 512, 512, (1, 1), True), (5, 512, 512, (1, 1), False),
        (5, 512, 512, (1, 1), False)],
}
_NET_REC = {
    2: [(3, 16, 32, (1, 1), False)],
    3: [(3, 32, 64, (1, 1), False), (3, 64, 64, (1, 1), False)],
    4: [(3, 64, 128, (2, 1), False), (3, 128, 128, (1, 1), False)],
    5: [(3, 128, 256, (1, 2), False)] + [(5, 256, 256, (1, 1), False)] * 4,
    6: [(5, 256, 512, (2, 1), True), (5, 512, 512, (1, 1), True), (5, 512, 512, (2, 1), False),
        (5, 512, 512, (1, 1), False)],
}


def _plant_row0(w, n=1):
    """make output channels 0..n-1 copy input channels 0..n-1 through the centre tap only"""
    for j in range(n):
        w[j] = 0.0
        w[j, w.shape[1] // 2, w.shape[2] // 2, j] = 1.0
    return w


def _lcnet_block(g: GraphBuilder, x, k, cout, stride, use_se, plant):
    """plant = number of leading channels carried through unchanged (identity taps)"""
    c = g.channels[x]
    wd = g.he((k, k, c), k * k)
    bd = g.small((c,))
    if plant:
        wd[:, :, :plant] = 0.0
        wd[k // 2, k // 2, :plant] = 1.0
        bd[:plant] = 0.0
    # LearnableRepLayer (deploy): conv -> [hardswish unless stride 2] ; post = the Act's learnable affine
    act = ACT_NONE if stride != (1, 1) else ACT_HSWISH
    x = g.dwconv(x, k, stride, act=act, w=wd, b=bd)
    if use_se:
        x = g.se(x, c // 4)
    wp = g.he((cout, 1, 1, c), c)
    bp = g.small((cout,))
    if plant:
        wp = _plant_row0(wp, plant)
        bp[:plant] = 0.0
    return g.conv(x, cout, (1, 1), act=ACT_HSWISH, w=wp, b=bp)


def build_det(seed: int = 42, scale: float = 0.75, signal_gain: float = 5.0) -> bytes:
    """PP-OCRv5_mobile_det shaped graph: PP-LCNetV3(0.75, det) + RSEFPN(96) + DBHead.

    Synthetic weights with a planted signal: channel 0 carries
    "darkness" from the stem through the stride-4 trunk into the DB head so that
    dark text lines map to p ~ 0.97 and light background to p ~ 0.05; every other
    weight is He-normal so the arithmetic volume is that of the real network.
    """
    g = GraphBuilder(KIND_DET, seed)
    md = lambda c: make_divisible(c * scale)
    # stem: conv3x3 s2 + BN (no activation)
    c1 = md(16)
    w = g.he((c1, 3, 3, 3), 27)
    b = g.small((c1,))
    w[0] = 0.0
    w[0, 1, 1, :] = -signal_gain
    b[0] = 0.0
    x = g.conv(0, c1, (3, 3), (2, 2), act=ACT_NONE, w=w, b=b)
    feats = []
    for stage in (2, 3, 4, 5, 6):
        for (k, _cin, cout, s, use_se) in _NET_DET[stage]:
            x = _lcnet_block(g, x, k, md(cout), s, use_se, plant=1 if stage <= 3 else 0)
        if stage >= 3:
            feats.append(x)
    mv = [16, 24, 56, 480]
    outs = []
    for i, f in enumerate(feats):
        co = int(mv[i] * scale)
        wq = g.he((co, 1, 1, g.channels[f]), g.channels[f])
        bq = g.small((co,))
        if i == 0:
            wq = _plant_row0(wq)
            bq[0] = 0.0
        outs.append(g.conv(f, co, (1, 1), act=ACT_NONE, w=wq, b=bq))
    # RSEFPN(out=96, shortcut=True)
    oc = 96
    ins = []
    for i, f in enumerate(outs):
        wq = g.he((oc, 1, 1, g.channels[f]), g.channels[f], 0.5)
        if i == 0:
            wq = _plant_row0(wq)
        else:
            wq[0] = 0.0  # keep the planted channel clean in the top-down sum
        t = g.conv(f, oc, (1, 1), act=ACT_NONE, bias=False, w=wq, b=np.zeros(oc, np.float32))
        ins.append(g.se(t, oc // 4, residual=True, slope=0.2, offset=0.5))
    in2, in3, in4, in5 = ins
    out4 = g.upadd(in4, in5)
    out3 = g.upadd(in3, out4)
    out2 = g.upadd(in2, out3)
    fuse = g.new_tensor(oc)
    q = oc // 4
    ps = []
    for i, f in enumerate((in5, out4, out3, out2)):
        wq = g.he((q, 3, 3, oc), 9 * oc, 0.5)
        if i == 3:
            wq = _plant_row0(wq)
        t = g.conv(f, q, (3, 3), act=ACT_NONE, bias=False, w=wq, b=np.zeros(q, np.float32))
        ps.append(g.se(t, q // 4, residual=True, slope=0.2, offset=0.5))
    p5, p4, p3, p2 = ps
    g.upsample_into(p5, 8, fuse, 0, oc)
    g.upsample_into(p4, 4, fuse, q, oc)
    g.upsample_into(p3, 2, fuse, 2 * q, oc)
    g.upsample_into(p2, 1, fuse, 3 * q, oc)
    # DBHead binarize branch (BN folded): conv3x3 -> relu -> deconv -> relu -> deconv -> sigmoid
    wq = g.he((q, 3, 3, oc), 9 * oc, 0.5)
    bq = g.small((q,))
    wq[0] = 0.0
    wq[0, 1, 1, 3 * q] = 1.0
    bq[0] = 0.0
    x = g.conv(fuse, q, (3, 3), act=ACT_RELU, w=wq, b=bq)
    wd = g.he((2, 2, q, q), q)
    bd = g.small((q,))
    wd[:, :, 0, :] = 0.0
    wd[:, :, 0, 0] = 1.0
    bd[0] = 0.0
    x = g.deconv2(x, q, act=ACT_RELU, w=wd, b=bd)
    wd = (g.rng.standard_normal((2, 2, 1, q)) * 0.002).astype(np.float32)
    wd[:, :, 0, 0] = 1.0
    x = g.deconv2(x, 1, act=ACT_SIGMOID, w=wd, b=np.array([-3.0], np.float32))
    return g.serialize()


def _svtr_block(g: GraphBuilder, x, heads, mlp_ratio=2.0):
    c = g.channels[x]
    n1 = g.layernorm(x, 1e-5)
    a = g.attn(n1, heads)
    x = g.add(x, a)
    n2 = g.layernorm(x, 1e-5)
    h = g.conv(n2, int(c * mlp_ratio), (1, 1), act=ACT_SWISH, gain=0.7)
    h = g.conv(h, c, (1, 1), act=ACT_NONE, gain=0.7)
    return g.add(x, h)


def build_rec(seed: int = 42, vocab: int = 18385, scale: float = 0.95, logit_gain: float = 6.0,
              blank_bias: float = 9.0) -> bytes:
    """PP-OCRv5_mobile_rec shaped graph: PP-LCNetV3(0.95, rec strides) + avgpool(3,2)
    + EncoderWithSVTR(dims 120, depth 2) + CTC Linear(120 -> vocab) + softmax.
    vocab = len(ppocrv5_dict.txt) + 2 = 18385 (decode.rs:392-423)."""
    g = GraphBuilder(KIND_REC, seed)
    md = lambda c: make_divisible(c * scale)
    # stem: NP planted channels = darkness seen through different 3x3 taps, so the
    # per-timestep feature vector follows the local stroke texture of the crop.
    NP = 8
    c1 = md(16)
    w = g.he((c1, 3, 3, 3), 27)
    b = g.small((c1,))
    taps = [(1, 1), (1, 0), (1, 2), (0, 1), (2, 1), (0, 0), (2, 2), (0, 2)]
    for j, (ty, tx) in enumerate(taps):
        w[j] = 0.0
        w[j, ty, tx, :] = -1.5
        w[j, 1, 1, :] += -0.5 * (j % 3)
        b[j] = 0.5 + 0.25 * j
    x = g.conv(0, c1, (3, 3), (2, 2), act=ACT_NONE, w=w, b=b)
    for stage in (2, 3, 4, 5, 6):
        for (k, _cin, cout, s, use_se) in _NET_REC[stage]:
            x = _lcnet_block(g, x, k, md(cout), s, use_se, plant=NP)
    x = g.avgpool(x, (3, 2), (3, 2))
    cin = g.channels[x]
    h = x
    z = g.conv(x, cin // 8, (1, 3), act=ACT_SWISH)
    z = g.conv(z, 120, (1, 1), act=ACT_SWISH)
    for _ in range(2):
        z = _svtr_block(g, z, 8)
    z = g.layernorm(z, 1e-6)
    cat = g.new_tensor(2 * cin)
    g.upsample_into(h, 1, cat, 0, 2 * cin)
    g.conv(z, cin, (1, 1), act=ACT_SWISH, out=cat, c_off=cin, c_total=2 * cin)
    # conv4 reads the planted backbone channels (first NP of the guide half) strongly
    w4 = g.he((cin // 8, 1, 3, 2 * cin), 3 * 2 * cin, 0.5)
    w4[:, :, :, :NP] = (g.rng.standard_normal((cin // 8, 1, 3, NP)) * 0.35).astype(np.float32)
    z = g.conv(cat, cin // 8, (1, 3), act=ACT_SWISH, w=w4)
    z = g.conv(z, 120, (1, 1), act=ACT_SWISH)
    w = (g.rng.standard_normal((vocab, 120)) * (logit_gain / np.sqrt(120.0))).astype(np.float32)
    b = g.small((vocab,), 0.1)
    b[0] = blank_bias
    g.ctc_head(z, vocab, w, b)
    return g.serialize()


# PP-LCNet (v1) stage table of the text-line orientation classifier: (k, in, out, stride, use_se).  PaddleClas'
# textline_orientation configuration keeps the width after the stem: stride_list [2, [2,1], [2,1], [2,1], [2,1]] on
# 80 x 160 inputs (EXT: recalled from the PULC configuration; the .onnx file is not available offline).
_NET_CLS = {
    2: [(3, 16, 32, (1, 1), False)],
    3: [(3, 32, 64, (2, 1), False), (3, 64, 64, (1, 1), False)],
    4: [(3, 64, 128, (2, 1), False), (3, 128, 128, (1, 1), False)],
    5: [(3, 128, 256, (2, 1), False)] + [(5, 256, 256, (1, 1), False)] * 5,
    6: [(5, 256, 512, (2, 1), True), (5, 512, 512, (1, 1), True)],
}
CLS_INPUT_SHAPE = (80, 160)  # TextLineOrientationAdapter::DEFAULT_INPUT_SHAPE, text_line_orientation_adapter.rs:48


def _lcnet_v1_block(g: GraphBuilder, x, k, cout, stride, use_se, plant):
    """PP-LCNet v1 DepthwiseSeparable: dw conv + BN + hardswish -> [SE] -> 1x1 conv + BN + hardswish.  The first
    `plant` channels pass through untouched by the random weights (identity centre tap, or a box filter where the
    layer is strided; identity rows in the 1x1 conv)."""
    c = g.channels[x]
    wd = g.he((k, k, c), k * k)
    bd = g.small((c,))
    if plant:
        wd[:, :, :plant] = 0.0
        if stride == (1, 1):
            wd[k // 2, k // 2, :plant] = 1.0
        else:  # strided layers average their window, so the planted statistics see every input row
            wd[:, :, :plant] = 1.0 / (k * k)
        bd[:plant] = 0.0
    x = g.dwconv(x, k, stride, act=ACT_HSWISH, w=wd, b=bd)
    if use_se:
        x = g.se(x, c // 4)
        if plant:  # gate of the planted channels saturated at 1
            g.ops[-1].w[2][:plant] = 0.0
            g.ops[-1].w[3][:plant] = 6.0
    wp = g.he((cout, 1, 1, c), c)
    bp = g.small((cout,))
    if plant:
        wp = _plant_row0(wp, plant)
        wp[plant:, :, :, :plant] = 0.0
        bp[:plant] = 0.0
    return g.conv(x, cout, (1, 1), act=ACT_HSWISH, w=wp, b=bp)


def build_cls(seed: int = 42, scale: float = 1.0, num_classes: int = 2, dark_ref: float = -5.55, dark_gain: float = 6.0) -> bytes:
    """PP-LCNet_x1_0_textline_ori shaped graph (oar-ocr-core/src/models/classification/pp_lcnet.rs runs it through
    ORT): stem 3x3 s2 -> 13 depthwise-separable blocks -> global average pool -> 1x1 conv 1280 + hardswish ->
    Linear(1280 -> num_classes) + softmax.  Output [B, 1, num_classes] probabilities; class 1 = "180".

    Synthetic weights with a planted, well-conditioned decision (everything else is He-normal so the arithmetic volume
    is the real network's): stem channel 0 = 10 + darkness of the pixel (always >= 3, where hardswish is the identity),
    carried through the trunk by identity taps / box filters, so the average pool yields D = 10 + mean darkness of the
    crop; the head computes logit_180 - logit_0 = 2 [hs(g (D - 10 - dark_ref)) - hs(-g (D - 10 - dark_ref))]: dark
    crops are called "180", light ones "0".  Not a real orientation cue -- it makes both classes occur on the synthetic
    pages with wide margins, which is what the parity tests need (the rotate180 that follows changes the recognised
    text, so a wrong or missing rotation cannot go unnoticed).
    """
    g = GraphBuilder(KIND_CLS, seed)
    g.base_gain = 0.8
    md = lambda c: make_divisible(c * scale)
    NP = 1
    c1 = md(16)
    w = g.he((c1, 3, 3, 3), 27)
    b = g.small((c1,))
    w[0] = 0.0
    w[0, 1, 1, :] = -1.0   # normalised input: dark < 0
    b[0] = 10.0
    x = g.conv(0, c1, (3, 3), (2, 2), act=ACT_HSWISH, w=w, b=b)
    for stage in (2, 3, 4, 5, 6):
        for (k, _cin, cout, s, use_se) in _NET_CLS[stage]:
            x = _lcnet_v1_block(g, x, k, md(cout), s, use_se, plant=NP)
    x = g.avgpool(x, (0, 0), (0, 0))  # (0, 0) window = global average pool: any input shape
    cin = g.channels[x]
    wl = g.he((1280, 1, 1, cin), cin)
    bl = np.zeros((1280,), np.float32)   # last_conv carries no bias in PaddleClas; the planted rows use one
    wl[:2] = 0.0
    wl[2:, :, :, :NP] = 0.0
    wl[0, 0, 0, 0], bl[0] = dark_gain, -dark_gain * (10.0 + dark_ref)
    wl[1, 0, 0, 0], bl[1] = -dark_gain, dark_gain * (10.0 + dark_ref)
    x = g.conv(x, 1280, (1, 1), act=ACT_HSWISH, w=wl, b=bl)
    wf = (g.rng.standard_normal((num_classes, 1280)) * 0.002).astype(np.float32)
    bf = np.zeros((num_classes,), np.float32)
    wf[:, :2] = 0.0
    if num_classes >= 2:
        wf[1, 0], wf[1, 1] = 1.0, -1.0
        wf[0, 0], wf[0, 1] = -1.0, 1.0
    g.ctc_head(x, num_classes, wf, bf)   # Linear + softmax over the single "timestep"
    return g.serialize()


# HGNetV2-L: the backbone of PP-DocLayout-L / RT-DETR-L and, with other strides, of the
# server-size OCR models (item 4).  Stage table as the reference's in-tree description states it
# (oar-ocr-vl/src/models/pp_doclayout/hgnetv2.rs:13-23).
_HG_STEM = (3, 32, 48)
_HG_STAGES = [  # (in, mid, out, blocks, downsample, light, kernel, layers)
    (48, 48, 128, 1, False, False, 3, 6),
    (128, 96, 512, 1, True, False, 3, 6),
    (512, 192, 1024, 3, True, True, 5, 6),
    (1024, 384, 2048, 1, True, True, 5, 6),
]


def _hg_block(g: GraphBuilder, x, cin, mid, cout, layers, k, light, residual):
    """HGNetV2 BasicLayer (hgnetv2.rs:161-262): `layers` convs in a chain, every output (and the input) concatenated,
    two 1x1 aggregation convs, optional identity residual.  The concat is channel-slice writes into one buffer."""
    total = cin + layers * mid
    cat = g.new_tensor(total)
    g.upsample_into(x, 1, cat, 0, total)
    h = x
    for i in range(layers):
        if light:  # ConvLayerLight: 1x1 conv (no activation) then depthwise k x k + ReLU
            h = g.conv(h, mid, (1, 1), act=ACT_NONE)
            h = g.dwconv(h, k, (1, 1), act=ACT_RELU)
        elif i == layers - 1:  # nothing else reads the last layer: it writes its concat slice directly
            g.conv(h, mid, (k, k), act=ACT_RELU, out=cat, c_off=cin + i * mid, c_total=total)
            continue
        else:
            h = g.conv(h, mid, (k, k), act=ACT_RELU)
        g.upsample_into(h, 1, cat, cin + i * mid, total)
    y = g.conv(cat, cout // 2, (1, 1), act=ACT_RELU)
    y = g.conv(y, cout, (1, 1), act=ACT_RELU)
    return g.add(y, x) if residual else y


def build_hgnetv2_l(seed: int = 42, return_idx=(3,), taps: list | None = None) -> bytes:
    """HGNetV2-L feature extractor (hgnetv2.rs: Embeddings :264-348, Stage :264-330, BasicLayer :161-262), BatchNorm
    folded into the conv biases, He-normal synthetic weights.  The graph's output is the last requested stage
    (strides 4 / 8 / 16 / 32, 128 / 512 / 1024 / 2048 channels); `taps`, when given, receives the tensor id of every
    stage output (a detector neck reads several of them).  Spec + oracle only for now (KIND_FEAT)."""
    g = GraphBuilder(KIND_FEAT, seed)
    _hgnetv2_l_into(g, return_idx, taps)
    return g.serialize()


def _hgnetv2_l_into(g: GraphBuilder, return_idx=(3,), taps: list | None = None):
    g.base_gain = 0.7
    c0, c1, c2 = _HG_STEM
    x = g.conv(0, c1, (3, 3), (2, 2), act=ACT_RELU)                       # stem1
    p = g.pad(x, 0, 0, 1, 1)                                              # pad_right_bottom
    b = g.conv(p, c1 // 2, (2, 2), pad=(0, 0), act=ACT_RELU)              # stem2a
    cat = g.new_tensor(2 * c1)
    g.upsample_into(g.maxpool(p, (2, 2), (1, 1)), 1, cat, 0, 2 * c1)
    g.conv(g.pad(b, 0, 0, 1, 1), c1, (2, 2), pad=(0, 0), act=ACT_RELU, out=cat, c_off=c1, c_total=2 * c1)  # stem2b
    x = g.conv(cat, c1, (3, 3), (2, 2), act=ACT_RELU)                     # stem3
    x = g.conv(x, c2, (1, 1), act=ACT_RELU)                               # stem4
    last = max(return_idx)
    for si, (cin, mid, cout, blocks, down, light, k, layers) in enumerate(_HG_STAGES):
        if down:  # depthwise 3x3 stride 2, no activation
            x = g.dwconv(x, 3, (2, 2), act=ACT_NONE)
        for bi in range(blocks):
            x = _hg_block(g, x, cin if bi == 0 else cout, mid, cout, layers, k, light, residual=bi != 0)
        if taps is not None:
            taps.append(x)
        if si == last:
            break
    return x


# Synthetic weight table of the layout detector's neck and decoder (RT-DETR-L: hybrid encoder + 6-layer deformable
# decoder, d = 256, 8 heads, FFN 1024, 3 levels x 4 points, 300 queries), by name.  The backbone's weights come from
# build_hgnetv2_l with the same seed.  Consumed by build_layout_encoder (the OARG form of backbone + encoder) and by
# the CPU oracle of the whole detector (oracle/rtdetr.py, which documents what each name is).
def layout_weights(seed: int = 42, num_labels: int = 23) -> dict:
    import math
    D, HEADS, FFN, LEVELS, POINTS, DEC_LAYERS = 256, 8, 1024, 3, 4, 6
    rng = np.random.default_rng(seed + 1)
    w = {}

    def lin(name, cin, cout, gain=1.0, bias=0.02):
        w[name + ".w"] = (rng.standard_normal((cout, cin)) * gain / math.sqrt(cin)).astype(np.float32)
        w[name + ".b"] = (rng.standard_normal(cout) * bias).astype(np.float32)

    def conv(name, cin, cout, k, gain=1.0):  # BatchNorm folded: a conv with bias
        w[name + ".w"] = (rng.standard_normal((cout, cin, k, k)) * gain * math.sqrt(2.0 / (cin * k * k))).astype(np.float32)
        w[name + ".b"] = (rng.standard_normal(cout) * 0.02).astype(np.float32)

    def ln(name):
        w[name + ".g"] = (1.0 + rng.standard_normal(D) * 0.02).astype(np.float32)
        w[name + ".b"] = (rng.standard_normal(D) * 0.02).astype(np.float32)

    def csp(name):
        conv(name + ".conv1", 2 * D, D, 1)
        conv(name + ".conv2", 2 * D, D, 1)
        for i in range(3):
            conv(f"{name}.rep{i}.c3", D, D, 3, 0.7)
            conv(f"{name}.rep{i}.c1", D, D, 1, 0.7)

    for l, c in enumerate((512, 1024, 2048)):
        conv(f"input_proj{l}", c, D, 1)
        conv(f"dec_input_proj{l}", D, D, 1)
    for n in ("q", "k", "v", "o"):
        lin(f"aifi.{n}", D, D)
    lin("aifi.fc1", D, FFN)
    lin("aifi.fc2", FFN, D)
    ln("aifi.ln1")
    ln("aifi.ln2")
    for i in range(2):
        conv(f"lateral{i}", D, D, 1)
        csp(f"fpn{i}")
        conv(f"down{i}", D, D, 3)
        csp(f"pan{i}")
    lin("enc_output", D, D)
    ln("enc_output_ln")
    lin("enc_score", D, num_labels, 2.0, 0.5)
    for i, (a, b) in enumerate(((D, D), (D, D), (D, 4))):
        lin(f"enc_bbox{i}", a, b, 1.0 if i < 2 else 0.3)
    lin("query_pos0", 4, 2 * D)
    lin("query_pos1", 2 * D, D)
    for i in range(DEC_LAYERS):
        p = f"dec{i}"
        for n in ("q", "k", "v", "o"):
            lin(f"{p}.sa.{n}", D, D)
        ln(f"{p}.ln1")
        lin(f"{p}.ca.offsets", D, HEADS * LEVELS * POINTS * 2, 0.5, 1.0)
        lin(f"{p}.ca.weights", D, HEADS * LEVELS * POINTS)
        lin(f"{p}.ca.value", D, D)
        lin(f"{p}.ca.out", D, D)
        ln(f"{p}.ln2")
        lin(f"{p}.fc1", D, FFN)
        lin(f"{p}.fc2", FFN, D)
        ln(f"{p}.ln3")
        lin(f"{p}.score", D, num_labels, 2.0, 0.5)
        for j, (a, b) in enumerate(((D, D), (D, D), (D, 4))):
            lin(f"{p}.bbox{j}", a, b, 1.0 if j < 2 else 0.3)
    return w


def build_layout_encoder(weights: dict, seed: int = 42, shapes_hw=None) -> bytes:
    """Backbone + hybrid encoder + decoder-input projection of the layout detector (RT-DETR-L) as ONE OARG graph whose
    output is the decoder memory [B, 1, sum(H_l W_l), 256] (8400 tokens at 640 x 640): HGNetV2-L -> 1x1 projections ->
    AIFI on the stride-32 map -> CCFM (top-down then bottom-up, CSPRep blocks) -> 1x1 projections -> OP_TOKENS.
    `weights` is the detector's weight table by name (the oracle's, oracle/rtdetr.py documents the names; this module
    never imports the oracle); the backbone is regenerated from `seed` exactly as build_hgnetv2_l does.  The RepVGG
    3x3 + 1x1 branches are merged into one 3x3 convolution (their sum before the activation is one convolution).
    `shapes_hw`: the three map sizes (strides 8 / 16 / 32) of the input the graph will see, needed only to lay the
    token rows out.  Spec + oracle only: ACT_GELU, OP_ATTN with positions, OP_TOKENS have no CUDA kernels yet."""
    W = {k: np.asarray(v, np.float32) for k, v in weights.items()}
    g = GraphBuilder(KIND_FEAT, seed)
    taps = []
    _hgnetv2_l_into(g, (3,), taps)
    d = 256

    def cw(name):  # [cout, cin, kh, kw] -> [cout, kh, kw, cin]
        return np.ascontiguousarray(W[name + ".w"].transpose(0, 2, 3, 1)), W[name + ".b"]

    def conv(x, name, k=1, s=1, act=ACT_SWISH, **kw):
        w, b = cw(name)
        return g.conv(x, w.shape[0], (k, k), (s, s), act=act, w=w, b=b, **kw)

    def lin_as_conv(x, name, act=ACT_NONE):
        w = W[name + ".w"]
        return g.conv(x, w.shape[0], (1, 1), act=act, w=np.ascontiguousarray(w[:, None, None, :]), b=W[name + ".b"])

    def csp(x, name):
        y = conv(x, name + ".conv1")
        for i in range(3):
            w3, b3 = cw(f"{name}.rep{i}.c3")
            w1, b1 = cw(f"{name}.rep{i}.c1")
            w3 = w3.copy()
            w3[:, 1, 1, :] += w1[:, 0, 0, :]
            y = g.conv(y, d, (3, 3), act=ACT_SWISH, w=w3, b=b3 + b1)
        return g.add(y, conv(x, name + ".conv2"))

    feats = [conv(taps[l + 1], f"input_proj{l}", act=ACT_NONE) for l in range(3)]
    # AIFI: post-norm transformer layer with positions on q / k
    t = feats[2]
    wqkv = np.concatenate([W["aifi.q.w"], W["aifi.k.w"], W["aifi.v.w"]], 0)
    bqkv = np.concatenate([W["aifi.q.b"], W["aifi.k.b"], W["aifi.v.b"]], 0)
    a = g.attn(t, 8, wqkv=wqkv, bqkv=bqkv, wp=W["aifi.o.w"], bp=W["aifi.o.b"], positions=True)
    t = g.layernorm(g.add(t, a), 1e-5, W["aifi.ln1.g"], W["aifi.ln1.b"])
    f = lin_as_conv(lin_as_conv(t, "aifi.fc1", ACT_GELU), "aifi.fc2")
    feats[2] = g.layernorm(g.add(t, f), 1e-5, W["aifi.ln2.g"], W["aifi.ln2.b"])
    # CCFM
    fpn = [feats[2]]
    for i in range(2):
        top = conv(fpn[-1], f"lateral{i}")
        fpn[-1] = top
        cat = g.new_tensor(2 * d)
        g.upsample_into(top, 2, cat, 0, 2 * d)
        g.upsample_into(feats[1 - i], 1, cat, d, 2 * d)
        fpn.append(csp(cat, f"fpn{i}"))
    fpn.reverse()
    pan = [fpn[0]]
    for i in range(2):
        cat = g.new_tensor(2 * d)
        conv(pan[-1], f"down{i}", 3, 2, out=cat, c_off=0, c_total=2 * d)
        g.upsample_into(fpn[i + 1], 1, cat, d, 2 * d)
        pan.append(csp(cat, f"pan{i}"))
    # decoder memory
    total = sum(h * w for h, w in shapes_hw)
    mem = g.new_tensor(d)
    off = 0
    for l, f in enumerate(pan):
        g.tokens(conv(f, f"dec_input_proj{l}", act=ACT_NONE), mem, off, total)
        off += shapes_hw[l][0] * shapes_hw[l][1]
    return g.serialize()


# Layout detector head (RT-DETR-L decoder) as a table of single layers in OARG form.  It is NOT a graph to execute in
# order: csrc/layout_net.cu runs the decoder (query selection, 6 layers of self-attention + multi-scale deformable
# attention + FFN, box refinement) and calls these layers by POSITION, so that their weights get the same packing /
# tensor-core kernels as every other 1x1 convolution, LayerNorm and attention block.  The order below is the ABI between
# this builder and layout_net.cu (LH_* constants there).
LAYOUT_HEAD_FIXED = 8        # enc_output, enc_output_ln, enc_score, enc_bbox0..2, query_pos0, query_pos1
LAYOUT_HEAD_PER_LAYER = 14   # sa, ln1, ca.offsets, ca.weights, ca.value, ca.out, ln2, fc1, fc2, ln3, score, bbox0..2


def build_layout_head(weights: dict, dec_layers: int = 6) -> bytes:
    W = {k: np.asarray(v, np.float32) for k, v in weights.items()}
    g = GraphBuilder(KIND_FEAT, 0)
    d = 256
    src = {}

    def inp(c):  # one placeholder tensor per input width (the ops are never chained)
        if c not in src:
            src[c] = g.new_tensor(c)
        return src[c]

    def lin(name, act=ACT_NONE):
        w = W[name + ".w"]
        return g.conv(inp(w.shape[1]), w.shape[0], (1, 1), act=act, w=np.ascontiguousarray(w[:, None, None, :]), b=W[name + ".b"])

    def ln(name):
        return g.layernorm(inp(d), 1e-5, W[name + ".g"], W[name + ".b"])

    lin("enc_output"); ln("enc_output_ln"); lin("enc_score")
    lin("enc_bbox0", ACT_RELU); lin("enc_bbox1", ACT_RELU); lin("enc_bbox2")
    lin("query_pos0", ACT_RELU); lin("query_pos1")
    assert len(g.ops) == LAYOUT_HEAD_FIXED
    for i in range(dec_layers):
        p = f"dec{i}"
        wqkv = np.concatenate([W[p + ".sa.q.w"], W[p + ".sa.k.w"], W[p + ".sa.v.w"]], 0)
        bqkv = np.concatenate([W[p + ".sa.q.b"], W[p + ".sa.k.b"], W[p + ".sa.v.b"]], 0)
        g.attn(inp(d), 8, wqkv=wqkv, bqkv=bqkv, wp=W[p + ".sa.o.w"], bp=W[p + ".sa.o.b"])
        ln(p + ".ln1")
        lin(p + ".ca.offsets"); lin(p + ".ca.weights"); lin(p + ".ca.value"); lin(p + ".ca.out")
        ln(p + ".ln2")
        lin(p + ".fc1", ACT_RELU); lin(p + ".fc2")
        ln(p + ".ln3")
        lin(p + ".score")
        lin(p + ".bbox0", ACT_RELU); lin(p + ".bbox1", ACT_RELU); lin(p + ".bbox2")
    assert len(g.ops) == LAYOUT_HEAD_FIXED + LAYOUT_HEAD_PER_LAYER * dec_layers
    return g.serialize()


# Server-size recogniser: PP-OCRv5_server_rec = PPHGNetV2_B4 (the table above; "B4" and "L" name the
# same widths) with text-recognition strides + the SVTR neck + CTC head of the mobile model.  EXT: the stride table is
# recalled from PaddleOCR's rec_pphgnetv2 (stem3 stride 1; stage downsamples (2,1), (1,2), (2,1), (2,1); every stage
# downsamples); no .onnx file is available offline to confirm it.
_HG_REC_STRIDES = [(2, 1), (1, 2), (2, 1), (2, 1)]


def build_rec_server(seed: int = 42, vocab: int = 18385) -> bytes:
    """PP-OCRv5_server_rec shaped graph: HGNetV2-B4 (rec strides: 48 x W -> 3 x W/4 x 2048) + avgpool(3,2) +
    EncoderWithSVTR(dims 120, depth 2) + CTC Linear(120 -> vocab) + softmax; He-normal synthetic weights.
    Spec + oracle only: the CUDA engine has no kernels for OP_PAD / OP_MAXPOOL yet and fails loudly on them."""
    g = GraphBuilder(KIND_REC, seed)
    g.base_gain = 0.7
    c0, c1, c2 = _HG_STEM
    x = g.conv(0, c1, (3, 3), (2, 2), act=ACT_RELU)
    p = g.pad(x, 0, 0, 1, 1)
    b = g.conv(p, c1 // 2, (2, 2), pad=(0, 0), act=ACT_RELU)
    cat = g.new_tensor(2 * c1)
    g.upsample_into(g.maxpool(p, (2, 2), (1, 1)), 1, cat, 0, 2 * c1)
    g.conv(g.pad(b, 0, 0, 1, 1), c1, (2, 2), pad=(0, 0), act=ACT_RELU, out=cat, c_off=c1, c_total=2 * c1)
    x = g.conv(cat, c1, (3, 3), (1, 1), act=ACT_RELU)   # stem3 keeps the resolution for text recognition
    x = g.conv(x, c2, (1, 1), act=ACT_RELU)
    for (cin, mid, cout, blocks, _down, light, k, layers), s in zip(_HG_STAGES, _HG_REC_STRIDES):
        x = g.dwconv(x, 3, s, act=ACT_NONE)
        for bi in range(blocks):
            x = _hg_block(g, x, cin if bi == 0 else cout, mid, cout, layers, k, light, residual=bi != 0)
    x = g.avgpool(x, (3, 2), (3, 2))
    cin = g.channels[x]
    h = x
    z = g.conv(x, cin // 8, (1, 3), act=ACT_SWISH)
    z = g.conv(z, 120, (1, 1), act=ACT_SWISH)
    for _ in range(2):
        z = _svtr_block(g, z, 8)
    z = g.layernorm(z, 1e-6)
    cat = g.new_tensor(2 * cin)
    g.upsample_into(h, 1, cat, 0, 2 * cin)
    g.conv(z, cin, (1, 1), act=ACT_SWISH, out=cat, c_off=cin, c_total=2 * cin)
    z = g.conv(cat, cin // 8, (1, 3), act=ACT_SWISH)
    z = g.conv(z, 120, (1, 1), act=ACT_SWISH)
    w = (g.rng.standard_normal((vocab, 120)) * (4.0 / np.sqrt(120.0))).astype(np.float32)
    g.ctc_head(z, vocab, w, g.small((vocab,), 0.1))
    return g.serialize()


def synthetic_dict(vocab: int = 18385) -> list[str]:
    """Character list for synthetic runs: vocab-2 distinct code points (CJK block
    onward), standing in for ppocrv5_dict.txt (one char per line, ocr.rs:386)."""
    n = vocab - 2
    out = []
    cp = 0x4E00
    while len(out) < n:
        out.append(chr(cp))
        cp += 1
    return out


_cache: dict = {}


def get_blob(kind: str, seed: int = 42, vocab: int = 18385) -> bytes:
    key = (kind, seed, vocab)
    if key not in _cache:
        _cache[key] = build_det(seed) if kind == "det" else (build_cls(seed) if kind == "cls" else build_rec(seed, vocab))
    return _cache[key]
