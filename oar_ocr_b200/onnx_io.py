"""ONNX <-> OARG conversion without the onnx package (protobuf wire format by hand).

The reference loads `pp-ocrv5_mobile_{det,rec}.onnx` into ONNX Runtime (oar-ocr-core/src/core/inference/
ort_infer_builders.rs:9-70); this library executes OARG layer lists (models.py).  This module closes the gap in both
directions:

  export_onnx(blob) -> bytes   an OARG graph as a standard ONNX model (opset 17, NCHW, input "x"), so a maintainer can
                               run our synthetic networks through the reference's own ONNX Runtime path and compare;
  import_onnx(data) -> bytes   an ONNX model -> OARG blob for `oar_model_load_blob`.  Handles the operator subset of
                               the two PP-OCR networks: Conv (dense / depthwise, BatchNormalization folded), Relu /
                               HardSwish / Sigmoid / HardSigmoid and their decomposed forms (HardSigmoid*x, Sigmoid*x),
                               squeeze-excite (GlobalAveragePool-Conv-Relu-Conv-HardSigmoid-Mul[-Add]), Resize-nearest,
                               Add, Concat, ConvTranspose 2x2/s2, AveragePool, LayerNormalization between NHWC
                               transposes, the single-input multi-head attention block and the MatMul+Softmax CTC head
                               in the form export_onnx writes them; for classifiers (PP-LCNet line orientation) a
                               plain GlobalAveragePool and the Flatten | Reshape | Squeeze -> Gemm | MatMul [+ Add] ->
                               Softmax tail as PaddleClas / torch exports spell it; Dropout / Identity pass through.
  python -m oar_ocr_b200.onnx_io convert model.onnx model.oarg | export model.oarg model.onnx

No real PP-OCR .onnx file is available offline, so import_onnx is verified by round trips of both synthetic networks
(tests/test_onnx_io.py: OARG -> ONNX -> OARG gives the same blob-level graph and identical oracle outputs) and by an
independent evaluation of the exported ONNX with the ONNX operator definitions.  Anything outside the subset raises
OCRError("ModelLoad") naming the node.
"""
from __future__ import annotations

import struct
import sys

import numpy as np

from . import models as M
from .ffi import OCRError

# ---------------------------------------------------------------------------------------------------------------
# protobuf wire format (only what ModelProto needs)
# ---------------------------------------------------------------------------------------------------------------


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _f_varint(field, v):
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field, b):
    if isinstance(b, str):
        b = b.encode()
    return _key(field, 2) + _varint(len(b)) + b


def _f_float(field, v):
    return _key(field, 5) + struct.pack("<f", float(v))


def _read_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7


def parse_message(buf) -> dict:
    """field number -> list of raw values (int for varint/fixed, bytes for length-delimited)"""
    buf = memoryview(buf)
    out, pos, n = {}, 0, len(buf)
    while pos < n:
        k, pos = _read_varint(buf, pos)
        field, wire = k >> 3, k & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise OCRError("ModelLoad", f"unsupported protobuf wire type {wire}")
        out.setdefault(field, []).append(v)
    return out


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _ints(vals) -> list:
    """repeated int64: packed (one bytes blob) or unpacked"""
    out = []
    for v in vals:
        if isinstance(v, int):
            out.append(_signed(v))
        else:
            pos, n = 0, len(v)
            while pos < n:
                x, pos = _read_varint(v, pos)
                out.append(_signed(x))
    return out


def _floats(vals) -> list:
    out = []
    for v in vals:
        b = bytes(v)
        out.extend(struct.unpack("<%df" % (len(b) // 4), b))
    return out


# ---------------------------------------------------------------------------------------------------------------
# ONNX building blocks
# ---------------------------------------------------------------------------------------------------------------
FLOAT, INT64 = 1, 7
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7


def tensor_proto(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    dt = FLOAT if arr.dtype == np.float32 else INT64
    assert arr.dtype in (np.float32, np.int64)
    msg = b"".join(_f_varint(1, d) for d in arr.shape) + _f_varint(2, dt) + _f_bytes(8, name) + _f_bytes(9, arr.tobytes())
    return msg


def attr(name: str, value) -> bytes:
    msg = _f_bytes(1, name)
    if isinstance(value, float):
        msg += _f_float(2, value) + _f_varint(20, A_FLOAT)
    elif isinstance(value, (int, np.integer)):
        msg += _f_varint(3, value) + _f_varint(20, A_INT)
    elif isinstance(value, str):
        msg += _f_bytes(4, value) + _f_varint(20, A_STRING)
    elif isinstance(value, (list, tuple)) and all(isinstance(v, (int, np.integer)) for v in value):
        msg += b"".join(_f_varint(8, v) for v in value) + _f_varint(20, A_INTS)
    elif isinstance(value, (list, tuple)):
        msg += b"".join(_f_float(7, v) for v in value) + _f_varint(20, A_FLOATS)
    else:
        raise TypeError(type(value))
    return msg


def node(op_type: str, inputs, outputs, name: str = "", **attrs) -> bytes:
    msg = b"".join(_f_bytes(1, i) for i in inputs) + b"".join(_f_bytes(2, o) for o in outputs)
    msg += _f_bytes(3, name or outputs[0]) + _f_bytes(4, op_type)
    msg += b"".join(_f_bytes(5, attr(k, v)) for k, v in attrs.items())
    return msg


def value_info(name: str, dims) -> bytes:
    shape = b""
    for d in dims:
        dim = _f_bytes(2, d) if isinstance(d, str) else _f_varint(1, d)
        shape += _f_bytes(1, dim)
    ttype = _f_varint(1, FLOAT) + _f_bytes(2, shape)
    return _f_bytes(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def model_proto(nodes, initializers, inputs, outputs, name="oar_ocr_b200") -> bytes:
    graph = b"".join(_f_bytes(1, n) for n in nodes) + _f_bytes(2, name)
    graph += b"".join(_f_bytes(5, t) for t in initializers)
    graph += b"".join(_f_bytes(11, i) for i in inputs) + b"".join(_f_bytes(12, o) for o in outputs)
    opset = _f_bytes(1, "") + _f_varint(2, 17)
    return _f_varint(1, 8) + _f_bytes(2, "oar-ocr-b200") + _f_bytes(7, graph) + _f_bytes(8, opset)


# ---------------------------------------------------------------------------------------------------------------
# OARG -> ONNX
# ---------------------------------------------------------------------------------------------------------------
_ACT_NODES = {M.ACT_RELU: "Relu", M.ACT_HSWISH: "HardSwish", M.ACT_SIGMOID: "Sigmoid"}


def _parse_oarg(blob: bytes):
    if blob[:4] != M.MAGIC:
        raise OCRError("ModelLoad", "not an OARG blob")
    _version, kind, n_ops, n_tensors, n_w = struct.unpack_from("<4IQ", blob, 4)
    ops, off = [], 28
    for _ in range(n_ops):
        r = struct.unpack_from("<4i12i4f4q4q", blob, off)
        off += 144
        ops.append(dict(type=r[0], in0=r[1], in1=r[2], out=r[3], p=list(r[4:16]), f=list(r[16:20]), w_off=r[20:24],
                        w_len=r[24:28]))
    w = np.frombuffer(blob, np.float32, n_w, off)
    return kind, n_tensors, ops, w


def export_onnx(blob: bytes) -> bytes:
    kind, _n_t, ops, W = _parse_oarg(blob)
    nodes, inits = [], []
    name_of = {0: "x"}       # OARG tensor id -> ONNX value name
    slices = {}              # OARG tensor id -> {c_off: value name} for tensors assembled from channel slices
    totals = {}
    counter = [0]

    def fresh(prefix):
        counter[0] += 1
        return f"{prefix}_{counter[0]}"

    def init(prefix, arr):
        n = fresh(prefix)
        inits.append(tensor_proto(n, arr))
        return n

    def wt(op, i, shape):
        return W[op["w_off"][i]:op["w_off"][i] + op["w_len"][i]].reshape(shape)

    def value(tid):
        """the ONNX name of an OARG tensor; channel-slice tensors become a Concat the first time they are read"""
        if tid in slices and tid not in name_of:
            parts = [slices[tid][k] for k in sorted(slices[tid])]
            out = fresh("concat")
            nodes.append(node("Concat", parts, [out], axis=1))
            name_of[tid] = out
        return name_of[tid]

    def emit_act(x, act, f):
        if act not in _ACT_NODES and act not in (M.ACT_NONE, M.ACT_SWISH, M.ACT_HSIGMOID):
            raise OCRError("ModelLoad", f"cannot export activation {act} (spec-only so far)")
        if act in _ACT_NODES:
            y = fresh("act")
            nodes.append(node(_ACT_NODES[act], [x], [y]))
            x = y
        elif act == M.ACT_SWISH:
            s, y = fresh("sig"), fresh("act")
            nodes.append(node("Sigmoid", [x], [s]))
            nodes.append(node("Mul", [x, s], [y]))
            x = y
        elif act == M.ACT_HSIGMOID:
            y = fresh("act")
            nodes.append(node("HardSigmoid", [x], [y], alpha=1.0 / 6.0, beta=0.5))
            x = y
        if f[0] != 1.0 or f[1] != 0.0:  # the learnable affine of PP-LCNetV3's activation layers
            a, b = init("post_scale", np.array([f[0]], np.float32)), init("post_bias", np.array([f[1]], np.float32))
            m, y = fresh("aff"), fresh("aff")
            nodes.append(node("Mul", [x, a], [m]))
            nodes.append(node("Add", [m, b], [y]))
            x = y
        return x

    def store(op, y):
        p = op["p"]
        if p[11]:
            slices.setdefault(op["out"], {})[p[10]] = y
            totals[op["out"]] = p[11]
        else:
            name_of[op["out"]] = y

    for op in ops:
        t, p, f = op["type"], op["p"], op["f"]
        a = value(op["in0"])
        if t == M.OP_CONV:
            kh, kw, sh, sw, ph, pw, cin, cout, act = p[:9]
            w = init("w", np.ascontiguousarray(wt(op, 0, (cout, kh, kw, cin)).transpose(0, 3, 1, 2)))
            b = init("b", wt(op, 1, (cout,)))
            y = fresh("conv")
            nodes.append(node("Conv", [a, w, b], [y], kernel_shape=[kh, kw], strides=[sh, sw], pads=[ph, pw, ph, pw],
                              dilations=[1, 1], group=1))
            store(op, emit_act(y, act, f))
        elif t == M.OP_DWCONV:
            kh, kw, sh, sw, ph, pw, c, act = p[:8]
            w = init("w", np.ascontiguousarray(wt(op, 0, (kh, kw, c)).transpose(2, 0, 1).reshape(c, 1, kh, kw)))
            b = init("b", wt(op, 1, (c,)))
            y = fresh("dwconv")
            nodes.append(node("Conv", [a, w, b], [y], kernel_shape=[kh, kw], strides=[sh, sw], pads=[ph, pw, ph, pw],
                              dilations=[1, 1], group=c))
            name_of[op["out"]] = emit_act(y, act, f)
        elif t == M.OP_SE:
            c, cm, residual = p[:3]
            g, h1, r, h2, s, y = (fresh(n) for n in ("gap", "se_fc1", "se_relu", "se_fc2", "se_gate", "se"))
            nodes.append(node("GlobalAveragePool", [a], [g]))
            nodes.append(node("Conv", [g, init("w", wt(op, 0, (cm, c, 1, 1))), init("b", wt(op, 1, (cm,)))], [h1],
                              kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0], dilations=[1, 1], group=1))
            nodes.append(node("Relu", [h1], [r]))
            nodes.append(node("Conv", [r, init("w", wt(op, 2, (c, cm, 1, 1))), init("b", wt(op, 3, (c,)))], [h2],
                              kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0], dilations=[1, 1], group=1))
            nodes.append(node("HardSigmoid", [h2], [s], alpha=float(f[0]), beta=float(f[1])))
            nodes.append(node("Mul", [a, s], [y]))
            if residual:
                z = fresh("se_res")
                nodes.append(node("Add", [a, y], [z]))
                y = z
            name_of[op["out"]] = y
        elif t == M.OP_ADD:
            y = fresh("add")
            nodes.append(node("Add", [a, value(op["in1"])], [y]))
            name_of[op["out"]] = y
        elif t in (M.OP_UPADD, M.OP_UPSAMPLE):
            src = value(op["in1"]) if t == M.OP_UPADD else a
            y = src
            if p[0] != 1:
                y = fresh("up")
                sc = init("scales", np.array([1.0, 1.0, p[0], p[0]], np.float32))
                nodes.append(node("Resize", [src, "", sc], [y], mode="nearest", nearest_mode="floor",
                                  coordinate_transformation_mode="asymmetric"))
            if t == M.OP_UPADD:
                z = fresh("upadd")
                nodes.append(node("Add", [a, y], [z]))
                name_of[op["out"]] = z
            else:
                store(op, y)
        elif t == M.OP_DECONV2:
            cin, cout, act = p[:3]
            w = init("w", np.ascontiguousarray(wt(op, 0, (2, 2, cout, cin)).transpose(3, 2, 0, 1)))
            y = fresh("deconv")
            nodes.append(node("ConvTranspose", [a, w, init("b", wt(op, 1, (cout,)))], [y], kernel_shape=[2, 2],
                              strides=[2, 2], pads=[0, 0, 0, 0], dilations=[1, 1], group=1))
            name_of[op["out"]] = emit_act(y, act, [1.0, 0.0])
        elif t == M.OP_AVGPOOL:
            y = fresh("pool")
            if p[0] == 0 and p[1] == 0:  # global pool (classifier trunk)
                nodes.append(node("GlobalAveragePool", [a], [y]))
            else:
                nodes.append(node("AveragePool", [a], [y], kernel_shape=[p[0], p[1]], strides=[p[2], p[3]],
                                  pads=[0, 0, 0, 0]))
            name_of[op["out"]] = y
        elif t == M.OP_PAD:
            y = fresh("pad")
            pads = init("pads", np.array([0, 0, p[0], p[1], 0, 0, p[2], p[3]], np.int64))  # NCHW begins, then ends
            nodes.append(node("Pad", [a, pads], [y], mode="constant"))
            name_of[op["out"]] = y
        elif t == M.OP_MAXPOOL:
            y = fresh("maxpool")
            nodes.append(node("MaxPool", [a], [y], kernel_shape=[p[0], p[1]], strides=[p[2], p[3]], pads=[0, 0, 0, 0]))
            name_of[op["out"]] = y
        elif t == M.OP_LAYERNORM:
            c = p[0]
            n1, n2, y = fresh("nhwc"), fresh("ln"), fresh("nchw")
            nodes.append(node("Transpose", [a], [n1], perm=[0, 2, 3, 1]))
            nodes.append(node("LayerNormalization", [n1, init("ln_g", wt(op, 0, (c,))), init("ln_b", wt(op, 1, (c,)))],
                              [n2], axis=-1, epsilon=float(f[0])))
            nodes.append(node("Transpose", [n2], [y], perm=[0, 3, 1, 2]))
            name_of[op["out"]] = y
        elif t == M.OP_ATTN:
            c, heads = p[:2]
            if p[2]:
                raise OCRError("ModelLoad", "cannot export attention with positions on q / k (spec-only so far)")
            d = c // heads
            names = [fresh("attn") for _ in range(20)]
            (t0, s0, qm, qa, q5, qt, q, k, v, qs, kt, sc, sm, av, at, ar, pm, pa, r4, y) = names
            shp_btc = init("shape", np.array([0, -1, c], np.int64))
            shp_5 = init("shape", np.array([0, 0, 3, heads, d], np.int64))
            nodes.append(node("Transpose", [a], [t0], perm=[0, 2, 3, 1]))                 # [B,H,W,C]
            nodes.append(node("Shape", [t0], [s0]))                                       # kept for the way back
            bt, sbt = fresh("attn"), fresh("attn")
            nodes.append(node("Reshape", [t0, shp_btc], [bt]))                            # [B,T,C]
            nodes.append(node("Shape", [bt], [sbt]))
            nodes.append(node("MatMul", [bt, init("w", np.ascontiguousarray(wt(op, 0, (3 * c, c)).T))], [qm]))
            nodes.append(node("Add", [qm, init("b", wt(op, 1, (3 * c,)))], [qa]))
            nodes.append(node("Reshape", [qa, shp_5], [q5]))                              # [B,T,3,h,d]
            nodes.append(node("Transpose", [q5], [qt], perm=[2, 0, 3, 1, 4]))             # [3,B,h,T,d]
            nodes.append(node("Split", [qt], [q, k, v], axis=0))                          # each [1,B,h,T,d]
            nodes.append(node("Mul", [q, init("scale", np.array([f[0]], np.float32))], [qs]))
            nodes.append(node("Transpose", [k], [kt], perm=[0, 1, 2, 4, 3]))
            nodes.append(node("MatMul", [qs, kt], [sc]))
            nodes.append(node("Softmax", [sc], [sm], axis=-1))
            nodes.append(node("MatMul", [sm, v], [av]))                                   # [1,B,h,T,d]
            nodes.append(node("Transpose", [av], [at], perm=[0, 1, 3, 2, 4]))             # [1,B,T,h,d]
            nodes.append(node("Reshape", [at, sbt], [ar]))                                # [B,T,C]
            nodes.append(node("MatMul", [ar, init("w", np.ascontiguousarray(wt(op, 2, (c, c)).T))], [pm]))
            nodes.append(node("Add", [pm, init("b", wt(op, 3, (c,)))], [pa]))
            nodes.append(node("Reshape", [pa, s0], [r4]))                                 # [B,H,W,C]
            nodes.append(node("Transpose", [r4], [y], perm=[0, 3, 1, 2]))
            name_of[op["out"]] = y
        elif t == M.OP_CTC_HEAD:
            c, vocab = p[:2]
            t0, bt, mm, lg, y = (fresh("head") for _ in range(5))
            nodes.append(node("Transpose", [a], [t0], perm=[0, 2, 3, 1]))
            nodes.append(node("Reshape", [t0, init("shape", np.array([0, -1, c], np.int64))], [bt]))
            nodes.append(node("MatMul", [bt, init("w", np.ascontiguousarray(wt(op, 0, (vocab, c)).T))], [mm]))
            nodes.append(node("Add", [mm, init("b", wt(op, 1, (vocab,)))], [lg]))
            nodes.append(node("Softmax", [lg], [y], axis=2))
            name_of[op["out"]] = y
        else:
            raise OCRError("ModelLoad", f"cannot export op type {t}")
    out_name = value(ops[-1]["out"])
    if kind == M.KIND_DET:
        out_dims = ["N", 1, "H", "W"]
    elif kind == M.KIND_FEAT:
        out_dims = ["N", "C", "H", "W"]
    else:
        out_dims = ["N", "T", ops[-1]["p"][1]]
    return model_proto(nodes, inits, [value_info("x", ["N", 3, "H", "W"])], [value_info(out_name, out_dims)])


# ---------------------------------------------------------------------------------------------------------------
# ONNX -> node list
# ---------------------------------------------------------------------------------------------------------------
def read_model(data: bytes):
    """ModelProto -> (nodes, initializers, graph inputs, graph outputs); a node is a dict with op_type, inputs,
    outputs, attrs (python values)"""
    model = parse_message(data)
    if 7 not in model:
        raise OCRError("ModelLoad", "ONNX model has no graph")
    g = parse_message(model[7][0])
    inits = {}
    for raw in g.get(5, []):
        t = parse_message(raw)
        dims = _ints(t.get(1, []))
        dt = t.get(2, [FLOAT])[0]
        name = bytes(t[8][0]).decode() if 8 in t else ""
        if 9 in t:
            arr = np.frombuffer(bytes(t[9][0]), np.float32 if dt == FLOAT else np.int64)
        elif dt == FLOAT:
            arr = np.array(_floats(t.get(4, [])), np.float32)
        elif dt == INT64:
            arr = np.array(_ints(t.get(7, [])), np.int64)
        else:
            raise OCRError("ModelLoad", f"initializer {name}: unsupported data type {dt}")
        inits[name] = arr.reshape(dims) if dims else arr.reshape(())
    nodes = []
    for raw in g.get(1, []):
        n = parse_message(raw)
        attrs = {}
        for ra in n.get(5, []):
            a = parse_message(ra)
            an = bytes(a[1][0]).decode()
            if 2 in a:
                attrs[an] = struct.unpack("<f", a[2][0])[0]
            elif 3 in a:
                attrs[an] = _signed(a[3][0])
            elif 4 in a:
                attrs[an] = bytes(a[4][0]).decode()
            elif 8 in a:
                attrs[an] = _ints(a[8])
            elif 7 in a:
                attrs[an] = _floats(a[7])
            else:
                attrs[an] = None
        nodes.append(dict(op_type=bytes(n[4][0]).decode(), inputs=[bytes(i).decode() for i in n.get(1, [])],
                          outputs=[bytes(o).decode() for o in n.get(2, [])], attrs=attrs,
                          name=bytes(n[3][0]).decode() if 3 in n else ""))

    def names(field):
        return [bytes(parse_message(v)[1][0]).decode() for v in g.get(field, [])]
    return nodes, inits, [n for n in names(11) if n not in inits], names(12)


# ---------------------------------------------------------------------------------------------------------------
# ONNX -> OARG
# ---------------------------------------------------------------------------------------------------------------
def import_onnx(data: bytes, seed_kind: int | None = None) -> bytes:
    nodes, inits, g_inputs, g_outputs = read_model(data)
    if len(g_inputs) != 1:
        raise OCRError("ModelLoad", f"expected one graph input, found {g_inputs}")
    consumers = {}
    for i, n in enumerate(nodes):
        for x in n["inputs"]:
            consumers.setdefault(x, []).append(i)
    for o in g_outputs:
        consumers.setdefault(o, []).append(-1)
    g = M.GraphBuilder(M.KIND_DET, 0)
    tid = {g_inputs[0]: 0}   # ONNX value name -> OARG tensor id
    done = set()
    kind = M.KIND_DET

    def fail(n, why):
        raise OCRError("ModelLoad", f"ONNX import: node '{n['name']}' ({n['op_type']}): {why}")

    def sole_consumer(name, op_type=None):
        c = consumers.get(name, [])
        if len(c) != 1 or c[0] < 0 or c[0] in done:
            return None
        n = nodes[c[0]]
        return c[0] if op_type is None or n["op_type"] == op_type else None

    def take_activation(y):
        """consume the activation (and learnable affine) that follows value y; returns (act, post, new value name)"""
        act, post = M.ACT_NONE, (1.0, 0.0)
        i = sole_consumer(y)
        if i is not None:
            n = nodes[i]
            simple = {"Relu": M.ACT_RELU, "HardSwish": M.ACT_HSWISH, "Sigmoid": M.ACT_SIGMOID}
            if n["op_type"] in simple and not (n["op_type"] == "Sigmoid" and _feeds_mul_with(n, y)):
                act, y = simple[n["op_type"]], n["outputs"][0]
                done.add(i)
            elif n["op_type"] == "HardSigmoid" and not _feeds_mul_with(n, y):
                if abs(n["attrs"].get("alpha", 0.2) - 1.0 / 6.0) > 1e-6 or abs(n["attrs"].get("beta", 0.5) - 0.5) > 1e-6:
                    fail(n, "HardSigmoid activation with alpha/beta other than 1/6, 0.5")
                act, y = M.ACT_HSIGMOID, n["outputs"][0]
                done.add(i)
        if act == M.ACT_NONE:  # decomposed forms: x * Sigmoid(x) (swish), x * HardSigmoid(x; 1/6, 0.5) (hardswish)
            c = [k for k in consumers.get(y, []) if k >= 0 and k not in done]
            if len(c) == 2:
                gate = [k for k in c if nodes[k]["op_type"] in ("Sigmoid", "HardSigmoid")]
                mul = [k for k in c if nodes[k]["op_type"] == "Mul"]
                if len(gate) == 1 and len(mul) == 1 and set(nodes[mul[0]]["inputs"]) == {y, nodes[gate[0]]["outputs"][0]}:
                    gn = nodes[gate[0]]
                    if gn["op_type"] == "Sigmoid":
                        act = M.ACT_SWISH
                    else:
                        if abs(gn["attrs"].get("alpha", 0.2) - 1.0 / 6.0) > 1e-6 or abs(gn["attrs"].get("beta", 0.5) - 0.5) > 1e-6:
                            fail(gn, "decomposed hardswish with alpha/beta other than 1/6, 0.5")
                        act = M.ACT_HSWISH
                    done.update((gate[0], mul[0]))
                    y = nodes[mul[0]]["outputs"][0]
        # learnable affine: Mul by a scalar initializer then Add of a scalar initializer
        i = sole_consumer(y, "Mul")
        if i is not None and act != M.ACT_NONE:
            other = [x for x in nodes[i]["inputs"] if x != y]
            if len(other) == 1 and other[0] in inits and inits[other[0]].size == 1:
                j = sole_consumer(nodes[i]["outputs"][0], "Add")
                if j is not None:
                    ob = [x for x in nodes[j]["inputs"] if x != nodes[i]["outputs"][0]]
                    if len(ob) == 1 and ob[0] in inits and inits[ob[0]].size == 1:
                        post = (float(inits[other[0]].reshape(-1)[0]), float(inits[ob[0]].reshape(-1)[0]))
                        done.update((i, j))
                        y = nodes[j]["outputs"][0]
        return act, post, y

    def _feeds_mul_with(n, x):
        i = sole_consumer(n["outputs"][0], "Mul")
        return i is not None and x in nodes[i]["inputs"]

    def conv_params(n):
        a = n["attrs"]
        w = inits.get(n["inputs"][1])
        if w is None:
            fail(n, "weights are not an initializer")
        kh, kw = a.get("kernel_shape", list(w.shape[2:]))
        sh, sw = a.get("strides", [1, 1])
        pads = a.get("pads", [0, 0, 0, 0])
        auto_pad = a.get("auto_pad", "NOTSET") or "NOTSET"
        if auto_pad == "VALID":
            pads = [0, 0, 0, 0]
        elif auto_pad in ("SAME_UPPER", "SAME_LOWER"):
            # total = (ceil(in / s) - 1) * s + k - in depends on the input size unless s == 1 (then k - 1); symmetric only
            # for odd kernels -- anything else cannot be written as the fixed symmetric padding the engine applies
            if (sh, sw) != (1, 1) or kh % 2 == 0 or kw % 2 == 0:
                fail(n, f"auto_pad={auto_pad} with stride {sh}x{sw} / kernel {kh}x{kw} (needs stride 1 and an odd kernel)")
            pads = [kh // 2, kw // 2, kh // 2, kw // 2]
        elif auto_pad != "NOTSET":
            fail(n, f"unknown auto_pad '{auto_pad}'")
        if pads[0] != pads[2] or pads[1] != pads[3]:
            fail(n, "asymmetric padding")
        if any(d != 1 for d in a.get("dilations", [1, 1])):
            fail(n, "dilation")
        b = inits[n["inputs"][2]] if len(n["inputs"]) > 2 and n["inputs"][2] else np.zeros(w.shape[0], np.float32)
        return w.astype(np.float32), b.astype(np.float32).copy(), (kh, kw), (sh, sw), (pads[0], pads[1]), a.get("group", 1)

    def fold_bn(y, w, b):
        i = sole_consumer(y, "BatchNormalization")
        if i is None:
            return y, w, b
        n = nodes[i]
        sc, bi, mean, var = (inits[x].astype(np.float32) for x in n["inputs"][1:5])
        k = sc / np.sqrt(var + np.float32(n["attrs"].get("epsilon", 1e-5)))
        done.add(i)
        return n["outputs"][0], w * k.reshape(-1, 1, 1, 1), (b - mean) * k + bi

    # Concat inputs are written as channel slices of one tensor: plan them before walking the graph
    slice_of = {}  # value name -> (concat output name, c_off)
    concat_total = {}
    for n in nodes:
        if n["op_type"] == "Concat":
            if n["attrs"].get("axis", 1) != 1:
                fail(n, "Concat on an axis other than channels")
            concat_total[n["outputs"][0]] = None
    i = 0
    while i < len(nodes):
        n = nodes[i]
        if i in done:
            i += 1
            continue
        done.add(i)
        ot, ins, outs = n["op_type"], n["inputs"], n["outputs"]
        if ot == "Conv":
            w, b, k, s, pad, group = conv_params(n)
            y, w, b = fold_bn(outs[0], w, b)
            x = tid[ins[0]]
            cin = g.channels[x]
            if group == 1:
                act, post, y = take_activation(y)
                tid[y] = g.conv(x, w.shape[0], k, s, pad, act=act, post=post,
                                w=np.ascontiguousarray(w.transpose(0, 2, 3, 1)), b=b)
            elif group == cin and w.shape[0] == cin and w.shape[1] == 1 and k[0] == k[1] and pad == (k[0] // 2, k[0] // 2):
                act, post, y = take_activation(y)
                tid[y] = g.dwconv(x, k[0], s, act=act, post=post,
                                  w=np.ascontiguousarray(w.reshape(cin, k[0], k[1]).transpose(1, 2, 0)), b=b)
            else:
                fail(n, f"grouped convolution (group={group}) other than depthwise")
        elif ot == "GlobalAveragePool":
            # squeeze-excite: GAP -> Conv1x1 -> Relu -> Conv1x1 -> HardSigmoid -> Mul(x, gate) [-> Add(x, .)]
            # anything else is a plain global pool (the classifier trunk's): AVGPOOL with a (0, 0) window
            seq, cur = [], outs[0]
            for want in ("Conv", "Relu", "Conv", "HardSigmoid", "Mul"):
                j = sole_consumer(cur, want)
                if j is None:
                    seq = None
                    break
                seq.append(j)
                cur = nodes[j]["outputs"][0]
            if seq is None:
                tid[outs[0]] = g.avgpool(tid[ins[0]], (0, 0), (0, 0))
                i += 1
                continue
            c1, c2, hs, mul = nodes[seq[0]], nodes[seq[2]], nodes[seq[3]], nodes[seq[4]]
            if ins[0] not in mul["inputs"]:
                fail(n, "squeeze-excite gate does not multiply the pooled tensor")
            done.update(seq)
            def se_fc(c):
                if c["inputs"][1] not in inits:
                    fail(c, "squeeze-excite weights are not an initializer")
                w = inits[c["inputs"][1]]
                if w.ndim != 4 or w.shape[2:] != (1, 1) or c["attrs"].get("group", 1) != 1:
                    fail(c, "squeeze-excite convolutions must be dense 1x1")
                has_b = len(c["inputs"]) > 2 and c["inputs"][2]
                if has_b and c["inputs"][2] not in inits:
                    fail(c, "squeeze-excite bias is not an initializer")
                return w, (inits[c["inputs"][2]] if has_b else np.zeros(w.shape[0], np.float32))
            (w1, b1), (w2, b2) = se_fc(c1), se_fc(c2)
            residual, y = False, mul["outputs"][0]
            j = sole_consumer(y, "Add")
            if j is not None and ins[0] in nodes[j]["inputs"]:
                residual, y = True, nodes[j]["outputs"][0]
                done.add(j)
            x = tid[ins[0]]
            c, cm = g.channels[x], w1.shape[0]
            out = g.new_tensor(c)
            g.ops.append(M.Op(M.OP_SE, x, -1, out, [c, cm, 1 if residual else 0] + [0] * 9,
                              [float(hs["attrs"].get("alpha", 0.2)), float(hs["attrs"].get("beta", 0.5)), 0.0, 0.0],
                              [w1.reshape(cm, c).astype(np.float32), b1.astype(np.float32),
                               w2.reshape(c, cm).astype(np.float32), b2.astype(np.float32)]))
            tid[y] = out
        elif ot == "Add":
            for x in ins:
                if x not in tid:
                    fail(n, f"operand '{x}' is " + ("an initializer (only activation + activation adds are supported)"
                                                   if x in inits else "not produced by a supported node"))
            tid[outs[0]] = g.add(tid[ins[0]], tid[ins[1]])
        elif ot == "Resize":
            if n["attrs"].get("mode", "nearest") != "nearest":
                fail(n, "Resize mode other than nearest")
            # integer-scale pixel replication out[y] = in[y // s]: 'asymmetric' with floor, or any of the modes that
            # coincide with it for integer scales.  half_pixel/pytorch_half_pixel + round_prefer_floor (the ONNX
            # defaults) also replicate (src = (y + .5)/s - .5 lands strictly inside cell y // s for integer s), and so
            # does half_pixel + floor only for s == 1; everything else is rejected.
            ctm = n["attrs"].get("coordinate_transformation_mode", "half_pixel")
            nm = n["attrs"].get("nearest_mode", "round_prefer_floor")
            ok = (ctm == "asymmetric" and nm == "floor") or \
                 (ctm in ("half_pixel", "pytorch_half_pixel") and nm in ("round_prefer_floor", "round_prefer_ceil"))
            if not ok:
                fail(n, f"Resize nearest with coordinate_transformation_mode={ctm}, nearest_mode={nm} is not "
                        "integer pixel replication")
            sc = inits.get(ins[2]) if len(ins) > 2 and ins[2] else None
            if sc is None or sc.size != 4 or sc[0] != 1 or sc[1] != 1 or sc[2] != sc[3] or sc[2] != int(sc[2]):
                fail(n, "Resize needs constant integer scales [1,1,s,s]")
            s = int(sc[2])
            j = sole_consumer(outs[0], "Add")
            if j is not None:
                other = [x for x in nodes[j]["inputs"] if x != outs[0]][0]
                done.add(j)
                tid[nodes[j]["outputs"][0]] = g.upadd(tid[other], tid[ins[0]], s)
            else:
                x = tid[ins[0]]
                out = g.new_tensor(g.channels[x])
                g.upsample_into(x, s, out, 0, 0)
                tid[outs[0]] = out
        elif ot == "Concat":
            parts = [tid[x] for x in ins]
            total = sum(g.channels[p] for p in parts)
            out = g.new_tensor(total)
            off = 0
            for pt in parts:
                # re-target the producer when it can write a slice itself (conv / upsample that nobody else reads)
                prod = [o for o in g.ops if o.out == pt]
                name = [x for x in ins if tid[x] == pt][0]
                only_here = len(consumers.get(name, [])) == 1 and sum(1 for o in g.ops if pt in (o.in0, o.in1)) == 0
                if len(prod) == 1 and prod[0].type in (M.OP_CONV, M.OP_UPSAMPLE) and only_here and prod[0].p[11] == 0:
                    prod[0].out, prod[0].p[10], prod[0].p[11] = out, off, total
                else:
                    g.upsample_into(pt, 1, out, off, total)
                off += g.channels[pt]
            tid[outs[0]] = out
        elif ot == "ConvTranspose":
            w = inits[ins[1]].astype(np.float32)  # [cin][cout][kh][kw]
            a = n["attrs"]
            if list(w.shape[2:]) != [2, 2] or a.get("strides", [1, 1]) != [2, 2] or any(a.get("pads", [0] * 4)):
                fail(n, "ConvTranspose other than 2x2 stride 2 without padding")
            b = inits[ins[2]].astype(np.float32) if len(ins) > 2 and ins[2] else np.zeros(w.shape[1], np.float32)
            y = outs[0]
            j = sole_consumer(y, "BatchNormalization")
            if j is not None:
                bn = nodes[j]
                sc, bi, mean, var = (inits[x].astype(np.float32) for x in bn["inputs"][1:5])
                kk = sc / np.sqrt(var + np.float32(bn["attrs"].get("epsilon", 1e-5)))
                w, b, y = w * kk.reshape(1, -1, 1, 1), (b - mean) * kk + bi, bn["outputs"][0]
                done.add(j)
            act, post, y = take_activation(y)
            if post != (1.0, 0.0):
                fail(n, "affine after a transposed convolution")
            tid[y] = g.deconv2(tid[ins[0]], w.shape[1], act=act, w=np.ascontiguousarray(w.transpose(2, 3, 1, 0)), b=b)
        elif ot == "Pad":
            a = n["attrs"]
            if a.get("mode", "constant") not in ("constant", b"constant"):
                fail(n, "Pad mode other than constant")
            if len(ins) < 2 or ins[1] not in inits:
                fail(n, "Pad amounts must be an initializer")
            pads = [int(v) for v in inits[ins[1]].ravel()]
            if len(pads) != 8 or any(pads[k] for k in (0, 1, 4, 5)) or min(pads) < 0:
                fail(n, "Pad must add zeros to the spatial dims of an NCHW tensor")
            if len(ins) > 2 and ins[2] and (ins[2] not in inits or float(inits[ins[2]].ravel()[0]) != 0.0):
                fail(n, "Pad with a non-zero constant")
            tid[outs[0]] = g.pad(tid[ins[0]], pads[2], pads[3], pads[6], pads[7])
        elif ot == "MaxPool":
            a = n["attrs"]
            if any(a.get("pads", [0] * 4)) or a.get("ceil_mode", 0) or any(d != 1 for d in a.get("dilations", [1, 1])):
                fail(n, "MaxPool with padding, ceil_mode or dilation")
            tid[outs[0]] = g.maxpool(tid[ins[0]], tuple(a["kernel_shape"]), tuple(a.get("strides", [1, 1])))
        elif ot == "AveragePool":
            a = n["attrs"]
            if any(a.get("pads", [0] * 4)):
                fail(n, "padded AveragePool")
            tid[outs[0]] = g.avgpool(tid[ins[0]], tuple(a["kernel_shape"]), tuple(a.get("strides", a["kernel_shape"])))
        elif ot == "Transpose" and n["attrs"].get("perm") == [0, 2, 3, 1]:
            j = sole_consumer(outs[0])
            nxt = nodes[j]["op_type"] if j is not None else None
            if nxt == "LayerNormalization":
                ln = nodes[j]
                k2 = sole_consumer(ln["outputs"][0], "Transpose")
                if k2 is None or nodes[k2]["attrs"].get("perm") != [0, 3, 1, 2]:
                    fail(ln, "LayerNormalization must sit between NHWC/NCHW transposes")
                done.update((j, k2))
                x = tid[ins[0]]
                c = g.channels[x]
                out = g.new_tensor(c)
                g.ops.append(M.Op(M.OP_LAYERNORM, x, -1, out, [c] + [0] * 11,
                                  [float(ln["attrs"].get("epsilon", 1e-5)), 0.0, 0.0, 0.0],
                                  [inits[ln["inputs"][1]].astype(np.float32), inits[ln["inputs"][2]].astype(np.float32)]))
                tid[nodes[k2]["outputs"][0]] = out
            elif len(consumers.get(outs[0], [])) == 2:
                i = _import_attention(g, nodes, inits, consumers, done, tid, i, fail)
            elif nxt == "Reshape":
                kind = M.KIND_REC
                _import_ctc_head(g, nodes, inits, done, tid, i, fail)
            else:
                fail(n, "NHWC transpose outside LayerNormalization / attention / CTC head")
        elif ot in ("Dropout", "Identity"):  # identities at inference time
            tid[outs[0]] = tid[ins[0]]
        elif ot in ("Flatten", "Squeeze", "Reshape"):
            # classifier tail as PaddleClas / torch exports spell it: pooled [B,C,1,1] -> Flatten | Reshape | Squeeze
            # -> Gemm | MatMul [+ Add] -> Softmax
            kind = M.KIND_CLS
            _import_cls_head(g, nodes, inits, consumers, done, tid, i, fail)
        else:
            fail(n, "operator outside the supported subset")
        i += 1
    if g_outputs[0] not in tid:
        raise OCRError("ModelLoad", "ONNX import: graph output was not produced")
    if g.ops[-1].out != tid[g_outputs[0]]:
        raise OCRError("ModelLoad", "ONNX import: the graph output is not the last operation")
    g.kind = kind if seed_kind is None else seed_kind
    return g.serialize()


def _chain(nodes, consumers, done, start_value, types, fail, ctx):
    """follow single-consumer links from start_value through the given op types; returns the node indices"""
    seq, cur = [], start_value
    for want in types:
        c = [k for k in consumers.get(cur, []) if k >= 0 and k not in done and nodes[k]["op_type"] == want]
        if len(c) != 1:
            fail(ctx, f"expected {want} after '{cur}'")
        seq.append(c[0])
        cur = nodes[c[0]]["outputs"][0]
    return seq


def _import_attention(g, nodes, inits, consumers, done, tid, i, fail):
    """the attention block as export_onnx writes it (single input, fused qkv projection, `heads` from the reshape)"""
    n = nodes[i]
    t0 = n["outputs"][0]
    shape_i = [k for k in consumers[t0] if nodes[k]["op_type"] == "Shape"]
    if len(shape_i) != 1:
        fail(n, "attention: missing Shape of the NHWC tensor")
    seq = _chain(nodes, consumers, done, t0, ["Reshape", "MatMul", "Add", "Reshape", "Transpose", "Split"], fail, n)
    shape_i += [k for k in consumers[nodes[seq[0]]["outputs"][0]] if nodes[k]["op_type"] == "Shape"]
    split = nodes[seq[-1]]
    q, k, v = split["outputs"]
    qs = _chain(nodes, consumers, done, q, ["Mul"], fail, n)
    kt = _chain(nodes, consumers, done, k, ["Transpose"], fail, n)
    tail = _chain(nodes, consumers, done, nodes[qs[0]]["outputs"][0],
                  ["MatMul", "Softmax", "MatMul", "Transpose", "Reshape", "MatMul", "Add", "Reshape", "Transpose"], fail, n)
    wqkv, bqkv = inits[nodes[seq[1]]["inputs"][1]], inits[[x for x in nodes[seq[2]]["inputs"] if x in inits][0]]
    shp5 = inits[nodes[seq[3]]["inputs"][1]]
    heads = int(shp5[3])
    scale = float(inits[[x for x in nodes[qs[0]]["inputs"] if x in inits][0]].reshape(-1)[0])
    wp, bp = inits[nodes[tail[5]]["inputs"][1]], inits[[x for x in nodes[tail[6]]["inputs"] if x in inits][0]]
    done.update(shape_i + seq + qs + kt + tail)
    x = tid[n["inputs"][0]]
    c = g.channels[x]
    out = g.new_tensor(c)
    g.ops.append(M.Op(M.OP_ATTN, x, -1, out, [c, heads] + [0] * 10, [scale, 0, 0, 0],
                      [np.ascontiguousarray(wqkv.T).astype(np.float32), bqkv.astype(np.float32),
                       np.ascontiguousarray(wp.T).astype(np.float32), bp.astype(np.float32)]))
    tid[nodes[tail[-1]]["outputs"][0]] = out
    return i


def _import_ctc_head(g, nodes, inits, done, tid, i, fail):
    n = nodes[i]
    consumers = {}
    for k, m in enumerate(nodes):
        for x in m["inputs"]:
            consumers.setdefault(x, []).append(k)
    seq = _chain(nodes, consumers, done, n["outputs"][0], ["Reshape", "MatMul", "Add", "Softmax"], fail, n)
    w = inits[nodes[seq[1]]["inputs"][1]]
    b = inits[[x for x in nodes[seq[2]]["inputs"] if x in inits][0]]
    done.update(seq)
    x = tid[n["inputs"][0]]
    tid[nodes[seq[-1]]["outputs"][0]] = g.ctc_head(x, w.shape[1], np.ascontiguousarray(w.T).astype(np.float32),
                                                   b.astype(np.float32))


def _import_cls_head(g, nodes, inits, consumers, done, tid, i, fail):
    """Flatten | Reshape | Squeeze -> Gemm | MatMul [+ Add] -> Softmax on pooled [B,C,1,1] features: the Linear + softmax
    head op on a one-step sequence (the engine checks at run time that the spatial size really is 1 x 1)"""
    n = nodes[i]

    def sole(value, types):
        c = [k for k in consumers.get(value, []) if k >= 0 and k not in done]
        return c[0] if len(c) == 1 and nodes[c[0]]["op_type"] in types else None

    x = tid[n["inputs"][0]]
    j = sole(n["outputs"][0], ("Gemm", "MatMul"))
    if j is None:
        fail(n, "flattened features must feed exactly one Gemm / MatMul")
    m = nodes[j]
    if m["inputs"][1] not in inits:
        fail(m, "classifier weights must be an initializer")
    w = inits[m["inputs"][1]].astype(np.float32)
    seq, cur, b = [j], m["outputs"][0], None
    if m["op_type"] == "Gemm":
        a = m["attrs"]
        if a.get("alpha", 1.0) != 1.0 or a.get("beta", 1.0) != 1.0 or a.get("transA", 0):
            fail(m, "Gemm with alpha/beta != 1 or transA")
        w = w if a.get("transB", 0) else w.T  # -> [classes, C]
        if len(m["inputs"]) > 2 and m["inputs"][2]:
            b = inits[m["inputs"][2]].astype(np.float32)
    else:
        w = w.T
        k = sole(cur, ("Add",))
        if k is not None:
            bias = [v for v in nodes[k]["inputs"] if v in inits]
            if len(bias) != 1:
                fail(nodes[k], "classifier bias must be an initializer")
            b = inits[bias[0]].astype(np.float32)
            seq.append(k)
            cur = nodes[k]["outputs"][0]
    k = sole(cur, ("Softmax",))
    if k is None:
        fail(m, "classifier head must end in Softmax")
    if nodes[k]["attrs"].get("axis", -1) not in (1, -1):
        fail(nodes[k], "Softmax over an axis other than the classes")
    seq.append(k)
    if w.ndim != 2 or w.shape[1] != g.channels[x]:
        fail(m, f"classifier weights {w.shape} do not match the {g.channels[x]} pooled channels")
    if b is None:
        b = np.zeros((w.shape[0],), np.float32)
    done.update(seq)
    tid[nodes[k]["outputs"][0]] = g.ctc_head(x, w.shape[0], np.ascontiguousarray(w), b.reshape(-1))


def main(argv):
    if len(argv) != 4 or argv[1] not in ("convert", "export"):
        print(__doc__)
        return 2
    with open(argv[2], "rb") as f:
        data = f.read()
    out = import_onnx(data) if argv[1] == "convert" else export_onnx(data)
    with open(argv[3], "wb") as f:
        f.write(out)
    print(f"{argv[1]}: {argv[2]} ({len(data)} bytes) -> {argv[3]} ({len(out)} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
