"""ONNX <-> OARG conversion (oar_ocr_b200/onnx_io.py) on CPU: protobuf wire format, export of both synthetic networks
evaluated with the ONNX operator definitions against the oracle, import round trips, and the alternative spellings an
exporter may use (BatchNormalization after Conv, decomposed hardswish / swish)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oar_ocr_b200 import models, onnx_io
from oar_ocr_b200.ffi import OCRError
from oracle.net import OracleNet


def _run_onnx(data: bytes, x: np.ndarray) -> np.ndarray:
    """a small evaluator of the ONNX operator subset (definitions as in the ONNX operator docs), fp32 torch-CPU"""
    nodes, inits, g_in, g_out = onnx_io.read_model(data)
    v = {k: torch.from_numpy(np.array(a)) for k, a in inits.items()}
    v[g_in[0]] = torch.from_numpy(x)
    for n in nodes:
        t, a = n["op_type"], n["attrs"]
        i = [v[k] if k else None for k in n["inputs"]]
        if t == "Conv":
            p = a["pads"]
            y = F.conv2d(i[0], i[1], i[2] if len(i) > 2 else None, a["strides"], (p[0], p[1]), a["dilations"], a["group"])
        elif t == "ConvTranspose":
            y = F.conv_transpose2d(i[0], i[1], i[2], a["strides"])
        elif t == "BatchNormalization":
            y = F.batch_norm(i[0], i[3], i[4], i[1], i[2], False, 0.0, a.get("epsilon", 1e-5))
        elif t == "Relu":
            y = F.relu(i[0])
        elif t == "HardSwish":
            y = F.hardswish(i[0])
        elif t == "Sigmoid":
            y = torch.sigmoid(i[0])
        elif t == "HardSigmoid":
            y = torch.clamp(i[0] * a.get("alpha", 0.2) + a.get("beta", 0.5), 0.0, 1.0)
        elif t == "Mul":
            y = i[0] * i[1]
        elif t == "Add":
            y = i[0] + i[1]
        elif t == "GlobalAveragePool":
            y = i[0].mean(dim=(2, 3), keepdim=True)
        elif t == "AveragePool":
            y = F.avg_pool2d(i[0], a["kernel_shape"], a["strides"])
        elif t == "MaxPool":
            y = F.max_pool2d(i[0], a["kernel_shape"], a["strides"])
        elif t == "Pad":
            pd = [int(q) for q in i[1]]
            y = F.pad(i[0], (pd[3], pd[7], pd[2], pd[6]))
        elif t == "Resize":
            assert a["mode"] == "nearest"
            y = F.interpolate(i[0], scale_factor=float(i[2][2]), mode="nearest")
        elif t == "Concat":
            y = torch.cat(i, dim=a["axis"])
        elif t == "Transpose":
            y = i[0].permute(a["perm"])
        elif t == "Shape":
            y = torch.tensor(list(i[0].shape), dtype=torch.int64)
        elif t == "Reshape":
            shp = [int(s) for s in i[1]]
            shp = [i[0].shape[k] if s == 0 else s for k, s in enumerate(shp)]
            y = i[0].reshape(shp)
        elif t == "LayerNormalization":
            assert a["axis"] == -1
            y = F.layer_norm(i[0], i[0].shape[-1:], i[1], i[2], a["epsilon"])
        elif t == "MatMul":
            y = i[0] @ i[1]
        elif t == "Flatten":
            y = i[0].flatten(a.get("axis", 1))
        elif t == "Squeeze":
            y = i[0].squeeze(-1).squeeze(-1)
        elif t in ("Dropout", "Identity"):
            y = i[0]
        elif t == "Gemm":
            y = i[0] @ (i[1].T if a.get("transB", 0) else i[1])
            if len(i) > 2 and i[2] is not None:
                y = y + i[2]
        elif t == "Softmax":
            y = torch.softmax(i[0], dim=a["axis"])
        elif t == "Split":
            parts = torch.chunk(i[0], len(n["outputs"]), dim=a["axis"])
            for name, part in zip(n["outputs"], parts):
                v[name] = part
            continue
        else:
            raise AssertionError("evaluator: " + t)
        v[n["outputs"][0]] = y
    return v[g_out[0]].numpy()


def test_wire_format_roundtrip():
    msg = onnx_io._f_varint(1, 300) + onnx_io._f_bytes(2, "abc") + onnx_io._f_float(3, 1.5) + onnx_io._f_varint(4, -1)
    got = onnx_io.parse_message(msg)
    assert got[1] == [300] and bytes(got[2][0]) == b"abc" and onnx_io._floats(got[3]) == [1.5]
    assert onnx_io._ints(got[4]) == [-1]
    t = onnx_io.parse_message(onnx_io.tensor_proto("w", np.arange(6, dtype=np.float32).reshape(2, 3)))
    assert onnx_io._ints(t[1]) == [2, 3] and bytes(t[8][0]) == b"w"
    n = onnx_io.node("Conv", ["x", "w"], ["y"], kernel_shape=[3, 3], group=1, epsilon=0.5, mode="nearest")
    nodes, inits, ins, outs = onnx_io.read_model(onnx_io.model_proto([n], [], [onnx_io.value_info("x", ["N", 3, 8, 8])],
                                                                    [onnx_io.value_info("y", [1])]))
    assert nodes[0]["op_type"] == "Conv" and nodes[0]["attrs"] == dict(kernel_shape=[3, 3], group=1, epsilon=0.5,
                                                                        mode="nearest")
    assert ins == ["x"] and outs == ["y"] and nodes[0]["inputs"] == ["x", "w"]


@pytest.mark.parametrize("kind", ["det", "rec", "cls", "hgnetv2"])
def test_export_matches_oracle_and_import_roundtrips(kind):
    if kind == "hgnetv2":  # two stages of the layout / server backbone: stem with Pad + MaxPool, concat blocks
        blob = models.build_hgnetv2_l(return_idx=(1,))
    else:
        blob = models.get_blob(kind, vocab=97) if kind == "rec" else models.get_blob(kind)
    rng = np.random.default_rng(3)
    shape = {"det": (2, 3, 64, 96), "rec": (2, 3, 48, 64), "cls": (2, 3, 80, 160), "hgnetv2": (1, 3, 64, 96)}[kind]
    x = rng.standard_normal(shape).astype(np.float32)
    want = OracleNet(blob).forward(x)
    data = onnx_io.export_onnx(blob)
    # 1. the exported ONNX, evaluated operator by operator, is the same network
    got = _run_onnx(data, x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5
    # 2. importing it back gives a graph with the same ops, parameters and weights: identical oracle output
    # (a classifier ends in the same MatMul + Softmax pattern as a CTC head: the caller's role names the kind)
    back = onnx_io.import_onnx(data, {"cls": models.KIND_CLS, "hgnetv2": models.KIND_FEAT}.get(kind))
    k0, _, ops0, w0 = onnx_io._parse_oarg(blob)
    k1, _, ops1, w1 = onnx_io._parse_oarg(back)
    assert k0 == k1 and sorted(o["type"] for o in ops0) == sorted(o["type"] for o in ops1)
    if [o["type"] for o in ops0] == [o["type"] for o in ops1]:  # (a Concat may place its copy-into-slice op later)
        for a, b in zip(ops0, ops1):
            assert a["p"][:10] == b["p"][:10] and a["f"] == pytest.approx(b["f"], rel=1e-7)
            for i in range(4):
                assert np.array_equal(w0[a["w_off"][i]:a["w_off"][i] + a["w_len"][i]],
                                      w1[b["w_off"][i]:b["w_off"][i] + b["w_len"][i]])
    assert np.array_equal(OracleNet(back).forward(x), want)


def test_import_folds_batchnorm_and_decomposed_activations():
    rng = np.random.default_rng(9)
    w = rng.standard_normal((8, 3, 3, 3)).astype(np.float32) * 0.2
    bn = [rng.uniform(0.5, 1.5, 8).astype(np.float32), rng.standard_normal(8).astype(np.float32) * 0.1,
          rng.standard_normal(8).astype(np.float32) * 0.1, rng.uniform(0.5, 1.5, 8).astype(np.float32)]
    w2 = rng.standard_normal((1, 8, 1, 1)).astype(np.float32) * 0.3
    T, N = onnx_io.tensor_proto, onnx_io.node
    nodes = [
        N("Conv", ["x", "w"], ["c"], kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1], dilations=[1, 1], group=1),
        N("BatchNormalization", ["c", "s", "b", "m", "v"], ["n"], epsilon=1e-5),
        N("HardSigmoid", ["n"], ["g"], alpha=1.0 / 6.0, beta=0.5),   # hardswish spelled as x * hardsigmoid(x)
        N("Mul", ["n", "g"], ["h"]),
        N("Conv", ["h", "w2", "b2"], ["c2"], kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0], dilations=[1, 1],
          group=1),
        N("Sigmoid", ["c2"], ["sg"]),                                # swish spelled as x * sigmoid(x)
        N("Mul", ["c2", "sg"], ["y"]),
    ]
    inits = [T("w", w), T("s", bn[0]), T("b", bn[1]), T("m", bn[2]), T("v", bn[3]), T("w2", w2),
             T("b2", np.array([0.05], np.float32))]
    data = onnx_io.model_proto(nodes, inits, [onnx_io.value_info("x", ["N", 3, "H", "W"])],
                               [onnx_io.value_info("y", ["N", 1, "H", "W"])])
    x = rng.standard_normal((2, 3, 16, 20)).astype(np.float32)
    want = _run_onnx(data, x)
    blob = onnx_io.import_onnx(data)
    _, _, ops, _ = onnx_io._parse_oarg(blob)
    assert [(o["type"], o["p"][8]) for o in ops] == [(models.OP_CONV, models.ACT_HSWISH), (models.OP_CONV, models.ACT_SWISH)]
    assert np.abs(OracleNet(blob).forward(x) - want).max() <= 1e-5


@pytest.mark.parametrize("tail", ["flatten_gemm", "reshape_matmul_add", "squeeze_matmul"])
def test_import_classifier_tail_spellings(tail):
    """a PaddleClas-style classifier: conv trunk -> GlobalAveragePool -> 1x1 conv + HardSwish -> Dropout -> flatten ->
    fc -> Softmax, with the spellings exporters use for flatten and fc; imported as kind CLS, same probabilities"""
    rng = np.random.default_rng(4)
    T, N = onnx_io.tensor_proto, onnx_io.node
    w = rng.standard_normal((8, 3, 3, 3)).astype(np.float32) * 0.3
    w2 = rng.standard_normal((16, 8, 1, 1)).astype(np.float32) * 0.4
    fc = rng.standard_normal((2, 16)).astype(np.float32)
    fb = np.array([0.1, -0.2], np.float32)
    nodes = [
        N("Conv", ["x", "w", "b"], ["c"], kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1], dilations=[1, 1], group=1),
        N("HardSwish", ["c"], ["h"]),
        N("GlobalAveragePool", ["h"], ["p"]),
        N("Conv", ["p", "w2"], ["c2"], kernel_shape=[1, 1], strides=[1, 1], pads=[0, 0, 0, 0], dilations=[1, 1], group=1),
        N("HardSwish", ["c2"], ["h2"]),
        N("Dropout", ["h2"], ["d"]),
    ]
    inits = [T("w", w), T("b", rng.standard_normal(8).astype(np.float32) * 0.1), T("w2", w2)]
    if tail == "flatten_gemm":
        nodes += [N("Flatten", ["d"], ["f"], axis=1), N("Gemm", ["f", "fc", "fb"], ["l"], alpha=1.0, beta=1.0, transB=1),
                  N("Softmax", ["l"], ["y"], axis=1)]
        inits += [T("fc", fc), T("fb", fb)]
    elif tail == "reshape_matmul_add":
        nodes += [N("Reshape", ["d", "shp"], ["f"]), N("MatMul", ["f", "fc"], ["mm"]), N("Add", ["mm", "fb"], ["l"]),
                  N("Softmax", ["l"], ["y"], axis=-1)]
        inits += [T("shp", np.array([0, -1], np.int64)), T("fc", np.ascontiguousarray(fc.T)), T("fb", fb)]
    else:
        nodes += [N("Squeeze", ["d"], ["f"]), N("MatMul", ["f", "fc"], ["l"]), N("Softmax", ["l"], ["y"], axis=1)]
        inits += [T("fc", np.ascontiguousarray(fc.T))]
    data = onnx_io.model_proto(nodes, inits, [onnx_io.value_info("x", ["N", 3, "H", "W"])],
                               [onnx_io.value_info("y", ["N", 2])])
    x = rng.standard_normal((3, 3, 20, 28)).astype(np.float32)
    want = _run_onnx(data, x)
    blob = onnx_io.import_onnx(data)
    kind, _, ops, _ = onnx_io._parse_oarg(blob)
    assert kind == models.KIND_CLS
    assert [o["type"] for o in ops] == [models.OP_CONV, models.OP_AVGPOOL, models.OP_CONV, models.OP_CTC_HEAD]
    got = OracleNet(blob).forward(x).reshape(3, 2)
    assert np.abs(got - want).max() <= 1e-6 and np.allclose(got.sum(1), 1.0, atol=1e-6)


def test_import_rejects_what_it_cannot_run():
    N = onnx_io.node
    data = onnx_io.model_proto([N("Gelu", ["x"], ["y"])], [], [onnx_io.value_info("x", ["N", 3, 8, 8])],
                               [onnx_io.value_info("y", ["N", 3, 8, 8])])
    with pytest.raises(OCRError) as e:
        onnx_io.import_onnx(data)
    assert "Gelu" in str(e.value) and "supported subset" in str(e.value)


def test_model_source_accepts_onnx_files(tmp_path, built_lib):
    """OAROCRBuilder takes the reference's own model files: an .onnx path (or ONNX bytes) is handed to the C ABI
    unchanged (oar_model_load_onnx converts it inside the library); the library's conversion and the offline tool's
    give the same OARG blob"""
    from oar_ocr_b200 import ffi
    from oar_ocr_b200.ocr import _resolve_model
    blob = models.get_blob("det")
    path = tmp_path / "det.onnx"
    path.write_bytes(onnx_io.export_onnx(blob))
    data = _resolve_model(str(path), "det")
    assert data == path.read_bytes() and _resolve_model(data, "det") == data
    got = ffi.onnx_to_oarg(data)
    assert got[:4] == b"OARG"
    x = np.random.default_rng(0).standard_normal((1, 3, 32, 32)).astype(np.float32)
    assert np.array_equal(OracleNet(got).forward(x), OracleNet(blob).forward(x))
    assert onnx_io.main(["onnx_io", "convert", str(path), str(tmp_path / "det.oarg")]) == 0
    assert (tmp_path / "det.oarg").read_bytes() == got
