"""ONNX behind the C ABI (oar_model_load_onnx; the reference's boundary is
OrtInfer::new(ModelSource::{Path, Memory}), oar-ocr-core/src/core/inference/ort_infer_builders.rs:9-70,
core/config/model_source.rs:20-28).  CPU tier: the library's own conversion (csrc/onnx_import.cu, reached through
oar_onnx_to_oarg, no device needed) gives byte-identical OARG blobs to the offline Python importer for every graph
the exporter can write and for the alternative operator spellings; malformed input fails with OCRError("ModelLoad").
GPU tier: a model loaded from ONNX bytes through oar_model_load_onnx runs and equals the OARG-loaded one."""
import struct

import numpy as np
import pytest

from oar_ocr_b200 import ffi, models, onnx_io
from oar_ocr_b200.ffi import OCRError


@pytest.mark.parametrize("kind", ["det", "rec", "cls"])
def test_cabi_import_equals_python_import(built_lib, kind):
    blob = models.get_blob(kind)
    data = onnx_io.export_onnx(blob)
    hint = models.KIND_CLS if kind == "cls" else None
    want = onnx_io.import_onnx(data, hint)
    got = ffi.onnx_to_oarg(data, models.KIND_CLS if kind == "cls" else -1)
    assert got == want
    ffi.validate_blob(got)
    # and the converted graph computes what the original layer list computes
    from oracle.net import OracleNet
    shape = {"det": (1, 3, 64, 64), "rec": (2, 3, 48, 64), "cls": (2, 3, 80, 160)}[kind]
    x = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    assert np.array_equal(OracleNet(got).forward(x), OracleNet(blob).forward(x))


def test_cabi_import_hgnetv2_stem_ops(built_lib):
    """Pad / MaxPool travel through the conversion (the engine refuses to LOAD them until their kernels exist)"""
    blob = models.get_blob("hgnetv2") if "hgnetv2" in getattr(models, "BLOB_KINDS", ()) else None
    if blob is None:
        g = models.GraphBuilder(models.KIND_DET, 3)
        x = g.conv(0, 16, (3, 3), (2, 2), act=models.ACT_RELU)
        x = g.pad(x, 0, 0, 1, 1)
        x = g.maxpool(x, (2, 2), (1, 1))
        x = g.conv(x, 1, (1, 1), act=models.ACT_SIGMOID)
        blob = g.serialize()
    data = onnx_io.export_onnx(blob)
    assert ffi.onnx_to_oarg(data) == onnx_io.import_onnx(data)


def _alt_spellings():
    """BatchNormalization after Conv, decomposed hardswish and swish, a ConvTranspose with BN, SE without biases,
    auto_pad, half_pixel nearest Resize: what another exporter may write for the same network"""
    N, T = onnx_io.node, onnx_io.tensor_proto
    rng = np.random.default_rng(5)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)  # noqa: E731
    inits = [T("w1", f(8, 3, 3, 3)), T("bn_s", 1 + 0.1 * f(8)), T("bn_b", f(8)), T("bn_m", f(8)), T("bn_v", np.abs(f(8)) + 0.5),
             T("w2", f(8, 1, 3, 3)), T("b2", f(8)), T("sw1", f(4, 8, 1, 1)), T("sw2", f(8, 4, 1, 1)),
             T("w3", f(8, 4, 2, 2)), T("b3", f(4)), T("bn2_s", 1 + 0.1 * f(4)), T("bn2_b", f(4)), T("bn2_m", f(4)),
             T("bn2_v", np.abs(f(4)) + 0.5), T("scales", np.array([1, 1, 2, 2], np.float32)), T("w4", f(1, 4, 1, 1)),
             T("b4", f(1))]
    nodes = [
        N("Conv", ["x", "w1"], ["c1"], kernel_shape=[3, 3], strides=[1, 1], auto_pad="SAME_UPPER"),
        N("BatchNormalization", ["c1", "bn_s", "bn_b", "bn_m", "bn_v"], ["n1"], epsilon=1e-3),
        N("HardSigmoid", ["n1"], ["g1"], alpha=1.0 / 6.0, beta=0.5),
        N("Mul", ["n1", "g1"], ["a1"]),                                     # decomposed hardswish
        N("Conv", ["a1", "w2", "b2"], ["c2"], kernel_shape=[3, 3], strides=[2, 2], pads=[1, 1, 1, 1], group=8),
        N("Sigmoid", ["c2"], ["g2"]),
        N("Mul", ["c2", "g2"], ["a2"]),                                     # decomposed swish
        N("GlobalAveragePool", ["a2"], ["p"]),
        N("Conv", ["p", "sw1"], ["s1"], kernel_shape=[1, 1]),               # squeeze-excite convs without biases
        N("Relu", ["s1"], ["s2"]),
        N("Conv", ["s2", "sw2"], ["s3"], kernel_shape=[1, 1]),
        N("HardSigmoid", ["s3"], ["s4"], alpha=0.2, beta=0.5),
        N("Mul", ["a2", "s4"], ["se"]),
        N("ConvTranspose", ["se", "w3", "b3"], ["d1"], kernel_shape=[2, 2], strides=[2, 2]),
        N("BatchNormalization", ["d1", "bn2_s", "bn2_b", "bn2_m", "bn2_v"], ["n2"]),
        N("Relu", ["n2"], ["r2"]),
        N("Resize", ["r2", "", "scales"], ["u"], mode="nearest"),           # ONNX defaults: half_pixel / round_prefer_floor
        N("Conv", ["u", "w4", "b4"], ["c4"], kernel_shape=[1, 1], auto_pad="VALID"),
        N("Sigmoid", ["c4"], ["y"]),
    ]
    return onnx_io.model_proto(nodes, inits, [onnx_io.value_info("x", ["N", 3, "H", "W"])],
                               [onnx_io.value_info("y", ["N", 1, "H", "W"])])


def test_cabi_import_alternative_spellings(built_lib):
    data = _alt_spellings()
    want = onnx_io.import_onnx(data)
    got = ffi.onnx_to_oarg(data)
    assert got == want
    _, _, ops, _ = onnx_io._parse_oarg(got)
    assert [o["type"] for o in ops] == [models.OP_CONV, models.OP_DWCONV, models.OP_SE, models.OP_DECONV2,
                                        models.OP_UPSAMPLE, models.OP_CONV]
    assert ops[0]["p"][4:6] == [1, 1] and ops[0]["p"][8] == models.ACT_HSWISH and ops[1]["p"][7] == models.ACT_SWISH
    # the converted graph computes what the ONNX operators define (written out with torch ops, fp32)
    import torch
    import torch.nn.functional as F
    from oracle.net import OracleNet
    _, inits, _, _ = onnx_io.read_model(data)
    t = {k: torch.from_numpy(np.array(v)) for k, v in inits.items()}
    x = np.random.default_rng(1).standard_normal((2, 3, 16, 12)).astype(np.float32)
    v = F.conv2d(torch.from_numpy(x), t["w1"], None, 1, 1)
    v = F.batch_norm(v, t["bn_m"], t["bn_v"], t["bn_s"], t["bn_b"], False, 0.0, 1e-3)
    v = v * torch.clamp(v / 6.0 + 0.5, 0.0, 1.0)
    v = F.conv2d(v, t["w2"], t["b2"], 2, 1, 1, 8)
    v = v * torch.sigmoid(v)
    g = v.mean(dim=(2, 3), keepdim=True)
    g = torch.clamp(F.conv2d(F.relu(F.conv2d(g, t["sw1"])), t["sw2"]) * 0.2 + 0.5, 0.0, 1.0)
    v = v * g
    v = F.conv_transpose2d(v, t["w3"], t["b3"], 2)
    v = F.relu(F.batch_norm(v, t["bn2_m"], t["bn2_v"], t["bn2_s"], t["bn2_b"], False, 0.0, 1e-5))
    v = F.interpolate(v, scale_factor=2.0, mode="nearest")
    want_y = torch.sigmoid(F.conv2d(v, t["w4"], t["b4"])).numpy()
    assert np.abs(OracleNet(got).forward(x) - want_y).max() <= 2e-5


def test_cabi_import_rejects_what_it_cannot_run(built_lib):
    N = onnx_io.node
    vi = onnx_io.value_info
    cases = {
        "supported subset": onnx_io.model_proto([N("Gelu", ["x"], ["y"])], [], [vi("x", ["N", 3, 8, 8])], [vi("y", ["N", 3, 8, 8])]),
        "auto_pad": onnx_io.model_proto(
            [N("Conv", ["x", "w"], ["y"], kernel_shape=[3, 3], strides=[2, 2], auto_pad="SAME_UPPER")],
            [onnx_io.tensor_proto("w", np.zeros((4, 3, 3, 3), np.float32))], [vi("x", ["N", 3, 8, 8])], [vi("y", ["N", 4, 4, 4])]),
        "pixel replication": onnx_io.model_proto(
            [N("Resize", ["x", "", "s"], ["y"], mode="nearest", coordinate_transformation_mode="align_corners")],
            [onnx_io.tensor_proto("s", np.array([1, 1, 2, 2], np.float32))], [vi("x", ["N", 3, 8, 8])], [vi("y", ["N", 3, 16, 16])]),
        "initializer": onnx_io.model_proto(
            [N("Add", ["x", "c"], ["y"])], [onnx_io.tensor_proto("c", np.zeros((1, 3, 1, 1), np.float32))],
            [vi("x", ["N", 3, 8, 8])], [vi("y", ["N", 3, 8, 8])]),
    }
    for needle, data in cases.items():
        for conv in (lambda d: ffi.onnx_to_oarg(d), onnx_io.import_onnx):
            with pytest.raises(OCRError) as e:
                conv(data)
            assert e.value.kind == "ModelLoad" and needle in str(e.value), (needle, str(e.value))
    for junk in (b"", b"\x00\x01\x02", b"\x3a\xff\xff\xff\xff\x0f", bytes(range(256)) * 4):
        with pytest.raises(OCRError) as e:
            ffi.onnx_to_oarg(junk)
        assert e.value.kind == "ModelLoad"


def test_validate_blob_rejects_corrupt_oarg(built_lib, det_blob):
    """oar_model_load_blob's structural checks, reachable without a device: truncated blobs, a 64-bit weight count that
    would wrap the size computation, out-of-range tensor ids / weight slices, a weight slice shorter than its
    convolution needs, zero strides"""
    ffi.validate_blob(det_blob)
    head = 28

    def patched(off, fmt, *vals):
        b = bytearray(det_blob)
        struct.pack_into(fmt, b, off, *vals)
        return bytes(b)

    bad = {
        "truncated": det_blob[: len(det_blob) // 2],
        "short header": det_blob[:20],
        "magic": b"XXXX" + det_blob[4:],
        "wrap n_w": patched(20, "<Q", (1 << 62) + 5),
        "n_ops": patched(12, "<I", 0xFFFFFFFF),
        "in0": patched(head + 4, "<i", 1 << 20),
        "in1": patched(head + 8, "<i", -7),
        "stride 0": patched(head + 16 + 2 * 4, "<i", 0),
        "w_off": patched(head + 16 + 48 + 16, "<q", (1 << 62)),
        "w_len short": patched(head + 16 + 48 + 16 + 32, "<q", 8),
        "w_off negative": patched(head + 16 + 48 + 16, "<q", -4),
    }
    for name, blob in bad.items():
        with pytest.raises(OCRError) as e:
            ffi.validate_blob(blob)
        assert e.value.kind == "ModelLoad", name


@pytest.mark.gpu
def test_load_onnx_through_the_c_abi_runs_like_oarg(ctx, det_blob, rec_blob):
    """oar_model_load_onnx: a model loaded from ONNX bytes gives the same outputs as the OARG-loaded one (bit for bit:
    the conversion reproduces the blob), and a role mismatch is refused"""
    from oar_ocr_b200 import synth
    det_onnx, rec_onnx = onnx_io.export_onnx(det_blob), onnx_io.export_onnx(rec_blob)
    d0, d1 = ffi.Model(ctx, det_blob), ffi.Model(ctx, det_onnx, ffi.KIND_DET)
    pages = [synth.page(3, 320), synth.page(4, 320)]
    for (b0, s0), (b1, s1) in zip(d0.det_run(pages), d1.det_run(pages)):
        assert np.array_equal(b0, b1) and np.array_equal(s0, s1) and len(b0) > 0
    r0, r1 = ffi.Model(ctx, rec_blob), ffi.Model(ctx, rec_onnx, ffi.KIND_REC)
    crops = [synth.crop(j, 48, 160 + 16 * j) for j in range(6)]
    a, b = r0.rec_run(crops, 18385), r1.rec_run(crops, 18385)
    assert all(np.array_equal(x, y) for x, y in zip(a["labels"], b["labels"])) and np.array_equal(a["scores"], b["scores"])
    with pytest.raises(OCRError) as e:
        ffi.Model(ctx, det_onnx, ffi.KIND_REC)
    assert e.value.kind == "ModelLoad" and "detection" in str(e.value)
