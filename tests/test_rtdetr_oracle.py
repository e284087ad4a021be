"""The oracle of the layout network (oracle-first): RT-DETR-L on torch-CPU (oracle/rtdetr.py).
No CUDA counterpart yet and no reference outputs to pin it (PARITY UNPINNED, see the module header); these tests check
the restated pieces against the reference's formulas / independent formulations and run the network end to end into the
layout post-process that IS built (oar_layout_postprocess)."""
import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def net():
    from oracle.rtdetr import RTDetrL
    return RTDetrL(seed=42)


def test_anchors_follow_the_reference():
    """model.rs:154-183: 0.05 * 2^level boxes on cell centres, logit space, invalid where a coordinate leaves
    (0.01, 0.99); 80^2 + 40^2 + 20^2 = 8400 tokens at 640 x 640"""
    from oracle.rtdetr import generate_anchors
    a, v = generate_anchors([(80, 80), (40, 40), (20, 20)])
    assert a.shape == (1, 8400, 4) and v.shape == (1, 8400, 1)
    c = np.array([0.5 / 80, 0.5 / 80, 0.05, 0.05], np.float32)  # first cell: centre 0.00625 < 0.01 -> invalid
    assert v[0, 0, 0] == 0 and a[0, 0, 0] == np.finfo(np.float32).max
    c = np.array([1.5 / 80, 1.5 / 80, 0.05, 0.05], np.float32)
    assert v[0, 81, 0] == 1 and np.allclose(a[0, 81].numpy(), np.log(c / (1 - c)), rtol=1e-6)
    assert np.allclose(torch.sigmoid(a[0, 6400 + 41, 2:]).numpy(), 0.1, atol=1e-6)  # level 1: 0.05 * 2
    assert int(v.sum()) == 78 * 78 + 40 * 40 + 20 * 20  # only the outer ring of the finest level is invalid


def test_sine_position_embedding_layout():
    """encoder.rs:179-216: [sin(y w), cos(y w), sin(x w), cos(x w)] with w_k = 10000^(-k / (dim/4))"""
    from oracle.rtdetr import sine_position_embedding
    p = sine_position_embedding(3, 5, 256)[0].numpy()
    assert p.shape == (15, 256)
    y, x, k = 2, 3, 7
    om = 1.0 / 10000.0 ** (k / 64)
    row = p[y * 5 + x]
    assert np.allclose([row[k], row[64 + k], row[128 + k], row[192 + k]],
                       [np.sin(y * om), np.cos(y * om), np.sin(x * om), np.cos(x * om)], atol=1e-6)


def test_deformable_attention_equals_a_naive_gather():
    """decoder.rs:337-470 spelled out: pixel = loc * size - 0.5, four corners, a corner counts only inside the map"""
    from oracle.rtdetr import POINTS, deformable_attention
    rng = np.random.default_rng(0)
    B, Q, heads, hd = 2, 5, 2, 3
    shapes = [(4, 6), (2, 3), (1, 2)]
    N = sum(h * w for h, w in shapes)
    value = torch.from_numpy(rng.standard_normal((B, N, heads, hd)).astype(np.float32))
    loc = torch.from_numpy(rng.uniform(-0.2, 1.2, (B, Q, heads, 3, POINTS, 2)).astype(np.float32))
    wts = torch.softmax(torch.from_numpy(rng.standard_normal((B, Q, heads, 3 * POINTS)).astype(np.float32)), -1)
    wts = wts.reshape(B, Q, heads, 3, POINTS)
    got = deformable_attention(value, shapes, loc, wts).numpy().reshape(B, Q, heads, hd)
    want = np.zeros((B, Q, heads, hd), np.float64)
    off = 0
    for lvl, (h, w) in enumerate(shapes):
        for b in range(B):
            for q in range(Q):
                for hh in range(heads):
                    for p in range(POINTS):
                        px = float(loc[b, q, hh, lvl, p, 0]) * w - 0.5
                        py = float(loc[b, q, hh, lvl, p, 1]) * h - 0.5
                        x0, y0 = int(np.floor(px)), int(np.floor(py))
                        for dx in (0, 1):
                            for dy in (0, 1):
                                xi, yi = x0 + dx, y0 + dy
                                if 0 <= xi < w and 0 <= yi < h:
                                    cw = (px - x0 if dx else 1 - (px - x0)) * (py - y0 if dy else 1 - (py - y0))
                                    want[b, q, hh] += cw * float(wts[b, q, hh, lvl, p]) * \
                                        value[b, off + yi * w + xi, hh].numpy()
        off += h * w
    assert np.abs(got - want).max() < 1e-5


def test_forward_shapes_and_batch_independence(net):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 3, 256, 320)).astype(np.float32)
    logits, boxes = net.forward(x)
    assert logits.shape == (2, 300, 23) and boxes.shape == (2, 300, 4)
    assert np.isfinite(logits).all() and boxes.min() > 0.0 and boxes.max() < 1.0
    l0, b0 = net.forward(x[:1])
    assert np.abs(l0[0] - logits[0]).max() < 1e-4 and np.abs(b0[0] - boxes[0]).max() < 1e-5  # no cross-image coupling
    l1, b1 = net.forward(x)
    assert np.array_equal(l1, logits) and np.array_equal(b1, boxes)  # deterministic


def test_network_rows_through_the_layout_postprocess(net, built_lib):
    """end to end on the CPU tier: oracle network -> exported-model rows -> oar_layout_postprocess (product, host code)
    == the oracle restatement of the adapter's post-process, bit for bit"""
    from oracle import cpu
    from oar_ocr_b200 import ffi
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 3, 320, 320)).astype(np.float32)
    sizes = [(1000.0, 800.0), (640.0, 640.0)]
    rows = net.rows(x, sizes)
    assert rows.shape == (2, 300, 6)
    assert (np.diff(rows[:, :, 1], axis=1) <= 0).all()  # top-300 comes out in score order
    assert rows[:, :, 0].min() >= 0 and rows[:, :, 0].max() < 23
    kw = dict(score_threshold=0.5, image_class_id=1, formula_class_id=7, class_merge_modes={0: 0, 2: 2, 7: 0})
    got = ffi.layout_postprocess(rows, sizes, 23, **kw)
    for b, (w, h) in enumerate(sizes):
        ob, oc, os_ = cpu.layout_postprocess(rows[b], w, h, 23, **kw)
        gb, gc, gs = got[b]
        assert gc.tolist() == oc.tolist() and np.array_equal(gb, ob) and np.array_equal(gs, os_)
        assert (gb[:, 0] >= 0).all() and (gb[:, 2] <= w).all() and (gb[:, 3] <= h).all() and len(gc) <= 100
    assert sum(len(g[1]) for g in got) > 0


def test_catmullrom_and_lanczos_resize_cross_checks():
    """image::imageops::resize with the filters of the layout detectors (scale_aware_detector.rs:52-80), restated from
    the crate's published sampler (unpinned: the crate is not vendored).  Cross-checks: CatmullRom is torch's
    antialiased bicubic (Keys a = -0.5) to within 2 grey levels on a smooth image, both filters keep a constant
    image constant (normalised weights) and a same-size resize is the identity"""
    import torch.nn.functional as F
    from oracle import cpu
    yy, xx = np.mgrid[0:90, 0:130].astype(np.float32)
    img = np.stack([127 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 40 + xx, 250 - 2 * yy], -1).clip(0, 255).astype(np.uint8)
    for (nw, nh) in ((64, 48), (200, 150)):
        got = cpu.resize_filter(img, nw, nh, "catmullrom").astype(np.float32)
        t = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)[None]
        want = F.interpolate(t, size=(nh, nw), mode="bicubic", antialias=True, align_corners=False)[0].permute(1, 2, 0)
        assert np.abs(got - want.clamp(0, 255).numpy()).max() <= 2.0
    flat = np.full((37, 53, 3), 93, np.uint8)
    for f in ("triangle", "catmullrom", "lanczos3"):
        assert (cpu.resize_filter(flat, 80, 21, f) == 93).all()
        assert np.array_equal(cpu.resize_filter(img, 130, 90, f), img)
    # Lanczos3 overshoots at a step edge (negative lobes), Triangle cannot
    step = np.zeros((8, 64, 3), np.uint8)
    step[:, 32:] = 200
    assert cpu.resize_filter(step, 256, 8, "lanczos3").max() > 200 and cpu.resize_filter(step, 256, 8, "triangle").max() == 200


def test_oracle_chain_for_config4(net, built_lib):
    """BASELINE.json configs[4] end to end on the CPU oracle at a reduced size: pages -> layout preprocess (CatmullRom to
    a fixed square, 1/255, RGB) -> RT-DETR-L -> rows -> post-process; the product's host half agrees with the oracle's"""
    from oracle import cpu
    from oar_ocr_b200 import ffi, synth
    pages = [synth.page(5, 480), synth.page(6, 320)[:200]]
    x, scale = cpu.layout_preprocess(pages, (256, 256))
    assert x.shape == (2, 3, 256, 256) and 0.0 <= x.min() and x.max() <= 1.0
    assert np.allclose(scale, [[256 / 480, 256 / 480], [256 / 200, 256 / 320]])
    sizes = [(float(p.shape[1]), float(p.shape[0])) for p in pages]
    rows = net.rows(x, sizes)
    got = ffi.layout_postprocess(rows, sizes, 23, score_threshold=0.3, image_class_id=1, formula_class_id=7)
    for b, (w, h) in enumerate(sizes):
        ob, oc, os_ = cpu.layout_postprocess(rows[b], w, h, 23, score_threshold=0.3, image_class_id=1, formula_class_id=7)
        assert got[b][1].tolist() == oc.tolist() and np.array_equal(got[b][0], ob) and np.array_equal(got[b][2], os_)


def test_encoder_as_oarg_layer_list_equals_the_oracle(net):
    """models.build_layout_encoder states backbone + hybrid encoder + decoder-input projections as ONE OARG graph
    (the spec the CUDA executor gets next round; new pieces: exact GELU, attention with 2-D positions on q / k, the
    token-rows op, RepVGG branches merged into one 3x3).  Executed by oracle/net.py it reproduces oracle/rtdetr.py's
    decoder memory."""
    from oar_ocr_b200 import models
    from oracle.net import OracleNet, parse
    x = np.random.default_rng(3).standard_normal((2, 3, 192, 256)).astype(np.float32)
    want, shapes = net.encode(x)
    assert shapes == [(24, 32), (12, 16), (6, 8)]
    blob = models.build_layout_encoder(models.layout_weights(42), seed=42, shapes_hw=shapes)
    kind, _, ops, _ = parse(blob)
    types = [o["type"] for o in ops]
    assert kind == models.KIND_FEAT and types.count(models.OP_TOKENS) == 3 and types.count(models.OP_ATTN) == 1
    got = OracleNet(blob).forward(x)  # [B, 256, 1, N]
    assert got.shape == (2, 256, 1, want.shape[1])
    got = got[:, :, 0].transpose(0, 2, 1)
    scale = np.abs(want.numpy()).max()
    assert np.abs(got - want.numpy()).max() <= 2e-5 * max(scale, 1.0)


def test_spec_only_ops_do_not_export_silently(net):
    """the ONNX exporter refuses what it cannot express yet instead of dropping it (positions, GELU, token rows)"""
    from oar_ocr_b200 import models, onnx_io
    from oar_ocr_b200.ffi import OCRError
    blob = models.build_layout_encoder(models.layout_weights(42), seed=42, shapes_hw=[(8, 8), (4, 4), (2, 2)])
    with pytest.raises(OCRError):
        onnx_io.export_onnx(blob)
