"""HGNetV2-L, the backbone of PP-DocLayout-L and of the server-size OCR models (item 4):
graph spec + oracle only so far.  Checks the layer list against the reference's in-tree description
(oar-ocr-vl/src/models/pp_doclayout/hgnetv2.rs:13-23, 161-348) and the oracle's execution of the two new ops."""
import numpy as np
import pytest


def test_stage_outputs_follow_the_reference_table():
    """strides 4 / 8 / 16 / 32 with 128 / 512 / 1024 / 2048 channels (STAGE_OUT_CHANNELS, STAGE_DOWNSAMPLE)"""
    from oar_ocr_b200 import models
    from oracle.net import OracleNet
    x = np.random.default_rng(0).standard_normal((1, 3, 96, 128)).astype(np.float32)
    for idx, (c, s) in enumerate([(128, 4), (512, 8), (1024, 16), (2048, 32)]):
        y = OracleNet(models.build_hgnetv2_l(return_idx=(idx,))).forward(x)
        assert y.shape == (1, c, 96 // s, 128 // s)
        assert np.isfinite(y).all() and y.min() >= 0.0 and y.max() > 0.0  # every block ends in ReLU (+ residual of ReLUs)


def test_layer_list_matches_the_reference_description():
    from oar_ocr_b200 import models
    from oracle.net import parse
    kind, _, ops, _ = parse(models.build_hgnetv2_l())
    assert kind == models.KIND_FEAT
    t = [o["type"] for o in ops]
    # stem: 5 convs, two right/bottom pads, one 2x2 stride-1 max pool (Embeddings::forward, hgnetv2.rs:333-348)
    assert t.count(models.OP_PAD) == 2 and t.count(models.OP_MAXPOOL) == 1
    # 3 depthwise stride-2 downsamples + (3 + 1 blocks) x 6 light layers' depthwise convs
    dw = [o for o in ops if o["type"] == models.OP_DWCONV]
    assert sum(1 for o in dw if tuple(o["p"][2:4]) == (2, 2)) == 3
    assert sum(1 for o in dw if tuple(o["p"][2:4]) == (1, 1)) == 4 * 6
    assert all(o["p"][0] == 5 for o in dw if tuple(o["p"][2:4]) == (1, 1))  # STAGE_KERNEL_SIZE 5 in the light stages
    # residual adds: blocks 2 and 3 of stage 3 only (STAGE_NUM_BLOCKS = [1, 1, 3, 1], residual = i != 0)
    assert t.count(models.OP_ADD) == 2
    # aggregation convs: total -> out/2 -> out with total = in + 6 * mid
    agg = [o for o in ops if o["type"] == models.OP_CONV and o["p"][0] == 1 and o["p"][6] in (48 + 6 * 48, 128 + 6 * 96,
                                                                                          512 + 6 * 192, 1024 + 6 * 192,
                                                                                          1024 + 6 * 384)]
    assert sorted(o["p"][7] for o in agg) == [64, 256, 512, 512, 512, 1024]


def test_server_recogniser_graph():
    """PP-OCRv5_server_rec shaped graph on the oracle: same sequence geometry as the mobile recogniser (T = W / 8), a
    softmax over the vocabulary per timestep, ~8x the mobile model's arithmetic (roofline/layers.json)"""
    import json
    import os
    from oar_ocr_b200 import models
    from oracle.net import OracleNet
    x = np.random.default_rng(1).standard_normal((2, 3, 48, 160)).astype(np.float32)
    y = OracleNet(models.build_rec_server(vocab=97)).forward(x)
    assert y.shape == (2, 20, 97) and np.allclose(y.sum(-1), 1.0, atol=1e-5)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "roofline", "layers.json")))
    assert d["rec_server"]["gflop_per_item"] > 5 * d["rec"]["gflop_per_item"]
    # wide dense convs: the 3-pass tensor-core time exceeds the HBM time of the per-layer traffic (tensor-bound)
    assert d["rec_server"]["tensor_ms_at_1590_tflops_x3"] > 0.5 * d["rec_server"]["hbm_ms_at_6650_gbs_fused"]


def test_new_ops_in_the_oracle():
    """OP_PAD = zero padding (top, left, bottom, right); OP_MAXPOOL = floor-mode max pool without padding"""
    from oar_ocr_b200 import models
    from oracle.net import OracleNet
    g = models.GraphBuilder(models.KIND_FEAT, 0)
    g.channels[0] = 3
    g.maxpool(g.pad(0, 0, 0, 1, 1), (2, 2), (1, 1))
    x = np.arange(2 * 3 * 3 * 4, dtype=np.float32).reshape(2, 3, 3, 4) - 30.0
    y = OracleNet(g.serialize()).forward(x)
    p = np.pad(x, ((0, 0), (0, 0), (0, 1), (0, 1)))
    want = np.maximum.reduce([p[:, :, :-1, :-1], p[:, :, 1:, :-1], p[:, :, :-1, 1:], p[:, :, 1:, 1:]])
    assert y.shape == x.shape and np.array_equal(y, want)
