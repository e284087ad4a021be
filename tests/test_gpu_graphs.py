"""CUDA graphs over the layer list (engine.cu: model_forward).  The second walk of a (model, lane, shape) is recorded,
every later one replayed: the results of the eager walk, the recording walk and the replays must be the same bits,
equal to the oracle like any other run, while the host-visible submissions drop.  Everything through the C ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-3
GRAPHS_ON = os.environ.get("OAR_GRAPHS", "1") != "0"


@pytest.fixture(scope="module")
def nets(ctx, det_blob, rec_blob):
    from oar_ocr_b200 import ffi
    return ffi.Model(ctx, det_blob), ffi.Model(ctx, rec_blob)


def _flat(results):
    out = []
    for g in results:
        for r in g.text_regions:
            out.append((np.asarray(r.bounding_box.points).copy(), np.asarray(r.label_indices).copy(),
                        float(r.confidence), int(r.detection_index)))
    return out


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])
        assert x[2] == y[2] and x[3] == y[3]  # bit-identical confidences: the same kernels on the same bytes


def test_replayed_graph_equals_eager_walk_and_oracle(nets, det_blob, rec_blob):
    from oracle import pipeline
    from oracle.net import OracleNet
    from oar_ocr_b200 import ffi, synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    # a shape no other test uses, so that run 1 really is this shape's first (eager) walk
    imgs = [synth.page(310 + i, 416) for i in range(3)]
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 3, 24)
    runs, submits, launches = [], [], []
    for _ in range(5):
        s0, l0 = ffi.submit_count(), ffi.launch_count()
        runs.append(_flat(ocr.predict(imgs)))
        submits.append(ffi.submit_count() - s0)
        launches.append(ffi.launch_count() - l0)
    assert len(runs[0]) >= 10
    for r in runs[1:]:
        _same(runs[0], r)
    # every run launches the same kernels; from the third on the networks' share arrives as graphs
    assert len(set(launches)) == 1, launches
    if GRAPHS_ON:
        assert submits[2] == submits[3] == submits[4], submits
        assert submits[2] * 2 < submits[0], submits
    else:
        assert len(set(submits)) == 1 and submits[0] == launches[0]
    # and the replayed result is the oracle's
    want = pipeline.predict(OracleNet(det_blob), OracleNet(rec_blob), imgs, 18385, image_batch_size=3, region_batch_size=24)
    got = ocr.predict(imgs)
    n = 0
    for g, w in zip(got, want):
        assert len(g.text_regions) == len(w)
        for r, o in zip(g.text_regions, w):
            assert np.array_equal(r.bounding_box.points, o["box"])
            assert np.array_equal(r.label_indices, o["labels"])
            assert abs(r.confidence - o["score"]) <= TOL
            n += 1
    assert n == len(runs[0])


def test_graph_inputs_are_read_at_replay_time(nets):
    """same shapes, different pixels: a replay must read the new pages / crops (the input table is re-copied into the
    graph's fixed slot in front of every launch), not the ones it was recorded with"""
    from oar_ocr_b200 import synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 2, 4096)
    a = [synth.page(400 + i, 352) for i in range(2)]
    b = [synth.page(500 + i, 352) for i in range(2)]
    first_a = _flat(ocr.predict(a))          # detector shape seen once
    first_b = _flat(ocr.predict(b))          # detector recorded on b's pages
    assert len(first_a) >= 5 and len(first_b) >= 5
    differs = len(first_a) != len(first_b) or any(not np.array_equal(x[0], y[0]) for x, y in zip(first_a, first_b))
    assert differs, "the two page sets must not produce the same boxes for this test to mean anything"
    for _ in range(2):
        _same(first_a, _flat(ocr.predict(a)))  # replays on a's pages
        _same(first_b, _flat(ocr.predict(b)))


def test_graphs_of_a_destroyed_model_are_dropped(ctx, rec_blob):
    """a graph is named by the model's uid, not its address: a model loaded after another one was destroyed (possibly at
    the same address) starts from its own eager walk"""
    from oar_ocr_b200 import ffi
    rng = np.random.default_rng(5)
    crops = [rng.integers(0, 256, (48, 200, 3), dtype=np.uint8) for _ in range(6)]
    def same(x, y):
        assert x["T"] == y["T"] and np.array_equal(x["scores"], y["scores"])
        assert all(np.array_equal(a, b) for a, b in zip(x["labels"], y["labels"]))

    outs = []
    for _ in range(2):
        m = ffi.Model(ctx, rec_blob)
        res = [m.rec_run(crops, 18385) for _ in range(3)]  # eager walk, recording walk, replay
        same(res[0], res[1])
        same(res[0], res[2])
        outs.append(res[0])
        m.close()
    same(outs[0], outs[1])
