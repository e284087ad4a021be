"""BASELINE.json's configurations as named parity cases (the bench measures configs[1]; the others are test cases).
This file sorts last on purpose: it re-walks paths the earlier files already cover, at the sizes BASELINE.json names.

configs[0]  single 640x640 synthetic image, text detection only, through the predictor API
            (examples/text_detection.rs -> TextDetectionPredictor, predictor default unclip_ratio 1.5)
configs[2]  recogniser only, 48x320 text-line crops              -> tests/test_gpu_fullsize.py::test_rec_512_crops
configs[1/3] det+rec on 960x960 pages, sharded                   -> test_gpu_fullsize.py, test_sharding_cpu.py
configs[4]  PP-DocLayout-L                                       -> host half only, tests/test_layout_post.py"""
import numpy as np
import pytest

LOGIT_TOL = 1e-3


def _config0_page():
    from oar_ocr_b200 import synth
    return synth.page(1, 640)  # "C1 uses the same generator at 640x640, seed 1"


def test_config0_oracle_is_well_conditioned(det_blob):
    """the CPU side of configs[0]: the oracle finds the page's text lines, none near the box threshold"""
    from oracle import pipeline
    from oracle.net import OracleNet
    (boxes, scores), = pipeline.det_forward(OracleNet(det_blob), [_config0_page()], unclip_ratio=1.5)
    assert len(boxes) >= 10 and scores.min() > 0.7  # box_thresh is 0.6


@pytest.mark.gpu
def test_config0_single_640_detection(ctx, det_blob):
    """configs[0] on the GPU through TextDetectionPredictor: boxes identical to the oracle, scores within 1e-3"""
    from oracle import pipeline
    from oracle.net import OracleNet
    from oar_ocr_b200 import ffi
    from oar_ocr_b200.ocr import TextDetectionConfig, TextDetectionPredictor
    pred = TextDetectionPredictor(ffi.Model(ctx, det_blob), TextDetectionConfig())  # predictor defaults: unclip 1.5
    got = pred.predict([_config0_page()]).detections[0]
    (boxes, scores), = pipeline.det_forward(OracleNet(det_blob), [_config0_page()], unclip_ratio=1.5)
    assert len(got) == len(boxes) >= 10
    for d, b, s in zip(got, boxes, scores):
        assert np.array_equal(d.bbox.points, b)
        assert abs(d.score - float(s)) <= LOGIT_TOL


@pytest.mark.gpu
def test_engine_refuses_unknown_model_kinds(ctx):
    """a blob whose header names a model kind the library does not know is refused outright (checked in the blob
    header before anything touches the device) instead of falling back to anything"""
    import struct
    from oar_ocr_b200 import ffi, models
    blob = bytearray(models.build_hgnetv2_l(return_idx=(0,)))
    struct.pack_into("<I", blob, 8, 9)
    with pytest.raises(ffi.OCRError) as e:
        ffi.Model(ctx, bytes(blob))
    assert e.value.code == ffi.OAR_E_MODEL and "unknown model kind" in str(e.value)
