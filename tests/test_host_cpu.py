"""CPU-side checks: the C-ABI library builds, loads and exports every declared symbol; host logic that needs no
device (config defaults, sort_quad_boxes, argument validation, builder validation) behaves like the reference."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    from oar_ocr_b200 import ffi
    header = open(os.path.join(ROOT, "include", "oar_b200.h")).read()
    declared = set(re.findall(r"\b(oar_[a-z0-9_]+)\s*\(", header))
    assert declared == set(ffi.SYMBOLS)
    lib = C.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_device_fails_loudly(built_lib):
    """there is no CPU fallback: without an sm_100 device every compute path is unreachable"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oar_ocr_b200 import ffi
    with pytest.raises(ffi.OCRError) as e:
        ffi.Context(0)
    assert e.value.code == ffi.OAR_E_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_config_defaults(built_lib):
    from oar_ocr_b200 import ffi
    d = ffi.det_config()
    # OAROCRBuilder defaults, src/oarocr/ocr.rs:351-364
    assert (round(d.thresh, 6), round(d.box_thresh, 6), d.unclip_ratio, d.max_candidates) == (0.3, 0.6, 2.0, 1000)
    assert (d.limit_side_len, d.limit_type, d.max_side_limit, d.min_size) == (960, 0, 4000, 3.0)
    p = ffi.pipeline_config()
    assert (p.image_batch_size, p.region_batch_size, p.n_chars) == (8, 64, 18385)  # accelerator defaults


def test_sort_quad_boxes_reference_cases(built_lib):
    """sorting.rs:740-808 through the C ABI (host-only entry point)"""
    from oar_ocr_b200 import ffi

    def fc(x1, y1, x2, y2):
        return [[x1, y1], [x2, y1], [x2, y2], [x1, y2]]

    b, _ = ffi.sort_quad_boxes(np.array([fc(10, 50, 50, 70), fc(10, 10, 50, 30), fc(10, 30, 50, 50)], np.float32))
    assert [float(x[:, 1].min()) for x in b] == [10.0, 30.0, 50.0]
    b, o = ffi.sort_quad_boxes(np.array([fc(60, 10, 100, 30), fc(10, 11, 50, 31), fc(10, 50, 50, 70),
                                         fc(60, 52, 100, 72)], np.float32))
    assert o.tolist() == [1, 0, 2, 3]
    b, o = ffi.sort_quad_boxes(np.zeros((0, 4, 2), np.float32))
    assert len(b) == 0


def test_sort_quad_boxes_matches_oracle(built_lib):
    from oar_ocr_b200 import ffi
    from oracle import cpu
    rng = np.random.default_rng(0)
    for n in (1, 5, 64, 300):
        tl = np.stack([rng.integers(0, 900, n), rng.integers(0, 40, n) * 9], -1).astype(np.float32)
        boxes = tl[:, None, :] + np.array([[0, 0], [80, 0], [80, 20], [0, 20]], np.float32)[None]
        gb, go = ffi.sort_quad_boxes(boxes)
        wb, wo = cpu.sort_quad_boxes(boxes)
        assert np.array_equal(go, wo) and np.array_equal(gb, wb)


def test_builder_batch_size_validation():
    """ocr.rs:1168-1195"""
    from oar_ocr_b200.ocr import MAX_BATCH_SIZE, OAROCRBuilder, OCRError
    OAROCRBuilder.validate_batch_size("image_batch_size", 1)
    OAROCRBuilder.validate_batch_size("region_batch_size", MAX_BATCH_SIZE)
    for name, v in (("image_batch_size", 0), ("region_batch_size", MAX_BATCH_SIZE + 1)):
        with pytest.raises(OCRError) as e:
            OAROCRBuilder.validate_batch_size(name, v)
        assert name in str(e.value) and f"1..={MAX_BATCH_SIZE}" in str(e.value)


def test_character_list_layout():
    """decode.rs:392-423: blank first, dictionary, trailing space"""
    from oar_ocr_b200.ocr import character_list
    from oar_ocr_b200.models import synthetic_dict
    chars = character_list(synthetic_dict())
    assert len(chars) == 18385 and chars[0] == "\0" and chars[-1] == " "


def test_dictionary_lines_follow_rust_semantics(tmp_path):
    """Rust `str::lines()` + `filter_map(|s| s.chars().next())` (ocr.rs:386, decode.rs:118-121): only '\\n' / '\\r\\n'
    end a line, empty lines are dropped, a multi-character line contributes its first char.  U+2028, U+0085, \\x0b,
    \\x0c and \\x1c are ordinary characters (Python's splitlines() would split on each and shift the class indices)."""
    from oar_ocr_b200.ocr import TextRecognitionPredictorBuilder, character_list, dict_lines
    content = "a\n\nbc\r\n\u2028\n\x0bq\n\x85\n\x1c\n\x0c\nz"
    lines = dict_lines(content)
    assert lines == ["a", "", "bc", "\u2028", "\x0bq", "\x85", "\x1c", "\x0c", "z"]
    chars = character_list(lines)
    assert chars == ["\0", "a", "b", "\u2028", "\x0b", "\x85", "\x1c", "\x0c", "z", " "]
    assert dict_lines("") == [] and dict_lines("x\n") == ["x"] and dict_lines("x\n\n") == ["x", ""]
    assert dict_lines("x\r") == ["x\r"] and dict_lines("x\r\n") == ["x"]  # a bare CR is not a terminator
    # the file-reading path must not translate newlines either (newline="")
    f = tmp_path / "dict.txt"
    f.write_bytes(content.encode("utf-8"))
    b = TextRecognitionPredictorBuilder().dict_path(str(f))
    with open(b._dict, "r", encoding="utf-8", newline="") as fh:
        assert character_list(dict_lines(fh.read())) == chars


def test_recognition_builder_requires_dict():
    from oar_ocr_b200.ocr import OCRError, TextRecognitionPredictor
    with pytest.raises(OCRError) as e:
        TextRecognitionPredictor.builder().build("synthetic")
    assert "dict_path" in str(e.value)


def test_model_blobs_are_well_formed(det_blob, rec_blob):
    from oracle.net import parse
    for blob, kind in ((det_blob, 0), (rec_blob, 1)):
        k, n_tensors, ops, weights = parse(blob)
        assert k == kind and len(ops) > 40
        assert struct.unpack_from("<I", blob, 4)[0] == 1
        for op in ops:
            assert 0 <= op["in0"] < n_tensors and 0 <= op["out"] < n_tensors
            for o, n in zip(op["w_off"], op["w_len"]):
                assert 0 <= o and o + n <= len(weights)
    # CTC head: V = 18385 = dict + blank + space
    assert parse(rec_blob)[2][-1]["p"][1] == 18385


def test_oracle_pipeline_smoke(det_blob, rec_blob):
    """the oracle's own end-to-end composition yields boxes and label sequences on a synthetic page"""
    from oar_ocr_b200 import synth
    from oracle import pipeline
    from oracle.net import OracleNet
    res = pipeline.predict(OracleNet(det_blob), OracleNet(rec_blob), [synth.page(9, 480)], 18385,
                           region_batch_size=4)[0]
    g = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_v1.npz"))
    assert len(res) == len(g["pipe_boxes"]) >= 5
    for i, r in enumerate(res):
        assert np.array_equal(r["box"], g["pipe_boxes"][i])
    assert np.array_equal(np.concatenate([r["labels"] for r in res]), g["pipe_labels"])


def test_roofline_layers_match_survey_estimates():
    """roofline/layers.json is regenerated from the OARG layer lists (SURVEY.md 8d asks for exactly that) and the
    algorithmic FLOPs land where the survey's architecture estimate puts them: ~10.2 GFLOP per 960x960 page for the
    detector, ~1.5 GFLOP per 48x320 crop for the recogniser."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fresh = json.loads(subprocess.run([sys.executable, os.path.join(root, "roofline", "make_layers.py")],
                                      capture_output=True, text=True, check=True).stdout)
    with open(os.path.join(root, "roofline", "layers.json")) as f:
        committed = json.load(f)
    for kind, want in (("det", 10.2), ("rec", 1.5)):
        assert committed[kind]["gflop_per_item"] == fresh[kind]["gflop_per_item"]  # committed file is in sync
        assert abs(fresh[kind]["gflop_per_item"] - want) / want < 0.05
        # the fused engine removes the depthwise round trip: strictly fewer bytes than one kernel per layer
        assert fresh[kind]["mb_total_fused_engine"] < fresh[kind]["mb_total_per_layer_kernels"]
    # the optional line-orientation classifier (PP-LCNet x1.0 at 80 x 160 with the width kept after the stem)
    assert committed["cls"]["gflop_per_item"] == fresh["cls"]["gflop_per_item"]
    assert 0.4 < fresh["cls"]["gflop_per_item"] < 0.8
    # spec-only graphs of the next rows: in sync, and the layout encoder graph carries ~93 of RT-DETR-L's 110 GFLOPs
    for kind in ("hgnetv2_l", "rec_server", "layout_encoder"):
        assert committed[kind]["gflop_per_item"] == fresh[kind]["gflop_per_item"]
    assert 85 < fresh["layout_encoder"]["gflop_per_item"] < 100


def test_abi_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in ffi.py must have exactly the layout a C compiler gives the structs of include/oar_b200.h
    (oar_ocr_result grew the optional word-box arrays; a drift here corrupts memory silently)."""
    import ctypes as C
    import os
    import subprocess
    from oar_ocr_b200 import ffi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "oar_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(oar_ocr_result), offsetof(oar_ocr_result, labels),
         offsetof(oar_ocr_result, ms_h2d), offsetof(oar_ocr_result, h2d_bytes), offsetof(oar_ocr_result, cols),
         offsetof(oar_ocr_result, max_wh_ratio), sizeof(oar_pipeline_config));
  printf("%zu %zu\n", sizeof(oar_det_config), offsetof(oar_pipeline_config, image_batch_size));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(oar_ocr_result, line_angle), offsetof(oar_ocr_result, ms_cls),
         sizeof(oar_layout_config), offsetof(oar_layout_config, class_thresholds),
         offsetof(oar_layout_config, image_class_id), offsetof(oar_layout_config, class_unclip));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    a, b, c = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")[:3]
    got = [int(x) for x in a.split()] + [int(x) for x in b.split()] + [int(x) for x in c.split()]
    R, P, L = ffi.OcrResult, ffi.PipelineConfig, ffi.LayoutConfig
    want = [C.sizeof(R), R.labels.offset, R.ms_h2d.offset, R.h2d_bytes.offset, R.cols.offset, R.max_wh_ratio.offset,
            C.sizeof(P), C.sizeof(ffi.DetConfig), P.image_batch_size.offset,
            R.line_angle.offset, R.ms_cls.offset, C.sizeof(L), L.class_thresholds.offset, L.image_class_id.offset,
            L.class_unclip.offset]
    assert got == want
