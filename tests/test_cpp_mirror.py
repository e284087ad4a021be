"""Builds and runs tests/cpp/api_mirror_test.cc: the C++ mirror of the reference's Rust API (include/oar_ocr.hpp)
over the C ABI.  CPU run: error behaviour only.  GPU run: a real predict() through the mirror."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(built_lib, tmp_path):
    exe = str(tmp_path / "api_mirror_test")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "api_mirror_test.cc"), "-o", exe, "-L", libdir,
                           "-loar_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_mirror_cpu(built_lib, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([_build(built_lib, tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_gpu(built_lib, tmp_path, det_blob, rec_blob):
    from oar_ocr_b200 import models
    d, r, c = tmp_path / "det.oarg", tmp_path / "rec.oarg", tmp_path / "cls.oarg"
    d.write_bytes(det_blob)
    r.write_bytes(rec_blob)
    c.write_bytes(models.get_blob("cls"))
    out = subprocess.run([_build(built_lib, tmp_path), str(d), str(r), str(c)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu path ok" in out.stdout
