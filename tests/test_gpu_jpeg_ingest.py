"""Encoded-image ingest: JPEG bytes -> nvJPEG -> HBM -> OAROCR::predict, through the C ABI.
The reference decodes on the CPU (load_image, core/utils/image.rs:88: image::open -> RGB8).  Checked here:
  * the device decode against an independent CPU decoder (Pillow / libjpeg-turbo) -- IDCT and chroma upsampling are
    implementation-defined within the JPEG standard's tolerance, so pixels may differ by a few grey levels, not more;
  * oar_pipeline_run_encoded == oar_pipeline_run on the very pixels the device decoded (bit for bit: the decode feeds
    the same pipeline, nothing else changes);
  * on these pages the boxes and label sequences also equal the oracle's on the CPU-decoded pixels."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _jpeg(img, quality=95, subsampling=0, progressive=False):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=quality, subsampling=subsampling, progressive=progressive)
    return buf.getvalue()


def _pil_decode(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("subsampling,progressive", [(0, False), (2, False), (0, True)])
def test_device_decode_close_to_cpu_decoder(ctx, subsampling, progressive):
    from oar_ocr_b200 import ffi, synth
    page = synth.page(600, 480)[:333, :411]  # odd sizes: partial MCUs at the right / bottom edge
    data = _jpeg(page, 92, subsampling, progressive)
    got = ffi.decode_jpeg(ctx, data)
    want = _pil_decode(data)
    assert got.shape == want.shape == (333, 411, 3)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= (3 if subsampling == 0 else 48), d.max()  # 4:2:0: the chroma upsampling filters differ at edges
    assert d.mean() <= (0.5 if subsampling == 0 else 2.0), d.mean()


def test_grayscale_jpeg(ctx):
    from PIL import Image
    from oar_ocr_b200 import ffi, synth
    g = synth.page(601, 320)[:, :, 1]
    buf = io.BytesIO()
    Image.fromarray(g, "L").save(buf, format="JPEG", quality=95)
    got = ffi.decode_jpeg(ctx, buf.getvalue())
    want = _pil_decode(buf.getvalue())
    assert got.shape == want.shape and np.abs(got.astype(int) - want.astype(int)).max() <= 2


def test_garbage_is_rejected(ctx):
    from oar_ocr_b200 import ffi
    with pytest.raises(ffi.OCRError):
        ffi.decode_jpeg(ctx, b"\x89PNG\r\n\x1a\n" + b"\0" * 64)


def test_pipeline_on_encoded_pages(ctx, det_blob, rec_blob):
    from oar_ocr_b200 import ffi, synth
    from oracle import pipeline
    from oracle.net import OracleNet
    det, rec = ffi.Model(ctx, det_blob), ffi.Model(ctx, rec_blob)
    pages = [synth.page(610 + i, 480) for i in range(3)] + [synth.page(620, 320)]
    jpegs = [_jpeg(p, 95, 0) for p in pages]
    cfg = ffi.pipeline_config(image_batch_size=4, region_batch_size=16, rec_score_thresh=0.0, n_chars=18385)
    cfg.det = ffi.det_config(unclip_ratio=2.0)
    enc = ffi.PipelineBuffers(4)
    res, sizes = ffi.pipeline_run_encoded(det, rec, jpegs, cfg, enc)
    assert sizes == [(480, 480)] * 3 + [(320, 320)]
    assert res.h2d_bytes == sum(len(j) for j in jpegs)  # the encoded bytes are all that crossed the bus
    # the same pipeline on the pixels the device decoded
    decoded = [ffi.decode_jpeg(ctx, j) for j in jpegs]
    arrs, ptrs, hs, ws = ffi._image_table(decoded)
    ref = ffi.PipelineBuffers(4)
    ffi.pipeline_run(det, rec, ptrs, hs, ws, False, cfg, ref)
    n = int(ref.region_off[4])
    assert n >= 30 and np.array_equal(enc.region_off, ref.region_off)
    assert np.array_equal(enc.boxes[:n], ref.boxes[:n]) and np.array_equal(enc.scores[:n], ref.scores[:n])
    nl = int(ref.label_off[n])
    assert np.array_equal(enc.label_off[:n + 1], ref.label_off[:n + 1]) and np.array_equal(enc.labels[:nl], ref.labels[:nl])
    # and against the oracle fed with the CPU decoder's pixels: the same number of regions on every page
    want = pipeline.predict(OracleNet(det_blob), OracleNet(rec_blob), [_pil_decode(j) for j in jpegs], 18385,
                            image_batch_size=4, region_batch_size=16)
    assert [int(enc.region_off[i + 1] - enc.region_off[i]) for i in range(4)] == [len(w) for w in want]
