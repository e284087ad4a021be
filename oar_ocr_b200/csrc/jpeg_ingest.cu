// jpeg_ingest.cu -- encoded pages straight into HBM.
//
// In the reference a page reaches the pipeline through load_image (oar-ocr-core/src/core/utils/image.rs:88:
// image::open -> to_rgb8) on the CPU.  Here the JPEG bytes are decoded by nvJPEG into the u8 HWC RGB page the rest of
// the path reads (NVJPEG_OUTPUT_RGBI = interleaved RGB), so neither the decoded pixels nor a host copy of them ever
// exist on the host.  nvJPEG is looked up at first use (dlopen of the toolkit's libnvjpeg.so.12): the library loads
// and every other entry point works on a box without it, and a missing nvJPEG is a loud OAR_E_UNSUPPORTED, not a
// CPU decode.  Baseline and progressive JPEG, grayscale or 3 components, as nvJPEG supports them.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "engine.cuh"

namespace oar {

namespace {

struct NvJpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) =
      nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                           nvjpegImage_t*, cudaStream_t) = nullptr;
  bool ok = false;
};

const NvJpegApi& nvjpeg_api() {
  static NvJpegApi api = [] {
    NvJpegApi a;
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
#define OAR_NVJ(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym))
    OAR_NVJ(CreateSimple, "nvjpegCreateSimple");
    OAR_NVJ(Destroy, "nvjpegDestroy");
    OAR_NVJ(JpegStateCreate, "nvjpegJpegStateCreate");
    OAR_NVJ(JpegStateDestroy, "nvjpegJpegStateDestroy");
    OAR_NVJ(GetImageInfo, "nvjpegGetImageInfo");
    OAR_NVJ(Decode, "nvjpegDecode");
#undef OAR_NVJ
    a.ok = a.CreateSimple && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo && a.Decode;
    return a;
  }();
  return api;
}

// one decoder (handle + state) per device, created on first use; calls on a context are serialised by its CallGuard,
// and contexts on the same device share the decoder under this mutex
struct Decoder {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
};
std::mutex g_dec_mu;
Decoder g_dec[64];

}  // namespace

// Decodes n JPEG streams into arena buffers of `ctx` (u8 HWC RGB); fills ptrs / hs / ws.
void decode_jpegs_to_device(oar_ctx* ctx, const uint8_t* const* data, const size_t* lens, int n, const uint8_t** ptrs,
                            int32_t* hs, int32_t* ws) {
  const NvJpegApi& api = nvjpeg_api();
  if (!api.ok)
    OAR_FAIL(OAR_E_UNSUPPORTED, "nvJPEG (libnvjpeg.so.12) is not available on this machine: encoded-image ingest is off");
  std::lock_guard<std::mutex> lock(g_dec_mu);
  if (ctx->device < 0 || ctx->device >= 64) OAR_FAIL(OAR_E_INVALID, "device id out of range");
  Decoder& d = g_dec[ctx->device];
  if (!d.handle) {
    if (api.CreateSimple(&d.handle) != NVJPEG_STATUS_SUCCESS || api.JpegStateCreate(d.handle, &d.state) != NVJPEG_STATUS_SUCCESS)
      OAR_FAIL(OAR_E_CUDA, "nvJPEG initialisation failed");
  }
  for (int i = 0; i < n; ++i) {
    if (!data[i] || lens[i] == 0) OAR_FAIL(OAR_E_INVALID, "image %d is empty", i);
    int comps = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    if (api.GetImageInfo(d.handle, data[i], lens[i], &comps, &ss, w, h) != NVJPEG_STATUS_SUCCESS || w[0] <= 0 || h[0] <= 0)
      OAR_FAIL(OAR_E_INVALID, "image %d: not a JPEG stream nvJPEG can parse", i);  // OCRError::ImageLoad in the reference
    uint8_t* dst = ctx->arena.get<uint8_t>((size_t)w[0] * h[0] * 3);
    nvjpegImage_t out{};
    out.channel[0] = dst;
    out.pitch[0] = (size_t)w[0] * 3;
    Launch l(ctx, "nvjpeg_decode", 0, (double)lens[i] + 3.0 * w[0] * h[0]);
    const nvjpegStatus_t rc = api.Decode(d.handle, d.state, data[i], lens[i], NVJPEG_OUTPUT_RGBI, &out, ctx->stream);
    if (rc != NVJPEG_STATUS_SUCCESS) OAR_FAIL(OAR_E_INVALID, "image %d: nvJPEG decode failed (status %d)", i, (int)rc);
    ptrs[i] = dst, hs[i] = h[0], ws[i] = w[0];
  }
}

bool jpeg_ingest_available() { return nvjpeg_api().ok; }

}  // namespace oar
