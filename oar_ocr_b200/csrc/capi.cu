// capi.cu -- the C ABI of liboar_b200.so (include/oar_b200.h) and the host side of the hot path.
//
// Host orchestration mirrors the reference's own call stack (file:line under the reference repo):
//   oar_pipeline_run   = OAROCR::predict                        src/oarocr/ocr.rs:518-659
//   det chunking       = det_batch_size loop                     src/oarocr/ocr.rs:550-592
//   same-shape groups  = DBModel::forward                        oar-ocr-core/src/models/detection/db.rs:281-335
//   limit-side resize  = DetResizeForTest::resize_image_type0    oar-ocr-core/src/processors/resize_detection.rs:243-319
//   reading order      = sort_quad_boxes                         oar-ocr-core/src/processors/sorting.rs:35-84
//   crop pool + flush  = MAX_POOLED_CROPS                        src/oarocr/ocr.rs:603-633
//   wh-ratio chunks    = OAROCR::recognize_global                src/oarocr/ocr.rs:802-897
//   rec batch tensor   = CRNNModel::preprocess_refs              oar-ocr-core/src/models/recognition/crnn.rs:71-125
//   line orientation   = OAROCR::classify_line_orientations      src/oarocr/ocr.rs:755-792
//                        PPLCNetModel::preprocess_refs           oar-ocr-core/src/models/classification/pp_lcnet.rs:139-196
// Everything numeric runs in the CUDA kernels of engine.cu / gemm_tc.cu / prepost.cu / dbpost.cu;
// the host only sequences launches, sorts a few hundred boxes and lays out result buffers.
#include <algorithm>
#include <cmath>
#include <exception>
#include <map>

#include <cstdlib>

#include <functional>
#include <memory>
#include <thread>

#include "engine.cuh"
#include "prepost.cuh"

namespace oar {

thread_local char g_err[1024] = {0};
std::atomic<long long> g_launches{0};
std::atomic<long long> g_submits{0};

void ensure_max_dynamic_smem(const void* kernel, int device, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(kernel, device);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return;
  OAR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[key] = bytes;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace oar

using namespace oar;

// ---------------------------------------------------------------------------
// context plumbing
// ---------------------------------------------------------------------------
void* oar_ctx::pinned_get(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  for (auto& s : pinned) {
    if (s.used + bytes <= s.cap) {
      void* p = s.base + s.used;
      s.used += bytes;
      return p;
    }
  }
  size_t cap = std::max(bytes, (size_t)16 << 20);
  char* p = nullptr;
  OAR_CUDA(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
  pinned.push_back(PinnedSlab{p, cap, bytes});
  return p;
}

void oar_ctx::pinned_reset() {
  if (pinned.size() > 1) {
    size_t total = 0;
    for (auto& s : pinned) {
      total += s.cap;
      cudaFreeHost(s.base);
    }
    pinned.clear();
    char* p = nullptr;
    OAR_CUDA(cudaHostAlloc(&p, total, cudaHostAllocDefault));
    pinned.push_back(PinnedSlab{p, total, 0});
  }
  for (auto& s : pinned) s.used = 0;
}

cudaEvent_t oar_ctx::next_event() {
  if (event_next == event_pool.size()) {
    cudaEvent_t e;
    OAR_CUDA(cudaEventCreate(&e));
    event_pool.push_back(e);
  }
  return event_pool[event_next++];
}

void oar_ctx::begin_call() {
  OAR_CUDA(cudaSetDevice(device));
  OAR_CUDA(cudaStreamSynchronize(stream));
  if (stream_aux) OAR_CUDA(cudaStreamSynchronize(stream_aux));
  if (stream_copy) OAR_CUDA(cudaStreamSynchronize(stream_copy));
  arena.reset();
  arena_aux.reset();
  // debugging aid: OAR_DBG_POISON=1 fills the arena with 0xFF (NaN as f32) before every call, so a kernel that reads
  // activations it (or its producer) never wrote shows up as NaN instead of silently reusing the previous call's data
  static const bool poison = getenv("OAR_DBG_POISON") != nullptr;
  if (poison)
    for (auto& sl : arena.slabs) OAR_CUDA(cudaMemsetAsync(sl.base, 0xFF, sl.cap, stream));
  pinned_reset();
  prof.clear();
  event_next = 0;
}

namespace {

struct CallGuard {
  std::lock_guard<std::mutex> lock;
  explicit CallGuard(oar_ctx* c) : lock(c->mu) { c->begin_call(); }
};

#define API_TRY try {
#define API_CATCH                                                  \
  }                                                                \
  catch (const oar::OarError& e) {                                 \
    cudaGetLastError();                                            \
    return e.code;                                                 \
  }                                                                \
  catch (const std::exception& e) {                                \
    oar::set_error("internal error: %s", e.what());                \
    return OAR_E_CUDA;                                             \
  }                                                                \
  return OAR_OK;

template <typename T>
T* to_device(oar_ctx* ctx, const T* host, size_t n) {
  T* d = ctx->arena.get<T>(n ? n : 1);
  if (n) {
    T* staged = (T*)ctx->pinned_get(n * sizeof(T));
    memcpy(staged, host, n * sizeof(T));
    OAR_CUDA(cudaMemcpyAsync(d, staged, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  }
  return d;
}

// DB detector normalisation constants as DBModelBuilder configures them (db.rs:409-415):
// scale 1/255, ImageNet mean/std applied in output (B,G,R) order, src channels [2,1,0];
// alpha = scale/std, beta = -mean/std (normalization.rs:142-143), all in f32.
void det_norm_coeffs(int src[3], float alpha[3], float beta[3]) {
  const float scale = 1.0f / 255.0f;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) {
    alpha[c] = scale / stdv[c];
    beta[c] = -mean[c] / stdv[c];
    src[c] = 2 - c;
  }
}

// Rust `as u32` from f32: saturating, NaN -> 0
uint32_t f32_as_u32(float v) {
  if (!(v > 0.0f)) return 0u;
  if (v >= 4294967296.0f) return 4294967295u;
  return (uint32_t)v;
}

// resize_image_type0 (resize_detection.rs:243-319): target dims only
void det_resize_dims(uint32_t h, uint32_t w, const oar_det_config& cfg, uint32_t* oh, uint32_t* ow) {
  uint32_t limit = (uint32_t)cfg.limit_side_len;
  float ratio = 1.0f;
  uint32_t mx = std::max(h, w), mn = std::min(h, w);
  if (cfg.limit_type == 0) {
    if (mx > limit) ratio = (float)limit / (float)mx;
  } else if (cfg.limit_type == 1) {
    if (mn < limit) ratio = (float)limit / (float)mn;
  } else {
    ratio = (float)limit / (float)mx;
  }
  uint32_t rh = f32_as_u32((float)h * ratio), rw = f32_as_u32((float)w * ratio);
  uint32_t side_cap = (uint32_t)cfg.max_side_limit;
  if (std::max(rh, rw) > side_cap) {
    float lr = (float)side_cap / (float)std::max(rh, rw);
    rh = f32_as_u32((float)rh * lr);
    rw = f32_as_u32((float)rw * lr);
  }
  rh = std::max((rh + 16) / 32 * 32, 32u);
  rw = std::max((rw + 16) / 32 * 32, 32u);
  *oh = rh;
  *ow = rw;
}

struct DevImage {
  const uint8_t* p;
  int h, w;
};

// one same-shape detection group in flight
struct DetGroup {
  std::vector<int> members;  // indices into the caller's image list
  int H = 0, W = 0;
  float *h_boxes = nullptr, *h_scores = nullptr;
  int32_t* h_counts = nullptr;
  DbPostStatus status;
  cudaEvent_t e_start = nullptr, e_net = nullptr, e_post = nullptr;
};

// Launches DB net -> DB post for one same-shape group; results land in pinned host memory once the streams are
// synchronised.  With `post_lane` the post-process (threshold, labelling, border tracing, box geometry: latency-bound,
// a fraction of the SMs) runs on the context's second lane behind an event, so it overlaps the NEXT group's network.
// `ready` (optional): per image of the caller's list, the event after which its pixels are in HBM.  With uploads in
// flight the network runs in parts of growing size -- a part's pages land while its predecessor is in the detector -- and
// the post-process still sees ONE batch (an image's map does not depend on its batch mates, and the post-process is
// latency-bound: two half-size passes would cost more than the copy that joins the halves).
void launch_det_group(oar_model* det, const std::vector<DevImage>& resized, const std::vector<int32_t>& src_h,
                      const std::vector<int32_t>& src_w, const oar_det_config& cfg, DetGroup& g, int comps_hint,
                      bool timed, bool post_lane, const std::vector<cudaEvent_t>* ready = nullptr) {
  oar_ctx* ctx = det->ctx;
  const int B = (int)g.members.size();
  std::vector<const uint8_t*> ptrs(B);
  std::vector<int32_t> sh(B), sw(B);
  bool aligned = true;
  for (int i = 0; i < B; ++i) {
    ptrs[i] = resized[g.members[i]].p;
    aligned = aligned && (((uintptr_t)ptrs[i] & 3) == 0);
    sh[i] = src_h[g.members[i]];
    sw[i] = src_w[g.members[i]];
  }
  if (timed) {
    g.e_start = ctx->next_event();
    g.e_net = ctx->next_event();
    g.e_post = ctx->next_event();
    cudaEventRecord(g.e_start, ctx->stream);
  }
  // persistent outputs first (the probability map too: the post lane reads it after this lane has moved on), then
  // per-group scratch that the next group may reuse (stream order)
  const int mc = cfg.max_candidates;
  DbPostOut out;
  out.boxes = ctx->arena.get<float>((size_t)B * mc * 8);
  out.scores = ctx->arena.get<float>((size_t)B * mc);
  out.counts = ctx->arena.get<int32_t>(B);
  static const int split_min = getenv("OAR_DET_SPLIT_MIN") ? atoi(getenv("OAR_DET_SPLIT_MIN")) : 16;  // 0 = never split
  bool pending = false;  // pages of this group still on their way?
  if (ready)
    for (int i = 0; i < B; ++i) pending = pending || (*ready)[g.members[i]] != nullptr;
  // Parts: the upload of the first part is the one nothing hides, and a page takes the detector three times as long as
  // the link (0.16 ms against 0.054 ms for 960 x 960), so every part may be up to three times its predecessor and its
  // upload still lands under the predecessor's kernels: 32 pages go as 4 + 12 + 16 (exposed upload 0.22 ms instead of
  // the 0.86 ms of two equal halves; 27.1 -> 26.4 ms per step with 8 + 24).  OAR_DET_SPLIT_PARTS=n forces n equal
  // parts, OAR_DET_SPLIT_FIRST=n two parts with n pages in the first.
  static const int split_parts = getenv("OAR_DET_SPLIT_PARTS") ? std::max(1, atoi(getenv("OAR_DET_SPLIT_PARTS"))) : 0;
  static const int split_first = getenv("OAR_DET_SPLIT_FIRST") ? atoi(getenv("OAR_DET_SPLIT_FIRST")) : 0;
  std::vector<int> bounds{0};  // part p = pages [bounds[p], bounds[p + 1])
  if (pending && split_min > 0 && B >= split_min) {
    if (split_parts > 0) {
      for (int p = 1; p <= std::min(split_parts, B); ++p) bounds.push_back(p * B / std::min(split_parts, B));
    } else if (split_first > 0 && split_first < B) {
      bounds.push_back(split_first), bounds.push_back(B);
    } else {
      int part = std::max(4, B / 8);
      while (bounds.back() + part < B) {
        bounds.push_back(bounds.back() + part);
        part *= 3;
      }
      // a small tail would be a poor detector batch: fold it into the part before it
      if (bounds.size() > 1 && B - bounds.back() < part / 6) bounds.pop_back();
      bounds.push_back(B);
    }
  } else {
    bounds.push_back(B);
  }
  const int halves = (int)bounds.size() - 1;
  const size_t plane = (size_t)g.H * g.W;
  float* prob_keep = (post_lane || halves > 1) ? ctx->arena.get<float>((size_t)B * plane) : nullptr;
  auto mark = ctx->arena.mark();
  const float* prob_p = nullptr;
  for (int hf = 0; hf < halves; ++hf) {
    const int b0 = bounds[hf], nb = bounds[hf + 1] - b0;
    if (ready)
      for (int i = b0; i < b0 + nb; ++i)
        if ((*ready)[g.members[i]]) OAR_CUDA(cudaStreamWaitEvent(ctx->stream, (*ready)[g.members[i]], 0));
    const uint8_t** d_table = (const uint8_t**)to_device(ctx, (const uint8_t* const*)ptrs.data() + b0, nb);
    // the network reads the u8 pages itself: NormalizeImage is folded into the stem convolution (engine.cu / fused_simt.cu)
    Tensor in;
    in.B = nb, in.H = g.H, in.W = g.W, in.C = 3;
    U8Input u8{};
    u8.mode = 0, u8.table = d_table, u8.B = nb, u8.H = g.H, u8.W = g.W, u8.table_aligned = aligned ? 1 : 0;
    det_norm_coeffs(u8.src, u8.a, u8.b);
    Tensor prob = model_forward(det, in, false, nullptr, &u8);
    if (prob.B != nb || prob.H != g.H || prob.W != g.W || prob.C != 1)
      OAR_FAIL(OAR_E_MODEL, "detector output %dx%dx%dx%d does not match its %dx%d input", prob.B, prob.H, prob.W, prob.C,
               g.H, g.W);
    prob_p = prob.p;
    if (prob_keep) {
      OAR_CUDA(cudaMemcpyAsync(prob_keep + (size_t)b0 * plane, prob.p, (size_t)nb * plane * sizeof(float),
                               cudaMemcpyDeviceToDevice, ctx->stream));
      prob_p = prob_keep;
      if (halves > 1 && hf + 1 < halves) ctx->arena.release_to(mark);  // the next half reuses the activations (stream order)
    }
  }
  if (timed) cudaEventRecord(g.e_net, ctx->stream);
  // post lane: the activations can go now (the kept copy of the map lives outside the mark); otherwise the post-process
  // below still reads the map inside them, and they are released after it has been enqueued
  if (post_lane) ctx->arena.release_to(mark);
  cudaStream_t main_stream = ctx->stream;
  if (post_lane) {
    cudaEvent_t ready = ctx->next_event();
    OAR_CUDA(cudaEventRecord(ready, main_stream));
    OAR_CUDA(cudaStreamWaitEvent(ctx->stream_aux, ready, 0));
    std::swap(ctx->stream, ctx->stream_aux);
    std::swap(ctx->arena, ctx->arena_aux);
  }
  try {
    auto pmark = post_lane ? ctx->arena.mark() : mark;
    g.status = db_postprocess_device(ctx, prob_p, B, g.H, g.W, sh.data(), sw.data(), cfg, out, comps_hint);
    g.h_boxes = (float*)ctx->pinned_get((size_t)B * mc * 8 * sizeof(float));
    g.h_scores = (float*)ctx->pinned_get((size_t)B * mc * sizeof(float));
    g.h_counts = (int32_t*)ctx->pinned_get((size_t)B * sizeof(int32_t));
    OAR_CUDA(cudaMemcpyAsync(g.h_counts, out.counts, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    OAR_CUDA(cudaMemcpyAsync(g.h_boxes, out.boxes, (size_t)B * mc * 8 * sizeof(float), cudaMemcpyDeviceToHost,
                             ctx->stream));
    OAR_CUDA(cudaMemcpyAsync(g.h_scores, out.scores, (size_t)B * mc * sizeof(float), cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (timed) cudaEventRecord(g.e_post, ctx->stream);
    if (post_lane) ctx->arena.release_to(pmark);
  } catch (...) {
    if (post_lane) {
      std::swap(ctx->stream, ctx->stream_aux);
      std::swap(ctx->arena, ctx->arena_aux);
    }
    throw;
  }
  if (post_lane) {
    std::swap(ctx->stream, ctx->stream_aux);
    std::swap(ctx->arena, ctx->arena_aux);
  } else {
    ctx->arena.release_to(mark);
  }
}

// DetResizeForTest::apply on device: returns the images the detector sees (resized in HBM when needed)
void det_resize_on_device(oar_ctx* ctx, const std::vector<DevImage>& imgs, const oar_det_config& cfg,
                          std::vector<DevImage>& resized) {
  const int n = (int)imgs.size();
  resized.resize(n);
  std::vector<ResizeJob> jobs;
  int max_sw = 0, max_dw = 0, max_dh = 0;
  for (int i = 0; i < n; ++i) {
    DevImage im = imgs[i];
    if (im.h <= 0 || im.w <= 0) OAR_FAIL(OAR_E_INVALID, "image %d has invalid dimensions %dx%d", i, im.w, im.h);
    if (im.h + im.w < 64) {
      // image_padding (resize_detection.rs:204-220): black canvas of at least 32x32, source at (0,0)
      int nw = std::max(im.w, 32), nh = std::max(im.h, 32);
      if (nw != im.w || nh != im.h) {
        uint8_t* pad = ctx->arena.get<uint8_t>((size_t)nw * nh * 3);
        OAR_CUDA(cudaMemsetAsync(pad, 0, (size_t)nw * nh * 3, ctx->stream));
        OAR_CUDA(cudaMemcpy2DAsync(pad, (size_t)nw * 3, im.p, (size_t)im.w * 3, (size_t)im.w * 3, im.h,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        im = DevImage{pad, nh, nw};
      }
    }
    uint32_t rh, rw;
    det_resize_dims((uint32_t)im.h, (uint32_t)im.w, cfg, &rh, &rw);
    if ((int)rh == im.h && (int)rw == im.w) {
      resized[i] = im;
      continue;
    }
    ResizeJob j;
    j.src = im.p, j.sw = im.w, j.sh = im.h, j.dw = (int)rw, j.dh = (int)rh;
    j.tmp = ctx->arena.get<float>((size_t)rh * im.w * 3);
    j.dst = ctx->arena.get<uint8_t>((size_t)rh * rw * 3);
    jobs.push_back(j);
    max_sw = std::max(max_sw, j.sw), max_dw = std::max(max_dw, j.dw), max_dh = std::max(max_dh, j.dh);
    resized[i] = DevImage{j.dst, (int)rh, (int)rw};
  }
  if (!jobs.empty()) {
    ResizeJob* d_jobs = to_device(ctx, jobs.data(), jobs.size());
    launch_resize_triangle(ctx, d_jobs, (int)jobs.size(), max_sw, max_dw, max_dh);
  }
}

struct DetResult {
  std::vector<std::vector<float>> boxes;   // per image: count*8, discovery order
  std::vector<std::vector<float>> scores;  // per image: count
  float ms_net = 0, ms_post = 0;
};

// TextDetectionAdapter::execute for one list of device images, chunked by image_batch_size.
// `ready` (optional): per image, the event after which its pixels are in HBM (uploads run on the copy stream).
void run_detection(oar_model* det, const std::vector<DevImage>& imgs, const oar_det_config& cfg, int batch_size,
                   DetResult& res, bool timed, const std::vector<cudaEvent_t>* ready = nullptr) {
  oar_ctx* ctx = det->ctx;
  const int n = (int)imgs.size();
  if (cfg.max_candidates <= 0) OAR_FAIL(OAR_E_INVALID, "max_candidates must be positive");
  res.boxes.assign(n, {});
  res.scores.assign(n, {});
  auto wait_for = [&](int i) {
    if (ready && (*ready)[i]) OAR_CUDA(cudaStreamWaitEvent(ctx->stream, (*ready)[i], 0));
  };
  std::vector<DevImage> resized;
  if (ready) {  // only the images that get resized need their pixels now
    oar_det_config c = cfg;
    for (int i = 0; i < n; ++i) {
      uint32_t rh, rw;
      det_resize_dims((uint32_t)std::max(imgs[i].h, 1), (uint32_t)std::max(imgs[i].w, 1), c, &rh, &rw);
      if ((int)rh != imgs[i].h || (int)rw != imgs[i].w || imgs[i].h + imgs[i].w < 64) wait_for(i);
    }
  }
  det_resize_on_device(ctx, imgs, cfg, resized);
  std::vector<int32_t> src_h(n), src_w(n);
  for (int i = 0; i < n; ++i) src_h[i] = imgs[i].h, src_w[i] = imgs[i].w;
  std::vector<DetGroup> groups;
  batch_size = std::max(batch_size, 1);
  // An image's boxes do not depend on its batch mates, so a same-shape group of the reference (db.rs:299-309) is cut
  // into sub-groups of at most DET_SUB pages: the post-process of one sub-group (second lane) and the upload of the
  // later pages (copy stream) overlap the network of the next.
  // Measured on B200 (32 pages 960x960): sub-groups of 8 made the step SLOWER (30.2 vs 29.1 ms) -- the post-process
  // kernels on the second lane take SM slots from the persistent one-CTA-per-SM network kernels, whose grids then run
  // with stragglers -- so the split stays opt-in (OAR_DET_SUB=8).
  static const int det_sub = getenv("OAR_DET_SUB") ? atoi(getenv("OAR_DET_SUB")) : 0;
  const bool lanes = det_sub > 0 && !ctx->profile && ctx->stream_aux;
  for (int start = 0; start < n; start += batch_size) {
    int end = std::min(n, start + batch_size);
    size_t first_group = groups.size();
    for (int i = start; i < end; ++i) {  // first-seen shape order, db.rs:299-309
      DetGroup* hit = nullptr;
      for (size_t gi = first_group; gi < groups.size(); ++gi)
        if (groups[gi].H == resized[i].h && groups[gi].W == resized[i].w &&
            (!lanes || (int)groups[gi].members.size() < det_sub))
          hit = &groups[gi];
      if (!hit) {
        groups.emplace_back();
        hit = &groups.back();
        hit->H = resized[i].h, hit->W = resized[i].w;
      }
      hit->members.push_back(i);
    }
  }
  const bool post_lane = lanes && groups.size() >= 2;
  for (auto& g : groups) launch_det_group(det, resized, src_h, src_w, cfg, g, 0, timed, post_lane, ready);
  if (post_lane) {
    cudaEvent_t join = ctx->next_event();
    OAR_CUDA(cudaEventRecord(join, ctx->stream_aux));
    OAR_CUDA(cudaStreamWaitEvent(ctx->stream, join, 0));
  }
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& g : groups) {
    int need = 0;
    int tries = 0;
    while (db_postprocess_check(g.status, &need) == 1) {
      if (++tries > 3) OAR_FAIL(OAR_E_CAPACITY, "DB post-process component bound did not converge");
      launch_det_group(det, resized, src_h, src_w, cfg, g, need, false, false);
      OAR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    const int mc = cfg.max_candidates;
    for (size_t k = 0; k < g.members.size(); ++k) {
      int cnt = g.h_counts[k];
      int img = g.members[k];
      res.boxes[img].assign(g.h_boxes + k * (size_t)mc * 8, g.h_boxes + k * (size_t)mc * 8 + (size_t)cnt * 8);
      res.scores[img].assign(g.h_scores + k * (size_t)mc, g.h_scores + k * (size_t)mc + cnt);
    }
    if (timed && g.e_start && tries == 0) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, g.e_start, g.e_net);
      cudaEventElapsedTime(&b, g.e_net, g.e_post);
      res.ms_net += a, res.ms_post += b;
    }
  }
}

// sort_quad_boxes (sorting.rs:35-84): stable sort by (y_min, x_min), then the adjacent-swap pass
void sort_quads_host(const float* boxes, int n, std::vector<int>& order) {
  order.resize(n);
  std::vector<float> ymin(n), xmin(n);
  for (int i = 0; i < n; ++i) {
    order[i] = i;
    const float* b = boxes + (size_t)i * 8;
    float mx = INFINITY, my = INFINITY;
    for (int k = 0; k < 4; ++k) {
      mx = std::fmin(mx, b[2 * k]);
      my = std::fmin(my, b[2 * k + 1]);
    }
    xmin[i] = mx, ymin[i] = my;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (ymin[a] < ymin[b]) return true;
    if (ymin[a] > ymin[b]) return false;
    return xmin[a] < xmin[b];
  });
  for (int i = 0; i + 1 < n; ++i) {
    for (int j = i; j >= 0; --j) {
      int cur = order[j], nxt = order[j + 1];
      if (std::fabs(ymin[nxt] - ymin[cur]) < 10.0f && xmin[nxt] < xmin[cur])
        std::swap(order[j], order[j + 1]);
      else
        break;
    }
  }
}

constexpr int REC_H = 48, REC_W = 320, REC_MAX_W = 3200;  // DEFAULT_REC_IMAGE_SHAPE, constants.rs:8
constexpr int MAX_POOLED_CROPS = 4096;                    // ocr.rs:603

int64_t f32_as_usize(float v) {
  if (!(v > 0.0f)) return 0;
  if (v >= 9.2e18f) return INT64_MAX;
  return (int64_t)v;
}

struct RecCrop {
  const uint8_t* p;  // device u8 HWC
  int h, w;
};

struct RecBatchOut {
  int n = 0, T = 0;
  int32_t *h_labels = nullptr, *h_cols = nullptr, *h_lens = nullptr;
  float* h_scores = nullptr;
};

// CRNNModel::forward_refs (crnn.rs:247-293) on one batch of device crops; outputs arrive in pinned
// memory after the stream is synchronised.
void launch_rec_batch(oar_model* rec, const RecCrop* crops, int n, int n_chars, RecBatchOut& out) {
  oar_ctx* ctx = rec->ctx;
  out.n = n;
  if (n == 0) return;
  const float base = (float)REC_W / (float)std::max(REC_H, 1);
  float max_ratio = base;
  for (int i = 0; i < n; ++i) {
    if (crops[i].h <= 0 || crops[i].w <= 0) OAR_FAIL(OAR_E_INVALID, "crop %d has invalid dimensions", i);
    max_ratio = std::fmax(max_ratio, (float)crops[i].w / (float)std::max(crops[i].h, 1));
  }
  const int tensor_w = (int)std::min<int64_t>(f32_as_usize((float)REC_H * max_ratio), REC_MAX_W);
  auto mark = ctx->arena.mark();
  std::vector<ResizeJob> jobs(n);
  std::vector<CrnnJob> cj(n);
  int max_sw = 0, max_dw = 0;
  for (int i = 0; i < n; ++i) {
    float ratio = (float)crops[i].w / (float)crops[i].h;
    int rw = (int)std::min<int64_t>(f32_as_usize(std::ceil((float)REC_H * ratio)), tensor_w);
    ResizeJob& j = jobs[i];
    j.src = crops[i].p, j.sw = crops[i].w, j.sh = crops[i].h, j.dw = rw, j.dh = REC_H;
    j.tmp = ctx->arena.get<float>((size_t)REC_H * crops[i].w * 3);
    j.dst = ctx->arena.get<uint8_t>((size_t)REC_H * std::max(rw, 1) * 3);
    cj[i].src = j.dst, cj[i].rw = rw;
    max_sw = std::max(max_sw, j.sw), max_dw = std::max(max_dw, rw);
  }
  ResizeJob* d_jobs = to_device(ctx, jobs.data(), jobs.size());
  CrnnJob* d_cj = to_device(ctx, cj.data(), cj.size());
  launch_resize_triangle(ctx, d_jobs, n, max_sw, max_dw, REC_H);
  Tensor in;
  in.B = n, in.H = REC_H, in.W = tensor_w, in.C = 3;
  U8Input u8{};  // normalize_crnn_chw_into folded into the stem convolution
  u8.mode = 1, u8.jobs = d_cj, u8.B = n, u8.H = REC_H, u8.W = tensor_w;
  CtcOut ctc;
  model_forward(rec, in, false, &ctc, &u8);
  if (ctc.B != n || ctc.T <= 0) OAR_FAIL(OAR_E_MODEL, "recognizer produced no CTC output");
  const int T = ctc.T;
  out.T = T;
  size_t bt = (size_t)n * T;
  int32_t* d_labels = ctx->arena.get<int32_t>(bt);
  int32_t* d_cols = ctx->arena.get<int32_t>(bt);
  int32_t* d_lens = ctx->arena.get<int32_t>(n);
  float* d_scores = ctx->arena.get<float>(n);
  launch_ctc_decode(ctx, ctc.idx, ctc.prob, n, T, n_chars, d_labels, d_cols, d_lens, d_scores);
  out.h_labels = (int32_t*)ctx->pinned_get(bt * sizeof(int32_t));
  out.h_cols = (int32_t*)ctx->pinned_get(bt * sizeof(int32_t));
  out.h_lens = (int32_t*)ctx->pinned_get(n * sizeof(int32_t));
  out.h_scores = (float*)ctx->pinned_get(n * sizeof(float));
  cudaStream_t st = ctx->stream;
  OAR_CUDA(cudaMemcpyAsync(out.h_labels, d_labels, bt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(out.h_cols, d_cols, bt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(out.h_lens, d_lens, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(out.h_scores, d_scores, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  ctx->arena.release_to(mark);
}

// ImageNet normalisation in RGB order as PPLCNetModelBuilder::build configures it (pp_lcnet.rs:400-412)
void cls_norm_coeffs(int src[3], float alpha[3], float beta[3]) {
  det_norm_coeffs(src, alpha, beta);
  for (int c = 0; c < 3; ++c) src[c] = c;
}

constexpr int CLS_H = 80, CLS_W = 160;  // TextLineOrientationAdapter::DEFAULT_INPUT_SHAPE
constexpr int CLS_CHUNK = 256;          // crops per classifier launch (bounds the activation arena, not a result knob)

// PPLCNetModel::forward_refs (pp_lcnet.rs:139-196, 200-243) on one batch of device crops: direct Triangle resize to
// iw x ih, normalise, network, top-1.  d_ids / d_scores [n] and (optionally) d_probs [n][C] are device outputs owned
// by the caller; returns C.  Scratch is released on return (stream order makes the reuse safe).
int launch_cls_batch(oar_model* cls, const RecCrop* crops, int n, int ih, int iw, int32_t* d_ids, float* d_scores,
                     float* d_probs, size_t probs_cap) {
  oar_ctx* ctx = cls->ctx;
  auto mark = ctx->arena.mark();
  std::vector<ResizeJob> jobs(n);
  std::vector<const uint8_t*> ptrs(n);
  int max_sw = 0;
  for (int i = 0; i < n; ++i) {
    if (crops[i].h <= 0 || crops[i].w <= 0) OAR_FAIL(OAR_E_INVALID, "crop %d has invalid dimensions", i);
    ResizeJob& j = jobs[i];
    j.src = crops[i].p, j.sw = crops[i].w, j.sh = crops[i].h, j.dw = iw, j.dh = ih;
    j.tmp = ctx->arena.get<float>((size_t)ih * crops[i].w * 3);
    j.dst = ctx->arena.get<uint8_t>((size_t)ih * iw * 3);
    ptrs[i] = j.dst;
    max_sw = std::max(max_sw, j.sw);
  }
  bool aligned = true;
  for (int i = 0; i < n; ++i) aligned = aligned && (((uintptr_t)ptrs[i] & 3) == 0);
  ResizeJob* d_jobs = to_device(ctx, jobs.data(), jobs.size());
  const uint8_t** d_table = (const uint8_t**)to_device(ctx, (const uint8_t* const*)ptrs.data(), (size_t)n);
  launch_resize_triangle(ctx, d_jobs, n, max_sw, iw, ih);
  Tensor in;
  in.B = n, in.H = ih, in.W = iw, in.C = 3;
  U8Input u8{};
  u8.mode = 0, u8.table = d_table, u8.B = n, u8.H = ih, u8.W = iw, u8.table_aligned = aligned ? 1 : 0;
  cls_norm_coeffs(u8.src, u8.a, u8.b);
  CtcOut head;
  Tensor probs = model_forward(cls, in, /*want_probs=*/true, &head, &u8);
  if (!probs.p || probs.B != n || probs.H * probs.W != 1 || probs.C <= 0)
    OAR_FAIL(OAR_E_MODEL, "classifier output %dx%dx%dx%d is not [n,1,1,classes]", probs.B, probs.H, probs.W, probs.C);
  const int C = probs.C;
  launch_cls_top1(ctx, probs.p, n, C, d_ids, d_scores);
  if (d_probs) {
    if ((size_t)n * C > probs_cap) OAR_FAIL(OAR_E_CAPACITY, "probabilities need %zu floats", (size_t)n * C);
    OAR_CUDA(cudaMemcpyAsync(d_probs, probs.p, (size_t)n * C * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  ctx->arena.release_to(mark);
  return C;
}

// Structural check of one OARG op record at load time: tensor ids in range, weight slices inside the weight array
// (no signed or unsigned wrap), parameters positive where the kernels divide or index by them, and every weight slice
// exactly as long as the op's parameters imply -- engine.cu indexes weights by those parameters, never by w_len, so a
// short slice would otherwise be read past its end.  Returns nullptr or the reason.
const char* validate_op(const OpRec& op, uint32_t n_tensors, uint64_t n_w) {
  if (op.in0 < 0 || op.in0 >= (int)n_tensors || op.out < 0 || op.out >= (int)n_tensors) return "tensor id out of range";
  if (op.in1 < -1 || op.in1 >= (int)n_tensors) return "second input id out of range";
  for (int k = 0; k < 4; ++k) {
    if (op.w_off[k] < 0 || op.w_len[k] < 0) return "negative weight slice";
    if ((uint64_t)op.w_off[k] > n_w || (uint64_t)op.w_len[k] > n_w - (uint64_t)op.w_off[k]) return "weight slice out of range";
  }
  const int32_t* p = op.p;
  auto pos = [](int64_t v, int64_t hi) { return v > 0 && v <= hi; };
  const int64_t CMAX = 1 << 20;  // channels / classes
  int64_t want[4] = {0, 0, 0, 0};
  switch (op.type) {
    case OP_CONV:
      if (!pos(p[0], 64) || !pos(p[1], 64) || !pos(p[2], 64) || !pos(p[3], 64) || p[4] < 0 || p[4] > 64 || p[5] < 0 ||
          p[5] > 64 || !pos(p[6], CMAX) || !pos(p[7], CMAX))
        return "convolution kernel / stride / channels out of range";
      if (p[10] < 0 || p[11] < 0 || (p[11] > 0 && (int64_t)p[10] + p[7] > p[11])) return "channel slice outside its tensor";
      want[0] = (int64_t)p[7] * p[0] * p[1] * p[6], want[1] = p[7];
      break;
    case OP_DWCONV:
      if (!pos(p[0], 64) || !pos(p[1], 64) || !pos(p[2], 64) || !pos(p[3], 64) || p[4] < 0 || p[4] > 64 || p[5] < 0 ||
          p[5] > 64 || !pos(p[6], CMAX))
        return "depthwise kernel / stride / channels out of range";
      want[0] = (int64_t)p[0] * p[1] * p[6], want[1] = p[6];
      break;
    case OP_SE:
      if (!pos(p[0], CMAX) || !pos(p[1], CMAX)) return "squeeze-excite widths out of range";
      want[0] = (int64_t)p[1] * p[0], want[1] = p[1], want[2] = (int64_t)p[0] * p[1], want[3] = p[0];
      break;
    case OP_ADD:
      if (op.in1 < 0) return "add needs two inputs";
      break;
    case OP_UPADD:
      if (op.in1 < 0 || !pos(p[0], 64)) return "upsample-add needs two inputs and a positive scale";
      break;
    case OP_UPSAMPLE:
      if (!pos(p[0], 64) || p[10] < 0 || p[11] < 0) return "upsample scale / slice out of range";
      break;
    case OP_DECONV2:
      if (!pos(p[0], CMAX) || !pos(p[1], CMAX)) return "transposed-conv channels out of range";
      want[0] = (int64_t)4 * p[1] * p[0], want[1] = p[1];
      break;
    case OP_AVGPOOL:
      if (!((p[0] == 0 && p[1] == 0) || (pos(p[0], 1 << 16) && pos(p[1], 1 << 16) && pos(p[2], 1 << 16) && pos(p[3], 1 << 16))))
        return "pool window / stride out of range";
      break;
    case OP_LAYERNORM:
      if (!pos(p[0], CMAX)) return "layer-norm width out of range";
      want[0] = p[0], want[1] = p[0];
      break;
    case OP_ATTN:
      if (!pos(p[0], CMAX) || !pos(p[1], p[0]) || p[0] % p[1]) return "attention width / heads out of range";
      want[0] = (int64_t)3 * p[0] * p[0], want[1] = 3 * (int64_t)p[0], want[2] = (int64_t)p[0] * p[0], want[3] = p[0];
      break;
    case OP_CTC_HEAD:
      if (!pos(p[0], CMAX) || !pos(p[1], CMAX)) return "head width / classes out of range";
      want[0] = (int64_t)p[1] * p[0], want[1] = p[1];
      break;
    case OP_PAD:
      if (p[0] < 0 || p[1] < 0 || p[2] < 0 || p[3] < 0 || p[0] > 4096 || p[1] > 4096 || p[2] > 4096 || p[3] > 4096)
        return "padding out of range";
      break;
    case OP_MAXPOOL:
      if (!pos(p[0], 64) || !pos(p[1], 64) || !pos(p[2], 64) || !pos(p[3], 64)) return "pool window / stride out of range";
      break;
    case OP_TOKENS:
      if (p[0] < 0 || !pos(p[1], 1 << 24) || p[0] >= p[1]) return "token rows out of range";
      break;
    default:
      return "unknown op type";
  }
  for (int k = 0; k < 4; ++k)
    if (op.w_len[k] != want[k]) return "weight slice length does not match the op's parameters";
  return nullptr;
}

struct BlobHeader {
  uint32_t version, kind, n_ops, n_tensors;
  uint64_t n_w;
};

// Parses and validates an OARG blob (host only).  Every size is bounded by the blob length BEFORE it is multiplied, so
// a crafted 64-bit weight count cannot wrap the length check.
void check_blob(const uint8_t* p, size_t len, BlobHeader& h, std::vector<OpRec>* ops_out) {
  if (len < 28 || memcmp(p, "OARG", 4) != 0) OAR_FAIL(OAR_E_MODEL, "not an OARG model blob");
  memcpy(&h.version, p + 4, 4);
  memcpy(&h.kind, p + 8, 4);
  memcpy(&h.n_ops, p + 12, 4);
  memcpy(&h.n_tensors, p + 16, 4);
  memcpy(&h.n_w, p + 20, 8);
  if (h.version != 1) OAR_FAIL(OAR_E_MODEL, "unsupported OARG version %u", h.version);
  if (h.kind > OAR_KIND_FEAT) OAR_FAIL(OAR_E_MODEL, "unknown model kind %u", h.kind);
  const size_t body = len - 28;
  if ((size_t)h.n_ops > body / sizeof(OpRec))
    OAR_FAIL(OAR_E_MODEL, "truncated model blob: %u ops do not fit %zu bytes", h.n_ops, len);
  const size_t w_bytes = body - (size_t)h.n_ops * sizeof(OpRec);
  if (h.n_w > w_bytes / 4)
    OAR_FAIL(OAR_E_MODEL, "truncated model blob: %zu bytes hold fewer than %llu weights", len, (unsigned long long)h.n_w);
  if (h.n_tensors == 0 || h.n_tensors > (1u << 20)) OAR_FAIL(OAR_E_MODEL, "implausible tensor count %u", h.n_tensors);
  if (h.n_ops == 0) OAR_FAIL(OAR_E_MODEL, "model has no operations");
  std::vector<OpRec> ops(h.n_ops);
  memcpy(ops.data(), p + 28, (size_t)h.n_ops * sizeof(OpRec));
  for (size_t i = 0; i < ops.size(); ++i) {
    const char* why = validate_op(ops[i], h.n_tensors, h.n_w);
    if (why) OAR_FAIL(OAR_E_MODEL, "op %zu (type %d): %s", i, ops[i].type, why);
  }
  if (ops_out) *ops_out = std::move(ops);
}

void require_device(oar_ctx* ctx) {
  if (!ctx) OAR_FAIL(OAR_E_INVALID, "null context");
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

void oar_det_config_default(oar_det_config* cfg) {
  if (!cfg) return;
  cfg->thresh = 0.3f;
  cfg->box_thresh = 0.6f;
  cfg->unclip_ratio = 2.0f;
  cfg->max_candidates = 1000;
  cfg->min_size = 3.0f;
  cfg->limit_side_len = 960;
  cfg->limit_type = 0;
  cfg->max_side_limit = 4000;
}

void oar_pipeline_config_default(oar_pipeline_config* cfg) {
  if (!cfg) return;
  oar_det_config_default(&cfg->det);
  cfg->image_batch_size = 8;
  cfg->region_batch_size = 64;
  cfg->rec_score_thresh = 0.0f;
  cfg->n_chars = 18385;
}

const char* oar_last_error(void) { return oar::g_err; }
int32_t oar_version(void) { return 100; }
int64_t oar_launch_count(void) { return oar::g_launches.load(); }
int64_t oar_submit_count(void) { return oar::g_submits.load(); }

int32_t oar_ctx_create(int32_t device_id, oar_ctx** out) {
  if (!out) {
    set_error("null output pointer");
    return OAR_E_INVALID;
  }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); liboar_b200 has no CPU fallback", cudaGetErrorString(e));
    return OAR_E_NO_DEVICE;
  }
  if (device_id < 0 || device_id >= count) {
    set_error("device_id %d out of range (0..%d)", device_id, count - 1);
    return OAR_E_INVALID;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess || prop.major != 10) {
    set_error("device %d is not an sm_100 (Blackwell B200) part; this library ships sm_100a code only", device_id);
    return OAR_E_NO_DEVICE;
  }
  API_TRY
  OAR_CUDA(cudaSetDevice(device_id));
  oar_ctx* c = new oar_ctx();
  c->device = device_id;
  c->sm_count = prop.multiProcessorCount;
  OAR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  OAR_CUDA(cudaStreamCreateWithFlags(&c->stream_aux, cudaStreamNonBlocking));
  OAR_CUDA(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
  *out = c;
  API_CATCH
}

void oar_ctx_destroy(oar_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stream_aux) cudaStreamSynchronize(ctx->stream_aux);
  graph_cache_free(ctx);
  ctx->arena.release();
  ctx->arena_aux.release();
  for (auto& s : ctx->pinned) cudaFreeHost(s.base);
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->timer0) cudaEventDestroy(ctx->timer0);
  if (ctx->timer1) cudaEventDestroy(ctx->timer1);
  cudaFree(ctx->flush_buf);
  cudaStreamDestroy(ctx->stream);
  if (ctx->stream_aux) cudaStreamDestroy(ctx->stream_aux);
  if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
  delete ctx;
}

int32_t oar_ctx_synchronize(oar_ctx* ctx) {
  API_TRY
  require_device(ctx);
  OAR_CUDA(cudaSetDevice(ctx->device));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_model_validate_blob(const void* bytes, size_t len) {
  API_TRY
  if (!bytes) OAR_FAIL(OAR_E_INVALID, "null argument");
  BlobHeader h;
  check_blob((const uint8_t*)bytes, len, h, nullptr);
  API_CATCH
}

int32_t oar_model_load_blob(oar_ctx* ctx, const void* bytes, size_t len, oar_model** out) {
  API_TRY
  require_device(ctx);
  if (!bytes || !out) OAR_FAIL(OAR_E_INVALID, "null argument");
  *out = nullptr;
  const uint8_t* p = (const uint8_t*)bytes;
  BlobHeader h;
  std::vector<OpRec> ops;
  check_blob(p, len, h, &ops);
  std::lock_guard<std::mutex> lock(ctx->mu);
  OAR_CUDA(cudaSetDevice(ctx->device));
  oar_model* m = new oar_model();
  m->ctx = ctx;
  static std::atomic<uint64_t> next_uid{1};
  m->uid = next_uid++;
  m->kind = (int)h.kind;
  m->n_tensors = (int)h.n_tensors;
  m->ops = std::move(ops);
  m->n_weights = h.n_w;
  cudaError_t e = cudaMalloc(&m->d_weights, std::max<size_t>(h.n_w, 1) * 4);
  if (e == cudaSuccess)
    e = cudaMemcpy(m->d_weights, p + 28 + m->ops.size() * sizeof(OpRec), h.n_w * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(m->d_weights);
    delete m;
    OAR_FAIL(OAR_E_CUDA, "weight upload failed: %s", cudaGetErrorString(e));
  }
  m->engine = 2;
  if (const char* cap = getenv("OAR_DBG_ENGINE_CAP")) m->engine = std::min(2, atoi(cap));
  try {
    tc_model_init(m);
  } catch (...) {
    tc_model_free(m);
    cudaFree(m->d_weights);
    delete m;
    throw;
  }
  *out = m;
  API_CATCH
}

int32_t oar_onnx_to_oarg(const void* onnx, size_t len, int32_t kind_hint, void* out, size_t cap, size_t* out_len) {
  API_TRY
  if (!onnx || !out_len) OAR_FAIL(OAR_E_INVALID, "null argument");
  if (kind_hint > OAR_KIND_CLS) OAR_FAIL(OAR_E_INVALID, "unknown model kind %d", kind_hint);
  std::vector<uint8_t> blob = onnx_to_oarg(onnx, len, kind_hint);
  *out_len = blob.size();
  if (!out) return OAR_OK;  // size query
  if (blob.size() > cap) OAR_FAIL(OAR_E_CAPACITY, "OARG blob needs %zu bytes, capacity %zu", blob.size(), cap);
  memcpy(out, blob.data(), blob.size());
  API_CATCH
}

int32_t oar_model_load_onnx(oar_ctx* ctx, const void* bytes, size_t len, int32_t kind, oar_model** out) {
  if (!bytes || !out) {
    set_error("null argument");
    return OAR_E_INVALID;
  }
  *out = nullptr;
  if (len >= 4 && memcmp(bytes, "OARG", 4) == 0) return oar_model_load_blob(ctx, bytes, len, out);
  std::vector<uint8_t> blob;
  {
    API_TRY
    require_device(ctx);
    if (kind < -1 || kind > OAR_KIND_CLS) OAR_FAIL(OAR_E_INVALID, "unknown model kind %d", kind);
    // a classifier's graph ends in the same MatMul + Softmax pattern as a CTC head: the caller's role decides
    blob = onnx_to_oarg(bytes, len, kind == OAR_KIND_CLS ? OAR_KIND_CLS : -1);
    uint32_t got;
    memcpy(&got, blob.data() + 8, 4);
    if (kind >= 0 && (int)got != kind)
      OAR_FAIL(OAR_E_MODEL, "the ONNX graph is a %s model, a %s model was requested", got == 0 ? "detection" : got == 1 ? "recognition" : "classification",
               kind == 0 ? "detection" : kind == 1 ? "recognition" : "classification");
    }
    catch (const oar::OarError& e) {
      cudaGetLastError();
      return e.code;
    }
    catch (const std::exception& e) {
      oar::set_error("internal error: %s", e.what());
      return OAR_E_MODEL;
    }
  }
  return oar_model_load_blob(ctx, blob.data(), blob.size(), out);
}

void oar_model_destroy(oar_model* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  if (m->ctx->stream_aux) cudaStreamSynchronize(m->ctx->stream_aux);
  graph_cache_drop_model(m->ctx, m->uid);
  tc_model_free(m);
  cudaFree(m->d_weights);
  delete m;
}

int32_t oar_model_kind(const oar_model* m) { return m ? m->kind : OAR_E_INVALID; }

int32_t oar_model_set_engine(oar_model* m, int32_t engine) {
  if (!m || engine < 0 || engine > 2) {
    set_error("invalid engine selector");
    return OAR_E_INVALID;
  }
  // debugging aid: OAR_DBG_ENGINE_CAP=1 maps engine 2 requests onto engine 1 (bisecting the fused kernels)
  static const char* cap = getenv("OAR_DBG_ENGINE_CAP");
  if (cap && engine > atoi(cap)) engine = atoi(cap);
  m->engine = engine;
  return OAR_OK;
}

int32_t oar_infer_f32(oar_model* m, const float* in, const int64_t in_shape[4], float* out, size_t out_cap,
                      int64_t out_shape[4]) {
  API_TRY
  if (!m || !in || !in_shape || !out || !out_shape) OAR_FAIL(OAR_E_INVALID, "null argument");
  int64_t B = in_shape[0], C = in_shape[1], H = in_shape[2], W = in_shape[3];
  if (B <= 0 || C != 3 || H <= 0 || W <= 0) OAR_FAIL(OAR_E_INVALID, "input must be [B,3,H,W] with positive dims");
  oar_ctx* ctx = m->ctx;
  CallGuard guard(ctx);
  size_t n = (size_t)B * C * H * W;
  float* d_nchw = ctx->arena.get<float>(n);
  OAR_CUDA(cudaMemcpyAsync(d_nchw, in, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  Tensor x;
  x.B = (int)B, x.H = (int)H, x.W = (int)W, x.C = 3;
  x.p = ctx->arena.get<float>(n);
  launch_nchw_to_nhwc(ctx, d_nchw, x.p, (int)B, 3, (int)H, (int)W);
  CtcOut ctc;
  Tensor y = model_forward(m, x, true, &ctc);
  if (!y.p) OAR_FAIL(OAR_E_MODEL, "model produced no output");
  if (m->kind == OAR_KIND_DET) {
    out_shape[0] = y.B, out_shape[1] = y.C, out_shape[2] = y.H, out_shape[3] = y.W;
    if (y.C != 1) OAR_FAIL(OAR_E_MODEL, "detector output has %d channels", y.C);
  } else if (m->kind == OAR_KIND_FEAT) {
    out_shape[0] = y.B, out_shape[1] = y.H, out_shape[2] = y.W, out_shape[3] = y.C;  // the feature map as stored: NHWC
  } else {
    out_shape[0] = y.B, out_shape[1] = (int64_t)y.H * y.W, out_shape[2] = y.C, out_shape[3] = 1;
  }
  if (y.numel() > out_cap) OAR_FAIL(OAR_E_CAPACITY, "output needs %zu floats, capacity %zu", y.numel(), out_cap);
  OAR_CUDA(cudaMemcpyAsync(out, y.p, y.numel() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_normalize_chw(oar_ctx* ctx, const uint8_t* rgb, int32_t batch, int32_t h, int32_t w,
                          const int32_t src_channels[3], const float alpha[3], const float beta[3], float* out) {
  API_TRY
  require_device(ctx);
  if (batch < 0 || h < 0 || w < 0) OAR_FAIL(OAR_E_INVALID, "negative dimension");
  for (int c = 0; c < 3; ++c)
    if (src_channels[c] < 0 || src_channels[c] > 2) OAR_FAIL(OAR_E_INVALID, "source channel out of range");
  size_t px = (size_t)batch * h * w;
  if (px == 0) return OAR_OK;
  if (!rgb || !out) OAR_FAIL(OAR_E_INVALID, "null buffer");
  CallGuard guard(ctx);
  uint8_t* d_in = ctx->arena.get<uint8_t>(px * 3);
  float* d_out = ctx->arena.get<float>(px * 3);
  OAR_CUDA(cudaMemcpyAsync(d_in, rgb, px * 3, cudaMemcpyHostToDevice, ctx->stream));
  int src[3] = {src_channels[0], src_channels[1], src_channels[2]};
  launch_normalize(ctx, d_in, nullptr, true, d_out, batch, h, w, src, alpha, beta, 0);
  OAR_CUDA(cudaMemcpyAsync(out, d_out, px * 12, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_db_postprocess(oar_ctx* ctx, const float* pred, int32_t batch, int32_t h, int32_t w, const int32_t* src_h,
                           const int32_t* src_w, const oar_det_config* cfg, float* boxes, float* scores,
                           int32_t* counts) {
  API_TRY
  require_device(ctx);
  if (!cfg || !counts) OAR_FAIL(OAR_E_INVALID, "null argument");
  if (batch <= 0) return OAR_OK;
  if (h <= 0 || w <= 0 || !pred || !boxes || !scores || !src_h || !src_w) OAR_FAIL(OAR_E_INVALID, "bad argument");
  if (cfg->max_candidates <= 0) OAR_FAIL(OAR_E_INVALID, "max_candidates must be positive");
  CallGuard guard(ctx);
  size_t total = (size_t)batch * h * w;
  const int mc = cfg->max_candidates;
  float* d_pred = ctx->arena.get<float>(total);
  OAR_CUDA(cudaMemcpyAsync(d_pred, pred, total * 4, cudaMemcpyHostToDevice, ctx->stream));
  DbPostOut out;
  out.boxes = ctx->arena.get<float>((size_t)batch * mc * 8);
  out.scores = ctx->arena.get<float>((size_t)batch * mc);
  out.counts = ctx->arena.get<int32_t>(batch);
  int hint = 0;
  for (int tries = 0;; ++tries) {
    auto mark = ctx->arena.mark();
    DbPostStatus st = db_postprocess_device(ctx, d_pred, batch, h, w, src_h, src_w, *cfg, out, hint);
    OAR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->arena.release_to(mark);
    if (db_postprocess_check(st, &hint) == 0) break;
    if (tries >= 3) OAR_FAIL(OAR_E_CAPACITY, "DB post-process component bound did not converge");
  }
  OAR_CUDA(cudaMemcpyAsync(counts, out.counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaMemcpyAsync(boxes, out.boxes, (size_t)batch * mc * 32, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaMemcpyAsync(scores, out.scores, (size_t)batch * mc * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_det_run(oar_model* det, const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                    const oar_det_config* cfg, float* boxes, float* scores, int32_t* counts) {
  API_TRY
  if (!det || det->kind != OAR_KIND_DET) OAR_FAIL(OAR_E_INVALID, "not a detection model");
  if (!cfg) OAR_FAIL(OAR_E_INVALID, "null config");
  if (n <= 0) return OAR_OK;  // DBModel::forward on an empty batch returns empty output, db.rs:288-293
  if (!images || !hs || !ws || !boxes || !scores || !counts) OAR_FAIL(OAR_E_INVALID, "null argument");
  oar_ctx* ctx = det->ctx;
  CallGuard guard(ctx);
  std::vector<DevImage> imgs(n);
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !images[i]) OAR_FAIL(OAR_E_INVALID, "image %d is empty", i);
    size_t bytes = (size_t)hs[i] * ws[i] * 3;
    uint8_t* d = ctx->arena.get<uint8_t>(bytes);
    OAR_CUDA(cudaMemcpyAsync(d, images[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
    imgs[i] = DevImage{d, hs[i], ws[i]};
  }
  DetResult res;
  run_detection(det, imgs, *cfg, n, res, false);
  const int mc = cfg->max_candidates;
  for (int i = 0; i < n; ++i) {
    counts[i] = (int32_t)res.scores[i].size();
    memcpy(boxes + (size_t)i * mc * 8, res.boxes[i].data(), res.boxes[i].size() * sizeof(float));
    memcpy(scores + (size_t)i * mc, res.scores[i].data(), res.scores[i].size() * sizeof(float));
  }
  API_CATCH
}

int32_t oar_sort_quad_boxes(float* boxes, int32_t n, int32_t* order) {
  API_TRY
  if (n < 0 || (n > 0 && !boxes)) OAR_FAIL(OAR_E_INVALID, "bad argument");
  std::vector<int> ord;
  sort_quads_host(boxes, n, ord);
  std::vector<float> tmp(boxes, boxes + (size_t)n * 8);
  for (int i = 0; i < n; ++i) {
    memcpy(boxes + (size_t)i * 8, tmp.data() + (size_t)ord[i] * 8, 8 * sizeof(float));
    if (order) order[i] = ord[i];
  }
  API_CATCH
}

int32_t oar_rotate_crop(oar_ctx* ctx, const uint8_t* image, int32_t h, int32_t w, const float* quads, int32_t n,
                        int32_t* out_w, int32_t* out_h, int32_t* status, uint8_t* out, size_t out_cap) {
  API_TRY
  require_device(ctx);
  if (n <= 0) return OAR_OK;
  if (!image || h <= 0 || w <= 0 || !quads || !out_w || !out_h || !status) OAR_FAIL(OAR_E_INVALID, "bad argument");
  CallGuard guard(ctx);
  size_t bytes = (size_t)h * w * 3;
  uint8_t* d_img = ctx->arena.get<uint8_t>(bytes);
  OAR_CUDA(cudaMemcpyAsync(d_img, image, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ImageRef ref{d_img, h, w};
  ImageRef* d_ref = to_device(ctx, &ref, 1);
  std::vector<CropPlan> plans(n);
  for (int i = 0; i < n; ++i) {
    memset(&plans[i], 0, sizeof(CropPlan));
    memcpy(plans[i].quad, quads + (size_t)i * 8, 8 * sizeof(float));
    plans[i].img = 0;
  }
  CropPlan* d_plans = to_device(ctx, plans.data(), plans.size());
  launch_crop_plan(ctx, d_plans, n, d_ref);
  CropPlan* h_plans = (CropPlan*)ctx->pinned_get(sizeof(CropPlan) * n);
  OAR_CUDA(cudaMemcpyAsync(h_plans, d_plans, sizeof(CropPlan) * n, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    status[i] = h_plans[i].status;
    out_w[i] = h_plans[i].status ? 0 : h_plans[i].ow;
    out_h[i] = h_plans[i].status ? 0 : h_plans[i].oh;
    h_plans[i].out_off = total * 3;
    if (!h_plans[i].status) total += (long long)h_plans[i].ow * h_plans[i].oh;
  }
  if (!out) return OAR_OK;
  if ((size_t)total * 3 > out_cap) OAR_FAIL(OAR_E_CAPACITY, "crops need %lld bytes, capacity %zu", total * 3, out_cap);
  if (total == 0) return OAR_OK;
  OAR_CUDA(cudaMemcpyAsync(d_plans, h_plans, sizeof(CropPlan) * n, cudaMemcpyHostToDevice, ctx->stream));
  uint8_t* pool = ctx->arena.get<uint8_t>((size_t)total * 3);
  launch_crop_warp(ctx, d_plans, n, d_ref, pool, total);
  OAR_CUDA(cudaMemcpyAsync(out, pool, (size_t)total * 3, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_crnn_preprocess(oar_ctx* ctx, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws,
                            int32_t n, float* out, size_t out_cap, int32_t* tensor_w) {
  API_TRY
  require_device(ctx);
  if (!tensor_w) OAR_FAIL(OAR_E_INVALID, "null argument");
  *tensor_w = 0;
  if (n <= 0) return OAR_OK;  // preprocess_refs on an empty batch yields a (0,0,0,0) tensor
  if (!crops || !hs || !ws) OAR_FAIL(OAR_E_INVALID, "null argument");
  float max_ratio = (float)REC_W / (float)REC_H;
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !crops[i]) OAR_FAIL(OAR_E_INVALID, "crop %d is empty", i);
    max_ratio = std::fmax(max_ratio, (float)ws[i] / (float)std::max(hs[i], 1));
  }
  const int tw = (int)std::min<int64_t>(f32_as_usize((float)REC_H * max_ratio), REC_MAX_W);
  *tensor_w = tw;
  if (!out) return OAR_OK;
  size_t total = (size_t)n * 3 * REC_H * tw;
  if (total > out_cap) OAR_FAIL(OAR_E_CAPACITY, "tensor needs %zu floats, capacity %zu", total, out_cap);
  CallGuard guard(ctx);
  std::vector<ResizeJob> jobs(n);
  std::vector<CrnnJob> cj(n);
  int max_sw = 0, max_dw = 0;
  for (int i = 0; i < n; ++i) {
    size_t bytes = (size_t)hs[i] * ws[i] * 3;
    uint8_t* d = ctx->arena.get<uint8_t>(bytes);
    OAR_CUDA(cudaMemcpyAsync(d, crops[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
    float ratio = (float)ws[i] / (float)hs[i];
    int rw = (int)std::min<int64_t>(f32_as_usize(std::ceil((float)REC_H * ratio)), tw);
    ResizeJob& j = jobs[i];
    j.src = d, j.sw = ws[i], j.sh = hs[i], j.dw = rw, j.dh = REC_H;
    j.tmp = ctx->arena.get<float>((size_t)REC_H * ws[i] * 3);
    j.dst = ctx->arena.get<uint8_t>((size_t)REC_H * std::max(rw, 1) * 3);
    cj[i].src = j.dst, cj[i].rw = rw;
    max_sw = std::max(max_sw, j.sw), max_dw = std::max(max_dw, rw);
  }
  ResizeJob* d_jobs = to_device(ctx, jobs.data(), jobs.size());
  CrnnJob* d_cj = to_device(ctx, cj.data(), cj.size());
  launch_resize_triangle(ctx, d_jobs, n, max_sw, max_dw, REC_H);
  float* d_out = ctx->arena.get<float>(total);
  launch_crnn_normalize(ctx, d_cj, n, REC_H, tw, d_out, 0);
  OAR_CUDA(cudaMemcpyAsync(out, d_out, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_ctc_decode(oar_ctx* ctx, const float* pred, int32_t b, int32_t t, int32_t v, int32_t n_chars,
                       int32_t* idx, float* prob, int32_t* labels, int32_t* cols, int32_t* lens, float* scores) {
  API_TRY
  require_device(ctx);
  if (b < 0 || t < 0 || v < 0) OAR_FAIL(OAR_E_INVALID, "negative dimension");
  // decode.rs:464-476: "no batch entries are returned when any dimension is zero" -- outputs untouched
  if (b == 0 || t == 0 || v == 0) return OAR_OK;
  if (!idx || !prob || !labels || !cols || !lens || !scores) OAR_FAIL(OAR_E_INVALID, "null argument");
  size_t rows = (size_t)b * t;
  if (!pred) OAR_FAIL(OAR_E_INVALID, "null argument");
  CallGuard guard(ctx);
  float* d_pred = ctx->arena.get<float>(rows * v);
  OAR_CUDA(cudaMemcpyAsync(d_pred, pred, rows * v * 4, cudaMemcpyHostToDevice, ctx->stream));
  int32_t* d_idx = ctx->arena.get<int32_t>(rows);
  float* d_prob = ctx->arena.get<float>(rows);
  int32_t* d_labels = ctx->arena.get<int32_t>(rows);
  int32_t* d_cols = ctx->arena.get<int32_t>(rows);
  int32_t* d_lens = ctx->arena.get<int32_t>(b);
  float* d_scores = ctx->arena.get<float>(b);
  launch_ctc_argmax(ctx, d_pred, (long long)rows, v, d_idx, d_prob);
  launch_ctc_decode(ctx, d_idx, d_prob, b, t, n_chars, d_labels, d_cols, d_lens, d_scores);
  cudaStream_t st = ctx->stream;
  OAR_CUDA(cudaMemcpyAsync(idx, d_idx, rows * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(prob, d_prob, rows * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(labels, d_labels, rows * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(cols, d_cols, rows * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(lens, d_lens, (size_t)b * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(scores, d_scores, (size_t)b * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaStreamSynchronize(st));
  API_CATCH
}

int32_t oar_rec_run(oar_model* rec, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                    int32_t n_chars, int32_t* labels, int32_t* cols, int32_t* lens, float* scores, int32_t t_cap,
                    int32_t* t_out) {
  return oar_rec_run_ex(rec, crops, hs, ws, n, 0, n_chars, labels, cols, lens, scores, t_cap, t_out);
}

int32_t oar_rec_run_ex(oar_model* rec, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                       int32_t crops_on_device, int32_t n_chars, int32_t* labels, int32_t* cols, int32_t* lens,
                       float* scores, int32_t t_cap, int32_t* t_out) {
  API_TRY
  if (!rec || rec->kind != OAR_KIND_REC) OAR_FAIL(OAR_E_INVALID, "not a recognition model");
  if (t_out) *t_out = 0;
  if (n <= 0) return OAR_OK;
  if (!crops || !hs || !ws || !labels || !cols || !lens || !scores) OAR_FAIL(OAR_E_INVALID, "null argument");
  oar_ctx* ctx = rec->ctx;
  CallGuard guard(ctx);
  std::vector<RecCrop> rc(n);
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !crops[i]) OAR_FAIL(OAR_E_INVALID, "crop %d is empty", i);
    total += ((size_t)hs[i] * ws[i] * 3 + 15) & ~(size_t)15;
  }
  if (crops_on_device) {
    for (int i = 0; i < n; ++i) rc[i] = RecCrop{crops[i], hs[i], ws[i]};
  } else {
    uint8_t* d = ctx->arena.get<uint8_t>(total);
    // crops in page-locked memory are copied from where they lie, neighbours in one transfer (a 512-crop batch is 24 MB:
    // staging it costs the host more than the recogniser costs the GPU); pageable crops go through one pinned staging
    // buffer and one copy (a copy per crop from pageable memory costs ~10 us each)
    bool pinned = true;
    for (int i = 0; i < n && pinned; ++i) {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, crops[i]) != cudaSuccess) {
        cudaGetLastError();
        pinned = false;
      } else {
        pinned = at.type == cudaMemoryTypeHost;
      }
    }
    if (pinned) {
      size_t off = 0, run_off = 0, run_bytes = 0;
      const uint8_t* run_src = nullptr;
      for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)hs[i] * ws[i] * 3, padded = (bytes + 15) & ~(size_t)15;
        rc[i] = RecCrop{d + off, hs[i], ws[i]};
        if (run_src && crops[i] == run_src + run_bytes && off == run_off + run_bytes) {
          run_bytes += bytes;  // continues the run on both sides (only possible while sizes are multiples of 16)
        } else {
          if (run_src) OAR_CUDA(cudaMemcpyAsync(d + run_off, run_src, run_bytes, cudaMemcpyHostToDevice, ctx->stream));
          run_src = crops[i], run_off = off, run_bytes = bytes;
        }
        off += padded;
      }
      if (run_src) OAR_CUDA(cudaMemcpyAsync(d + run_off, run_src, run_bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      uint8_t* h = (uint8_t*)ctx->pinned_get(total);
      size_t off = 0;
      for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)hs[i] * ws[i] * 3;
        memcpy(h + off, crops[i], bytes);
        rc[i] = RecCrop{d + off, hs[i], ws[i]};
        off += (bytes + 15) & ~(size_t)15;
      }
      OAR_CUDA(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  RecBatchOut out;
  launch_rec_batch(rec, rc.data(), n, n_chars, out);
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (out.T > t_cap) OAR_FAIL(OAR_E_CAPACITY, "sequence length %d exceeds capacity %d", out.T, t_cap);
  if (t_out) *t_out = out.T;
  for (int i = 0; i < n; ++i) {
    memcpy(labels + (size_t)i * t_cap, out.h_labels + (size_t)i * out.T, (size_t)out.T * 4);
    memcpy(cols + (size_t)i * t_cap, out.h_cols + (size_t)i * out.T, (size_t)out.T * 4);
    lens[i] = out.h_lens[i];
    scores[i] = out.h_scores[i];
  }
  API_CATCH
}

int32_t oar_cls_run(oar_model* cls, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                    int32_t input_h, int32_t input_w, int32_t* class_ids, float* scores, float* probs,
                    size_t probs_cap, int32_t* n_classes) {
  API_TRY
  if (!cls || cls->kind != OAR_KIND_CLS) OAR_FAIL(OAR_E_INVALID, "not a classification model");
  if (n_classes) *n_classes = 0;
  if (input_h <= 0 || input_w <= 0) OAR_FAIL(OAR_E_INVALID, "input shape must be positive");
  if (n <= 0) return OAR_OK;  // the adapter returns an empty TextLineOrientationOutput
  if (!crops || !hs || !ws || !class_ids || !scores) OAR_FAIL(OAR_E_INVALID, "null argument");
  oar_ctx* ctx = cls->ctx;
  CallGuard guard(ctx);
  std::vector<RecCrop> rc(n);
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !crops[i]) OAR_FAIL(OAR_E_INVALID, "crop %d is empty", i);
    size_t bytes = (size_t)hs[i] * ws[i] * 3;
    uint8_t* d = ctx->arena.get<uint8_t>(bytes);
    OAR_CUDA(cudaMemcpyAsync(d, crops[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc[i] = RecCrop{d, hs[i], ws[i]};
  }
  int32_t* d_ids = ctx->arena.get<int32_t>(n);
  float* d_sc = ctx->arena.get<float>(n);
  // class count is a property of the graph: the head op's output width
  int C = 0;
  for (const OpRec& op : cls->ops)
    if (op.type == OP_CTC_HEAD) C = op.p[1];
  if (C <= 0) OAR_FAIL(OAR_E_MODEL, "classifier graph has no Linear+Softmax head");
  float* d_probs = probs ? ctx->arena.get<float>((size_t)n * C) : nullptr;
  if (probs && (size_t)n * C > probs_cap)
    OAR_FAIL(OAR_E_CAPACITY, "probabilities need %zu floats, capacity %zu", (size_t)n * C, probs_cap);
  for (int s0 = 0; s0 < n; s0 += CLS_CHUNK) {
    int m = std::min(CLS_CHUNK, n - s0);
    int got = launch_cls_batch(cls, rc.data() + s0, m, input_h, input_w, d_ids + s0, d_sc + s0,
                               d_probs ? d_probs + (size_t)s0 * C : nullptr, (size_t)m * C);
    if (got != C) OAR_FAIL(OAR_E_MODEL, "classifier produced %d classes, graph declares %d", got, C);
  }
  cudaStream_t st = ctx->stream;
  OAR_CUDA(cudaMemcpyAsync(class_ids, d_ids, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaMemcpyAsync(scores, d_sc, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (probs) OAR_CUDA(cudaMemcpyAsync(probs, d_probs, (size_t)n * C * 4, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaStreamSynchronize(st));
  if (n_classes) *n_classes = C;
  API_CATCH
}

int32_t oar_rotate180(oar_ctx* ctx, const uint8_t* image, int32_t h, int32_t w, uint8_t* out) {
  API_TRY
  require_device(ctx);
  if (h < 0 || w < 0) OAR_FAIL(OAR_E_INVALID, "negative dimension");
  size_t bytes = (size_t)h * w * 3;
  if (bytes == 0) return OAR_OK;
  if (!image || !out) OAR_FAIL(OAR_E_INVALID, "null buffer");
  CallGuard guard(ctx);
  uint8_t* d = ctx->arena.get<uint8_t>(bytes);
  OAR_CUDA(cudaMemcpyAsync(d, image, bytes, cudaMemcpyHostToDevice, ctx->stream));
  Rot180Job job{d, h * w};
  Rot180Job* d_job = to_device(ctx, &job, 1);
  launch_rotate180(ctx, d_job, 1, h * w, nullptr);
  OAR_CUDA(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

// ---------------------------------------------------------------------------
// OAROCR::predict in stages (ocr.rs:518-659).  One PageStage per context: pages -> HBM, detection + reading-order
// sort, crop (+ optional line orientation).  Recognition works on a flat list of crop references in global
// (image, detection) order, which may point into the crop pools of several contexts (oar_pipeline_run_multi).
// ---------------------------------------------------------------------------
namespace {

struct PageStage {
  oar_ctx* ctx = nullptr;
  int n = 0;  // images of this stage
  std::vector<DevImage> imgs;
  std::vector<cudaEvent_t> ready;                // per image: its upload has landed (null: already resident)
  std::vector<std::vector<float>> sorted_boxes;  // per image: count * 8 floats, reading order
  std::vector<int> box_first;                    // [n + 1] prefix sums of the box counts
  int n_boxes = 0;
  CropPlan* h_plans = nullptr;  // pinned, [n_boxes]: status, dims, wh_ratio, pool offset
  uint8_t* pool = nullptr;      // device crop pool
  std::vector<int> valid_of;    // box -> index among the valid crops (orientation stage)
  int32_t* h_cls_ids = nullptr; // pinned, per valid crop (after the final synchronise)
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  float ms_det = 0, ms_post = 0;
  cudaEvent_t ev[6] = {};
};

// pages into HBM.  Host pages go through the context's copy stream, one event per page: the launch stream waits for
// exactly the pages a detection sub-group needs, so the upload of the later pages overlaps the network of the earlier.
void stage_upload(PageStage& S, const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int n,
                  int on_device) {
  oar_ctx* ctx = S.ctx;
  S.n = n;
  S.imgs.resize(n);
  S.ready.assign(n, nullptr);
  static const bool no_copy_stream = getenv("OAR_DBG_ONE_STREAM") != nullptr;
  const bool async = !on_device && !no_copy_stream && ctx->stream_copy && n > 1;
  cudaStream_t cs = async ? ctx->stream_copy : ctx->stream;
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !images[i]) OAR_FAIL(OAR_E_INVALID, "image %d is empty", i);
    if (on_device) {
      S.imgs[i] = DevImage{images[i], hs[i], ws[i]};
    } else {
      size_t bytes = (size_t)hs[i] * ws[i] * 3;
      uint8_t* d = ctx->arena.get<uint8_t>(bytes);
      OAR_CUDA(cudaMemcpyAsync(d, images[i], bytes, cudaMemcpyHostToDevice, cs));
      if (async) {
        S.ready[i] = ctx->next_event();
        OAR_CUDA(cudaEventRecord(S.ready[i], cs));
      }
      S.imgs[i] = DevImage{d, hs[i], ws[i]};
      S.h2d_bytes += (int64_t)bytes;
    }
  }
}

// detection (chunks of image_batch_size) + sort_quad_boxes per image
void stage_detect(PageStage& S, oar_model* det, const oar_pipeline_config* cfg) {
  DetResult dres;
  run_detection(det, S.imgs, cfg->det, cfg->image_batch_size, dres, true, S.ready.empty() ? nullptr : &S.ready);
  S.ms_det = dres.ms_net, S.ms_post = dres.ms_post;
  const int n = S.n;
  S.sorted_boxes.assign(n, {});
  S.box_first.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    int cnt = (int)dres.scores[i].size();
    std::vector<int> ord;
    sort_quads_host(dres.boxes[i].data(), cnt, ord);
    S.sorted_boxes[i].resize((size_t)cnt * 8);
    for (int k = 0; k < cnt; ++k)
      memcpy(&S.sorted_boxes[i][(size_t)k * 8], &dres.boxes[i][(size_t)ord[k] * 8], 8 * sizeof(float));
    S.box_first[i + 1] = S.box_first[i] + cnt;
    S.d2h_bytes += (int64_t)cnt * 36 + 4;
  }
  S.n_boxes = S.box_first[n];
}

// caller-supplied boxes (oar_crop_rec_run): box k of the call belongs to image img_index[k]; order within an image kept
void stage_set_boxes(PageStage& S, const float* boxes, const int32_t* img_index, int n_boxes, std::vector<int>& slot_of) {
  const int n = S.n;
  S.sorted_boxes.assign(n, {});
  S.box_first.assign(n + 1, 0);
  for (int k = 0; k < n_boxes; ++k) {
    if (img_index[k] < 0 || img_index[k] >= n) OAR_FAIL(OAR_E_INVALID, "box %d names image %d of %d", k, img_index[k], n);
    ++S.box_first[img_index[k] + 1];
  }
  for (int i = 0; i < n; ++i) S.box_first[i + 1] += S.box_first[i];
  std::vector<int> fill(S.box_first.begin(), S.box_first.end() - 1);
  for (int i = 0; i < n; ++i) S.sorted_boxes[i].resize((size_t)(S.box_first[i + 1] - S.box_first[i]) * 8);
  slot_of.assign(n_boxes, 0);
  for (int k = 0; k < n_boxes; ++k) {
    const int i = img_index[k], pos = fill[i]++;
    memcpy(&S.sorted_boxes[i][(size_t)(pos - S.box_first[i]) * 8], boxes + (size_t)k * 8, 8 * sizeof(float));
    slot_of[k] = pos;
  }
  S.n_boxes = n_boxes;
}

// get_rotate_crop_image for every box: plan on device, sizes back to the host, one warp launch for all
void stage_crop(PageStage& S) {
  oar_ctx* ctx = S.ctx;
  cudaStream_t st = ctx->stream;
  for (cudaEvent_t e : S.ready)  // every page must have landed (a no-op after detection, which waited group by group)
    if (e) OAR_CUDA(cudaStreamWaitEvent(st, e, 0));
  const int n = S.n, n_boxes = S.n_boxes;
  S.pool = nullptr;
  S.h_plans = nullptr;
  if (n_boxes <= 0) return;
  std::vector<CropPlan> plans(n_boxes);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < S.box_first[i + 1] - S.box_first[i]; ++k) {
      CropPlan& p = plans[S.box_first[i] + k];
      memset(&p, 0, sizeof(CropPlan));
      memcpy(p.quad, &S.sorted_boxes[i][(size_t)k * 8], 8 * sizeof(float));
      p.img = i;
    }
  std::vector<ImageRef> refs(n);
  for (int i = 0; i < n; ++i) refs[i] = ImageRef{S.imgs[i].p, S.imgs[i].h, S.imgs[i].w};
  ImageRef* d_refs = to_device(ctx, refs.data(), refs.size());
  CropPlan* d_plans = to_device(ctx, plans.data(), plans.size());
  launch_crop_plan(ctx, d_plans, n_boxes, d_refs);
  S.h_plans = (CropPlan*)ctx->pinned_get(sizeof(CropPlan) * n_boxes);
  OAR_CUDA(cudaMemcpyAsync(S.h_plans, d_plans, sizeof(CropPlan) * n_boxes, cudaMemcpyDeviceToHost, st));
  OAR_CUDA(cudaStreamSynchronize(st));
  long long total_px = 0;
  for (int i = 0; i < n_boxes; ++i) {
    S.h_plans[i].out_off = total_px * 3;
    if (S.h_plans[i].status == 0) total_px += (long long)S.h_plans[i].ow * S.h_plans[i].oh;
  }
  if (total_px > 0) {
    OAR_CUDA(cudaMemcpyAsync(d_plans, S.h_plans, sizeof(CropPlan) * n_boxes, cudaMemcpyHostToDevice, st));
    S.pool = ctx->arena.get<uint8_t>((size_t)total_px * 3);
    launch_crop_warp(ctx, d_plans, n_boxes, d_refs, S.pool, total_px);
  }
}

// Line orientation (ocr.rs:615, 755-792): classify every crop, rotate class-1 crops by 180 degrees in the pool.
// The reference calls the adapter once per image; the classifier treats every crop independently (fixed 80 x 160
// input, per-sample pooling), so one pass over all crops in chunks gives the same classes.  No host round trip:
// the rotation kernel reads the class ids on the device; the host reads them after the final synchronise.
void stage_orient(PageStage& S, oar_model* cls) {
  oar_ctx* ctx = S.ctx;
  S.valid_of.assign(S.n_boxes, -1);
  S.h_cls_ids = nullptr;
  if (!cls || !S.pool) return;
  std::vector<int> valid;
  for (int i = 0; i < S.n_boxes; ++i)
    if (S.h_plans[i].status == 0) S.valid_of[i] = (int)valid.size(), valid.push_back(i);
  const int nv = (int)valid.size();
  int32_t* d_ids = ctx->arena.get<int32_t>(nv);
  float* d_sc = ctx->arena.get<float>(nv);
  std::vector<Rot180Job> rj(nv);
  int max_npix = 0;
  for (int s0 = 0; s0 < nv; s0 += CLS_CHUNK) {
    int m = std::min(CLS_CHUNK, nv - s0);
    std::vector<RecCrop> rc(m);
    for (int k = 0; k < m; ++k) {
      const CropPlan& p = S.h_plans[valid[s0 + k]];
      rc[k] = RecCrop{S.pool + p.out_off, p.oh, p.ow};
      rj[s0 + k] = Rot180Job{S.pool + p.out_off, p.oh * p.ow};
      max_npix = std::max(max_npix, p.oh * p.ow);
    }
    launch_cls_batch(cls, rc.data(), m, CLS_H, CLS_W, d_ids + s0, d_sc + s0, nullptr, 0);
  }
  Rot180Job* d_rj = to_device(ctx, rj.data(), rj.size());
  launch_rotate180(ctx, d_rj, nv, max_npix, d_ids);
  S.h_cls_ids = (int32_t*)ctx->pinned_get((size_t)nv * sizeof(int32_t));
  OAR_CUDA(cudaMemcpyAsync(S.h_cls_ids, d_ids, (size_t)nv * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  S.d2h_bytes += (int64_t)nv * 4;
}

// one crop of the global pool: where its pixels live and what recognize_global needs to know about it
struct CropRef {
  int stage;  // owning PageStage (= context)
  int box;    // box index inside that stage
  const uint8_t* p;
  int h, w;
  float ratio;
};
// recognize_global (ocr.rs:802-897): waves of at most MAX_POOLED_CROPS crops in pool order, each stably sorted by
// wh_ratio and cut into chunks of region_batch_size.  chunk = [first, first + n) of `order`.
struct RecChunk {
  size_t first;
  int n;
  RecBatchOut out;
  int stage = 0;  // the context that recognises it
};
struct RecPlan {
  std::vector<int> order;  // indices into the CropRef list, wave by wave in wh-ratio order
  std::vector<RecChunk> chunks;
};
void plan_recognition(const std::vector<CropRef>& refs, int region_batch_size, RecPlan& plan) {
  plan.order.clear();
  plan.chunks.clear();
  plan.order.reserve(refs.size());
  for (size_t w0 = 0; w0 < refs.size(); w0 += MAX_POOLED_CROPS) {
    const size_t w1 = std::min(refs.size(), w0 + (size_t)MAX_POOLED_CROPS);
    const size_t base = plan.order.size();
    for (size_t i = w0; i < w1; ++i) plan.order.push_back((int)i);
    std::stable_sort(plan.order.begin() + base, plan.order.end(),
                     [&](int a, int b) { return refs[a].ratio < refs[b].ratio; });
    for (size_t s0 = base; s0 < plan.order.size(); s0 += region_batch_size) {
      RecChunk c;
      c.first = s0;
      c.n = (int)std::min<size_t>(region_batch_size, plan.order.size() - s0);
      plan.chunks.push_back(c);
    }
  }
}

// per crop reference, after recognition: where its results are
struct RecResult {
  int len = -1;  // -1: not recognised
  const int32_t* labels = nullptr;
  const int32_t* cols = nullptr;
  int T = 0;
  float score = 0.0f, chunk_max_ratio = 0.0f;
};
void collect_results(const std::vector<CropRef>& refs, const RecPlan& plan, float rec_score_thresh,
                     std::vector<RecResult>& res, int64_t* d2h_bytes) {
  res.assign(refs.size(), RecResult{});
  const float base_rec_ratio = (float)REC_W / (float)REC_H;  // DEFAULT_REC_IMAGE_SHAPE, ocr.rs:817
  for (const RecChunk& ch : plan.chunks) {
    float chunk_max = base_rec_ratio;
    for (int k = 0; k < ch.n; ++k) chunk_max = std::fmax(chunk_max, refs[plan.order[ch.first + k]].ratio);
    for (int k = 0; k < ch.n; ++k) {
      RecResult& r = res[plan.order[ch.first + k]];
      r.cols = ch.out.h_cols + (size_t)k * ch.out.T;
      r.labels = ch.out.h_labels + (size_t)k * ch.out.T;
      r.T = ch.out.T;
      r.chunk_max_ratio = chunk_max;
      r.score = ch.out.h_scores[k];
      // TextRecognitionAdapter::execute: score below the threshold keeps the slot with empty text
      r.len = r.score >= rec_score_thresh ? ch.out.h_lens[k] : 0;
    }
    if (d2h_bytes) *d2h_bytes += (int64_t)ch.n * (ch.out.T * 8 + 8);
  }
}

// crop references of one stage in (image, detection) order; failed crops are skipped (processors.rs:104-106)
void append_refs(const PageStage& S, int stage_index, std::vector<CropRef>& refs, std::vector<int>* ref_of_box) {
  if (ref_of_box) ref_of_box->assign(S.n_boxes, -1);
  for (int i = 0; i < S.n_boxes; ++i) {
    const CropPlan& p = S.h_plans[i];
    if (p.status != 0) continue;
    if (ref_of_box) (*ref_of_box)[i] = (int)refs.size();
    refs.push_back(CropRef{stage_index, i, S.pool + p.out_off, p.oh, p.ow, p.wh_ratio});
  }
}

// scatter to per-image, detection-index order (ocr.rs:879-892, 637-656); images [img0, img0 + S.n) of the result
void scatter_stage(const PageStage& S, int img0, const std::vector<int>& ref_of_box, const std::vector<RecResult>& res,
                   oar_ocr_result* out, int& r, long long& nl) {
  for (int i = 0; i < S.n; ++i) {
    out->region_off[img0 + i] = r;
    for (int k = 0; k < S.box_first[i + 1] - S.box_first[i]; ++k) {
      const int box = S.box_first[i] + k;
      const int ref = ref_of_box[box];
      if (ref < 0 || res[ref].len < 0) continue;
      const RecResult& rr = res[ref];
      if (r >= out->cap_regions || nl + rr.len > out->cap_labels)
        OAR_FAIL(OAR_E_CAPACITY, "result buffers too small (regions %d, labels %d)", out->cap_regions, out->cap_labels);
      if (out->boxes) memcpy(out->boxes + (size_t)r * 8, &S.sorted_boxes[i][(size_t)k * 8], 8 * sizeof(float));
      if (out->scores) out->scores[r] = rr.score;
      if (out->det_index) out->det_index[r] = k;
      if (out->label_off) out->label_off[r] = (int32_t)nl;
      if (out->labels && rr.len > 0) memcpy(out->labels + nl, rr.labels, (size_t)rr.len * 4);
      if (out->cols && rr.len > 0) memcpy(out->cols + nl, rr.cols, (size_t)rr.len * 4);
      if (out->seq_len) out->seq_len[r] = rr.T;
      if (out->wh_ratio) out->wh_ratio[r] = S.h_plans[box].wh_ratio;
      if (out->max_wh_ratio) out->max_wh_ratio[r] = rr.chunk_max_ratio;
      if (out->line_angle) out->line_angle[r] = S.h_cls_ids ? (float)S.h_cls_ids[S.valid_of[box]] * 180.0f : -1.0f;
      nl += rr.len;
      ++r;
    }
  }
}

// Recognition batches of one context.  The batches are independent, so they alternate between the context's two launch
// lanes (stream + arena each): the latency-bound tail of one batch (SVTR neck: ~30 small launches, CTC combine, decode)
// overlaps the bandwidth-bound backbone of the next.  A lane's arena is released batch by batch in its own stream order;
// the crop pool and the job tables the batches read were written on the main lane before the fork event.
void recognize_chunks(oar_model* rec, const std::vector<CropRef>& refs, RecPlan& plan, int n_chars, int only_stage = -1) {
  oar_ctx* ctx = rec->ctx;
  // (Round 2 found run-to-run differences with two batches in flight and bisected them -- tools/det_diff.py,
  // profiles/r2_two_lane_bisect.txt -- to lcblock_tc handing its TMA stage back before the loads from it had completed,
  // fused_tc.cu: dep_zero.  Fixed there; OAR_REC_LANES=1 keeps the single-lane form for A/B runs.)
  static const bool one_lane = getenv("OAR_DBG_ONE_STREAM") != nullptr ||
                               (getenv("OAR_REC_LANES") && atoi(getenv("OAR_REC_LANES")) < 2);
  int mine = 0;
  for (const RecChunk& ch : plan.chunks) mine += (only_stage < 0 || ch.stage == only_stage) ? 1 : 0;
  const bool two = !one_lane && !ctx->profile && mine >= 2 && ctx->stream_aux;
  cudaStream_t main_stream = ctx->stream;
  bool on_aux = false;
  auto lane = [&](bool aux) {
    if (aux == on_aux) return;
    std::swap(ctx->stream, ctx->stream_aux);
    std::swap(ctx->arena, ctx->arena_aux);
    on_aux = aux;
  };
  if (two) {
    cudaEvent_t fork = ctx->next_event();
    OAR_CUDA(cudaEventRecord(fork, main_stream));
    OAR_CUDA(cudaStreamWaitEvent(ctx->stream_aux, fork, 0));
  }
  std::vector<RecCrop> rc;
  int i = 0;
  try {
    for (RecChunk& ch : plan.chunks) {
      if (only_stage >= 0 && ch.stage != only_stage) continue;
      lane(two && (i++ & 1));
      rc.resize(ch.n);
      for (int k = 0; k < ch.n; ++k) {
        const CropRef& c = refs[plan.order[ch.first + k]];
        rc[k] = RecCrop{c.p, c.h, c.w};
      }
      launch_rec_batch(rec, rc.data(), ch.n, n_chars, ch.out);
      static const bool serial = getenv("OAR_DBG_LANE_SERIAL") != nullptr;  // bisecting aid: two lanes, no overlap
      if (two && serial) {
        cudaEvent_t done = ctx->next_event();
        OAR_CUDA(cudaEventRecord(done, ctx->stream));
        OAR_CUDA(cudaStreamWaitEvent(ctx->stream_aux, done, 0));
      }
    }
  } catch (...) {
    lane(false);
    throw;
  }
  lane(false);
  if (two) {
    cudaEvent_t join = ctx->next_event();
    OAR_CUDA(cudaEventRecord(join, ctx->stream_aux));
    OAR_CUDA(cudaStreamWaitEvent(main_stream, join, 0));
  }
}

void check_pipeline_args(oar_model* det, oar_model* rec, oar_model* cls, const uint8_t* const* images, const int32_t* hs,
                         const int32_t* ws, int32_t n, const oar_pipeline_config* cfg, oar_ocr_result* out) {
  if (!det || det->kind != OAR_KIND_DET || !rec || rec->kind != OAR_KIND_REC)
    OAR_FAIL(OAR_E_INVALID, "pipeline needs one detection and one recognition model");
  if (det->ctx != rec->ctx) OAR_FAIL(OAR_E_INVALID, "both models must live on the same context");
  if (cls && (cls->kind != OAR_KIND_CLS || cls->ctx != det->ctx))
    OAR_FAIL(OAR_E_INVALID, "the line-orientation model must be a classifier on the same context");
  if (!cfg || !out) OAR_FAIL(OAR_E_INVALID, "null argument");
  // OAROCR::predict rejects an empty image list (ocr.rs:525-532)
  if (n <= 0 || !images || !hs || !ws) OAR_FAIL(OAR_E_INVALID, "images: expected non-empty slice, got empty slice");
  if (cfg->image_batch_size <= 0 || cfg->region_batch_size <= 0)
    OAR_FAIL(OAR_E_INVALID, "batch sizes must be positive");  // ocr.rs:1168-1195
  if (!out->region_off) OAR_FAIL(OAR_E_INVALID, "null result buffers");
}

}  // namespace

int32_t oar_pipeline_run(oar_model* det, oar_model* rec, const uint8_t* const* images, const int32_t* hs,
                         const int32_t* ws, int32_t n, int32_t images_on_device, const oar_pipeline_config* cfg,
                         oar_ocr_result* out) {
  return oar_pipeline_run_cls(det, rec, nullptr, images, hs, ws, n, images_on_device, cfg, out);
}

namespace {
// everything after the pages are in HBM (S.imgs filled, ev[0] / ev[1] recorded by the caller)
void pipeline_after_upload(PageStage& S, oar_model* det, oar_model* rec, oar_model* cls, const oar_pipeline_config* cfg,
                           oar_ocr_result* out) {
  oar_ctx* ctx = S.ctx;
  cudaStream_t st = ctx->stream;
  const int n = S.n;
  stage_detect(S, det, cfg);
  cudaEventRecord(S.ev[2], st);
  stage_crop(S);
  cudaEventRecord(S.ev[3], st);
  stage_orient(S, cls);
  cudaEventRecord(S.ev[5], st);

  // ---- recognition: pool crops across images, flush at MAX_POOLED_CROPS, sort by wh_ratio, chunk
  std::vector<CropRef> refs;
  std::vector<int> ref_of_box;
  if (S.n_boxes > 0) append_refs(S, 0, refs, &ref_of_box);
  RecPlan plan;
  plan_recognition(refs, cfg->region_batch_size, plan);
  recognize_chunks(rec, refs, plan, cfg->n_chars);
  cudaEventRecord(S.ev[4], st);
  OAR_CUDA(cudaStreamSynchronize(st));

  std::vector<RecResult> res;
  int64_t d2h = 0;
  collect_results(refs, plan, cfg->rec_score_thresh, res, &d2h);
  int r = 0;
  long long nl = 0;
  scatter_stage(S, 0, ref_of_box, res, out, r, nl);
  out->region_off[n] = r;
  if (out->label_off) out->label_off[r] = (int32_t)nl;
  out->ms_det = S.ms_det, out->ms_post = S.ms_post;
  out->h2d_bytes = S.h2d_bytes, out->d2h_bytes = S.d2h_bytes + d2h;
  cudaEventElapsedTime(&out->ms_h2d, S.ev[0], S.ev[1]);
  cudaEventElapsedTime(&out->ms_crop, S.ev[2], S.ev[3]);
  cudaEventElapsedTime(&out->ms_rec, S.ev[3], S.ev[4]);
  cudaEventElapsedTime(&out->ms_cls, S.ev[3], S.ev[5]);
  cudaEventElapsedTime(&out->ms_total, S.ev[0], S.ev[4]);
}
}  // namespace

int32_t oar_pipeline_run_cls(oar_model* det, oar_model* rec, oar_model* cls, const uint8_t* const* images,
                             const int32_t* hs, const int32_t* ws, int32_t n, int32_t images_on_device,
                             const oar_pipeline_config* cfg, oar_ocr_result* out) {
  API_TRY
  check_pipeline_args(det, rec, cls, images, hs, ws, n, cfg, out);
  oar_ctx* ctx = det->ctx;
  CallGuard guard(ctx);
  PageStage S;
  S.ctx = ctx;
  for (auto& e : S.ev) e = ctx->next_event();
  out->ms_h2d = out->ms_det = out->ms_post = out->ms_crop = out->ms_rec = out->ms_total = out->ms_cls = 0.0f;
  out->h2d_bytes = out->d2h_bytes = 0;
  cudaEventRecord(S.ev[0], ctx->stream);
  stage_upload(S, images, hs, ws, n, images_on_device);
  cudaEventRecord(S.ev[1], ctx->stream);
  pipeline_after_upload(S, det, rec, cls, cfg, out);
  API_CATCH
}

// OAROCR::predict on ENCODED pages: JPEG bytes -> nvJPEG -> HBM -> the same pipeline.  ms_h2d
// covers upload + decode; out_hs / out_ws (may be NULL) receive the decoded page sizes.
int32_t oar_pipeline_run_encoded(oar_model* det, oar_model* rec, const uint8_t* const* jpegs, const size_t* lens, int32_t n,
                                 const oar_pipeline_config* cfg, oar_ocr_result* out, int32_t* out_hs, int32_t* out_ws) {
  API_TRY
  if (n <= 0 || !jpegs || !lens) OAR_FAIL(OAR_E_INVALID, "images: expected non-empty slice, got empty slice");
  std::vector<int32_t> hs(n, 1), ws(n, 1);
  check_pipeline_args(det, rec, nullptr, jpegs, hs.data(), ws.data(), n, cfg, out);
  oar_ctx* ctx = det->ctx;
  CallGuard guard(ctx);
  PageStage S;
  S.ctx = ctx;
  for (auto& e : S.ev) e = ctx->next_event();
  out->ms_h2d = out->ms_det = out->ms_post = out->ms_crop = out->ms_rec = out->ms_total = out->ms_cls = 0.0f;
  out->h2d_bytes = out->d2h_bytes = 0;
  cudaEventRecord(S.ev[0], ctx->stream);
  std::vector<const uint8_t*> ptrs(n);
  decode_jpegs_to_device(ctx, jpegs, lens, n, ptrs.data(), hs.data(), ws.data());
  S.n = n;
  S.imgs.resize(n);
  for (int i = 0; i < n; ++i) {
    S.imgs[i] = DevImage{ptrs[i], hs[i], ws[i]};
    S.h2d_bytes += (int64_t)lens[i];
    if (out_hs) out_hs[i] = hs[i];
    if (out_ws) out_ws[i] = ws[i];
  }
  cudaEventRecord(S.ev[1], ctx->stream);
  pipeline_after_upload(S, det, rec, nullptr, cfg, out);
  API_CATCH
}

// load_image (oar-ocr-core/src/core/utils/image.rs:88) for a JPEG stream, decoded on the device: RGB8 pixels back to the
// host (parity checks of the ingest path; the pipeline itself never brings them back).  rgb == NULL: only the size.
int32_t oar_decode_jpeg(oar_ctx* ctx, const uint8_t* jpeg, size_t len, uint8_t* rgb, size_t rgb_cap, int32_t* h, int32_t* w) {
  API_TRY
  require_device(ctx);
  if (!jpeg || !h || !w) OAR_FAIL(OAR_E_INVALID, "null argument");
  CallGuard guard(ctx);
  const uint8_t* p = nullptr;
  decode_jpegs_to_device(ctx, &jpeg, &len, 1, &p, h, w);
  if (rgb) {
    const size_t bytes = (size_t)*h * *w * 3;
    if (bytes > rgb_cap) OAR_FAIL(OAR_E_CAPACITY, "decoded page needs %zu bytes, capacity %zu", bytes, rgb_cap);
    OAR_CUDA(cudaMemcpyAsync(rgb, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

// The second half of OAROCR::predict with caller-supplied boxes (oar_crop_rec_run): every box is cropped
// from its page (get_rotate_crop_image, transform.rs:76-502) and the crops are recognised as recognize_global batches
// them (ocr.rs:802-897).  Per box k: status[k] = 0 ok / 1 crop failed (skipped), lens[k], labels/cols [k][t_cap],
// scores[k], seq_len[k] = T of its batch.
int32_t oar_crop_rec_run(oar_model* rec, const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                         int32_t images_on_device, const float* boxes, const int32_t* img_index, int32_t n_boxes,
                         int32_t region_batch_size, int32_t n_chars, float rec_score_thresh, int32_t* status,
                         int32_t* labels, int32_t* cols, int32_t* lens, float* scores, int32_t* seq_len, int32_t t_cap) {
  API_TRY
  if (!rec || rec->kind != OAR_KIND_REC) OAR_FAIL(OAR_E_INVALID, "not a recognition model");
  if (n <= 0 || !images || !hs || !ws) OAR_FAIL(OAR_E_INVALID, "images: expected non-empty slice, got empty slice");
  if (region_batch_size <= 0) OAR_FAIL(OAR_E_INVALID, "batch sizes must be positive");
  if (n_boxes <= 0) return OAR_OK;
  if (!boxes || !img_index || !status || !labels || !cols || !lens || !scores) OAR_FAIL(OAR_E_INVALID, "null argument");
  oar_ctx* ctx = rec->ctx;
  CallGuard guard(ctx);
  PageStage S;
  S.ctx = ctx;
  stage_upload(S, images, hs, ws, n, images_on_device);
  std::vector<int> slot_of;
  stage_set_boxes(S, boxes, img_index, n_boxes, slot_of);
  stage_crop(S);
  stage_orient(S, nullptr);
  std::vector<CropRef> refs;
  std::vector<int> ref_of_box;
  append_refs(S, 0, refs, &ref_of_box);
  RecPlan plan;
  plan_recognition(refs, region_batch_size, plan);
  recognize_chunks(rec, refs, plan, n_chars);
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<RecResult> res;
  collect_results(refs, plan, rec_score_thresh, res, nullptr);
  for (int k = 0; k < n_boxes; ++k) {
    const int ref = ref_of_box[slot_of[k]];
    status[k] = ref < 0 ? 1 : 0;
    lens[k] = 0, scores[k] = 0.0f;
    if (seq_len) seq_len[k] = 0;
    if (ref < 0) continue;
    const RecResult& rr = res[ref];
    if (rr.T > t_cap) OAR_FAIL(OAR_E_CAPACITY, "sequence length %d exceeds capacity %d", rr.T, t_cap);
    lens[k] = rr.len, scores[k] = rr.score;
    if (seq_len) seq_len[k] = rr.T;
    if (rr.len > 0) {
      memcpy(labels + (size_t)k * t_cap, rr.labels, (size_t)rr.len * 4);
      memcpy(cols + (size_t)k * t_cap, rr.cols, (size_t)rr.len * 4);
    }
  }
  API_CATCH
}

// OAROCR::predict over several GPUs inside one process (one host thread + CUDA context per GPU), equal to
// ONE un-sharded predict(): pages are dealt to the contexts in contiguous blocks; every context uploads, detects and
// crops its block; the host pools ALL crops in (image, detection) order and plans recognize_global once; whole chunks
// are dealt round-robin and a context fetches the crops that live in another context's pool over NVLink
// (cudaMemcpyPeerAsync) before recognising its chunks.  No collective: sizes and results travel through host memory.
int32_t oar_pipeline_run_multi(oar_model* const* dets, oar_model* const* recs, int32_t n_ctx,
                               const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                               const oar_pipeline_config* cfg, oar_ocr_result* out) {
  API_TRY
  if (n_ctx <= 0 || !dets || !recs) OAR_FAIL(OAR_E_INVALID, "need at least one context");
  for (int g = 0; g < n_ctx; ++g) {
    check_pipeline_args(dets[g], recs[g], nullptr, images, hs, ws, n, cfg, out);
    for (int h = 0; h < g; ++h)
      if (dets[h]->ctx == dets[g]->ctx) OAR_FAIL(OAR_E_INVALID, "contexts %d and %d are the same", h, g);
  }
  const int R = std::min<int>(n_ctx, n);
  std::vector<PageStage> stages(R);
  std::vector<std::unique_ptr<CallGuard>> guards;
  for (int g = 0; g < R; ++g) guards.emplace_back(new CallGuard(dets[g]->ctx));
  std::vector<int> first(R + 1, 0);
  for (int g = 0; g < R; ++g) first[g + 1] = first[g] + n / R + (g < n % R ? 1 : 0);
  out->ms_h2d = out->ms_det = out->ms_post = out->ms_crop = out->ms_rec = out->ms_total = out->ms_cls = 0.0f;
  out->h2d_bytes = out->d2h_bytes = 0;

  // direct peer access where the hardware has it (NVLink / NVSwitch on a B200 box); without it the peer copies below
  // are staged by the driver.  "Already enabled" and "not supported" are both fine.
  for (int g = 0; g < R; ++g)
    for (int h = 0; h < R; ++h) {
      const int dg = dets[g]->ctx->device, dh = dets[h]->ctx->device;
      if (dg == dh) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, dg, dh) == cudaSuccess && can) {
        cudaSetDevice(dg);
        cudaDeviceEnablePeerAccess(dh, 0);
        cudaGetLastError();
      }
    }
  // a worker thread per context; errors come back as (code, message)
  std::vector<int> codes(R, OAR_OK);
  std::vector<std::string> msgs(R);
  auto run_on_all = [&](const std::function<void(int)>& fn) {
    std::vector<std::thread> th;
    for (int g = 0; g < R; ++g)
      th.emplace_back([&, g] {
        try {
          OAR_CUDA(cudaSetDevice(dets[g]->ctx->device));
          fn(g);
        } catch (const oar::OarError& e) {
          cudaGetLastError();
          codes[g] = e.code, msgs[g] = oar::g_err;
        } catch (const std::exception& e) {
          codes[g] = OAR_E_CUDA, msgs[g] = e.what();
        }
      });
    for (auto& t : th) t.join();
    for (int g = 0; g < R; ++g)
      if (codes[g] != OAR_OK) OAR_FAIL(codes[g], "context %d: %s", g, msgs[g].c_str());
  };

  // ---- phase A: upload, detect, sort, crop -- every context on its block of pages
  run_on_all([&](int g) {
    PageStage& S = stages[g];
    S.ctx = dets[g]->ctx;
    for (auto& e : S.ev) e = S.ctx->next_event();
    cudaEventRecord(S.ev[0], S.ctx->stream);
    stage_upload(S, images + first[g], hs + first[g], ws + first[g], first[g + 1] - first[g], 0);
    cudaEventRecord(S.ev[1], S.ctx->stream);
    stage_detect(S, dets[g], cfg);
    cudaEventRecord(S.ev[2], S.ctx->stream);
    stage_crop(S);
    stage_orient(S, nullptr);
    cudaEventRecord(S.ev[3], S.ctx->stream);
    OAR_CUDA(cudaStreamSynchronize(S.ctx->stream));  // the pools are complete before anybody fetches from them
  });

  // ---- phase B: one global recognize_global plan over all pools, chunks dealt round-robin
  std::vector<CropRef> refs;
  std::vector<std::vector<int>> ref_of_box(R);
  for (int g = 0; g < R; ++g)
    if (stages[g].n_boxes > 0) append_refs(stages[g], g, refs, &ref_of_box[g]);
  RecPlan plan;
  plan_recognition(refs, cfg->region_batch_size, plan);
  for (size_t c = 0; c < plan.chunks.size(); ++c) plan.chunks[c].stage = (int)(c % R);

  // ---- phase C: every context recognises its chunks; foreign crops come over the peer link first
  run_on_all([&](int g) {
    oar_ctx* ctx = recs[g]->ctx;
    std::vector<RecCrop> rc;
    for (RecChunk& ch : plan.chunks) {
      if (ch.stage != g) continue;
      rc.resize(ch.n);
      for (int k = 0; k < ch.n; ++k) {
        const CropRef& c = refs[plan.order[ch.first + k]];
        const uint8_t* p = c.p;
        if (c.stage != g) {
          const size_t bytes = (size_t)c.h * c.w * 3;
          uint8_t* local = ctx->arena.get<uint8_t>(bytes);
          OAR_CUDA(cudaMemcpyPeerAsync(local, ctx->device, c.p, stages[c.stage].ctx->device, bytes, ctx->stream));
          p = local;
        }
        rc[k] = RecCrop{p, c.h, c.w};
      }
      launch_rec_batch(recs[g], rc.data(), ch.n, cfg->n_chars, ch.out);
    }
    cudaEventRecord(stages[g].ev[4], ctx->stream);
    OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  });

  // ---- phase D: scatter in image order
  std::vector<RecResult> res;
  int64_t d2h = 0;
  collect_results(refs, plan, cfg->rec_score_thresh, res, &d2h);
  int r = 0;
  long long nl = 0;
  for (int g = 0; g < R; ++g) {
    if (stages[g].n_boxes == 0) ref_of_box[g].clear();
    scatter_stage(stages[g], first[g], ref_of_box[g], res, out, r, nl);
    out->h2d_bytes += stages[g].h2d_bytes;
    out->d2h_bytes += stages[g].d2h_bytes;
    float a = 0, b = 0, c = 0, t = 0;
    cudaSetDevice(stages[g].ctx->device);
    cudaEventElapsedTime(&a, stages[g].ev[0], stages[g].ev[1]);
    cudaEventElapsedTime(&b, stages[g].ev[2], stages[g].ev[3]);
    cudaEventElapsedTime(&c, stages[g].ev[3], stages[g].ev[4]);
    cudaEventElapsedTime(&t, stages[g].ev[0], stages[g].ev[4]);
    // the call's stage times are those of its slowest context
    out->ms_h2d = std::max(out->ms_h2d, a), out->ms_crop = std::max(out->ms_crop, b);
    out->ms_rec = std::max(out->ms_rec, c), out->ms_total = std::max(out->ms_total, t);
    out->ms_det = std::max(out->ms_det, stages[g].ms_det), out->ms_post = std::max(out->ms_post, stages[g].ms_post);
  }
  out->d2h_bytes += d2h;
  out->region_off[n] = r;
  if (out->label_off) out->label_off[r] = (int32_t)nl;
  API_CATCH
}

// ---- layout detection: LayoutDetectionAdapter::execute on the device ---------------------
int32_t oar_layout_rows(oar_model* encoder, oar_model* head, const uint8_t* const* images, const int32_t* hs,
                        const int32_t* ws, int32_t n, int32_t images_on_device, int32_t input_h, int32_t input_w,
                        float* rows, size_t rows_cap) {
  API_TRY
  if (!encoder || !head || encoder->kind != OAR_KIND_FEAT || head->kind != OAR_KIND_FEAT)
    OAR_FAIL(OAR_E_INVALID, "layout detection needs an encoder model and a head model (feature-extractor kind)");
  if (encoder->ctx != head->ctx) OAR_FAIL(OAR_E_INVALID, "both models must live on the same context");
  if (n <= 0 || !images || !hs || !ws) OAR_FAIL(OAR_E_INVALID, "images: expected non-empty slice, got empty slice");
  if (!rows) OAR_FAIL(OAR_E_INVALID, "null argument");
  if ((size_t)n * 300 * 6 > rows_cap) OAR_FAIL(OAR_E_CAPACITY, "rows need %zu floats, capacity %zu", (size_t)n * 1800, rows_cap);
  oar_ctx* ctx = encoder->ctx;
  CallGuard guard(ctx);
  const float* d_rows = layout_rows_device(encoder, head, images, hs, ws, n, images_on_device, input_h, input_w);
  OAR_CUDA(cudaMemcpyAsync(rows, d_rows, (size_t)n * 1800 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_layout_run(oar_model* encoder, oar_model* head, const uint8_t* const* images, const int32_t* hs,
                       const int32_t* ws, int32_t n, int32_t input_h, int32_t input_w, const oar_layout_config* cfg,
                       float* boxes, int32_t* classes, float* scores, int32_t* counts) {
  if (!cfg || !boxes || !classes || !scores || !counts) {
    oar::set_error("null argument");
    return OAR_E_INVALID;
  }
  if (n <= 0) {
    oar::set_error("images: expected non-empty slice, got empty slice");
    return OAR_E_INVALID;
  }
  std::vector<float> rows((size_t)n * 1800);
  int32_t rc = oar_layout_rows(encoder, head, images, hs, ws, n, 0, input_h, input_w, rows.data(), rows.size());
  if (rc != OAR_OK) return rc;
  std::vector<float> sw(n), sh(n);
  for (int i = 0; i < n; ++i) sw[i] = (float)ws[i], sh[i] = (float)hs[i];
  // postprocess_pp_doclayout (layout_detection_adapter.rs:631-1116): host code in the reference and here (csrc/layout.cu)
  return oar_layout_postprocess(rows.data(), n, 300, 6, sw.data(), sh.data(), cfg, boxes, classes, scores, counts);
}

int32_t oar_device_alloc(oar_ctx* ctx, size_t bytes, void** out) {
  API_TRY
  require_device(ctx);
  if (!out) OAR_FAIL(OAR_E_INVALID, "null argument");
  OAR_CUDA(cudaSetDevice(ctx->device));
  OAR_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  API_CATCH
}

int32_t oar_device_free(oar_ctx* ctx, void* p) {
  API_TRY
  require_device(ctx);
  OAR_CUDA(cudaSetDevice(ctx->device));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  OAR_CUDA(cudaFree(p));
  API_CATCH
}

int32_t oar_memcpy_h2d(oar_ctx* ctx, void* dst, const void* src, size_t bytes) {
  API_TRY
  require_device(ctx);
  OAR_CUDA(cudaSetDevice(ctx->device));
  OAR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  API_CATCH
}

int32_t oar_timer_start(oar_ctx* ctx) {
  API_TRY
  require_device(ctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  OAR_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->timer0) {
    OAR_CUDA(cudaEventCreate(&ctx->timer0));
    OAR_CUDA(cudaEventCreate(&ctx->timer1));
  }
  OAR_CUDA(cudaStreamSynchronize(ctx->stream));
  OAR_CUDA(cudaEventRecord(ctx->timer0, ctx->stream));
  API_CATCH
}

int32_t oar_timer_stop(oar_ctx* ctx, float* ms) {
  API_TRY
  require_device(ctx);
  if (!ms || !ctx->timer0) OAR_FAIL(OAR_E_INVALID, "timer was not started");
  std::lock_guard<std::mutex> lock(ctx->mu);
  OAR_CUDA(cudaSetDevice(ctx->device));
  OAR_CUDA(cudaEventRecord(ctx->timer1, ctx->stream));
  OAR_CUDA(cudaEventSynchronize(ctx->timer1));
  OAR_CUDA(cudaEventElapsedTime(ms, ctx->timer0, ctx->timer1));
  API_CATCH
}

int32_t oar_l2_flush(oar_ctx* ctx) {
  API_TRY
  require_device(ctx);
  std::lock_guard<std::mutex> lock(ctx->mu);
  OAR_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)256 << 20;
  if (!ctx->flush_buf) OAR_CUDA(cudaMalloc(&ctx->flush_buf, bytes));
  OAR_CUDA(cudaMemsetAsync(ctx->flush_buf, 0, bytes, ctx->stream));
  API_CATCH
}

int32_t oar_profile_enable(oar_ctx* ctx, int32_t on) {
  if (!ctx) return OAR_E_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->profile = on != 0;
  return OAR_OK;
}

int32_t oar_profile_read(oar_ctx* ctx, oar_kernel_record* recs, int32_t cap) {
  if (!ctx) return 0;
  std::lock_guard<std::mutex> lock(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  int n = 0;
  for (auto& r : ctx->prof) {
    if (n >= cap) break;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.0f;
    }
    recs[n].name = r.name;
    recs[n].ms = ms;
    recs[n].flops = r.flops;
    recs[n].bytes = r.bytes;
    ++n;
  }
  return n;
}

}  // extern "C"
