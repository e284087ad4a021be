// common.cuh -- shared host/device plumbing for liboar_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/oar_b200.h"

namespace oar {

void set_error(const char* fmt, ...);
extern thread_local char g_err[1024];
extern std::atomic<long long> g_launches;  // contexts on different GPUs launch concurrently
// host-visible submissions: a kernel launched directly counts one, a replayed CUDA graph (engine.cu: model_forward)
// counts one however many kernels it holds -- those still count in g_launches
extern std::atomic<long long> g_submits;
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); safe to call from concurrent contexts
void ensure_max_dynamic_smem(const void* kernel, int device, int bytes);

struct OarError {
  int code;
};

#define OAR_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      oar::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
      throw oar::OarError{OAR_E_CUDA};                                                      \
    }                                                                                       \
  } while (0)

#define OAR_FAIL(code, ...)        \
  do {                             \
    oar::set_error(__VA_ARGS__);   \
    throw oar::OarError{code};     \
  } while (0)

// Bump allocator over one growable device slab; reset per call.  Memory laid
// out for a 180 GB part: no reuse games, every activation keeps its own range.
struct Arena {
  struct Slab {
    char* base;
    size_t cap, used;
  };
  std::vector<Slab> slabs;
  size_t min_slab = (size_t)256 << 20;
  size_t total_alloc = 0;  // bytes handed out since creation (never decreases): the footprint of a graph walk is a difference
  bool fixed = false;      // the private arena of a captured CUDA graph: one slab sized beforehand, growing is an error
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    total_alloc += bytes;
    for (auto& s : slabs) {
      if (s.used + bytes <= s.cap) {
        void* p = s.base + s.used;
        s.used += bytes;
        return p;
      }
    }
    if (fixed) OAR_FAIL(OAR_E_CUDA, "graph arena exhausted (%zu more bytes wanted)", bytes);
    size_t cap = bytes > min_slab ? bytes : min_slab;
    char* p = nullptr;
    OAR_CUDA(cudaMalloc(&p, cap));
    slabs.push_back(Slab{p, cap, bytes});
    return p;
  }
  template <typename T>
  T* get(size_t n) {
    return reinterpret_cast<T*>(alloc(n * sizeof(T)));
  }
  // stack discipline for per-chunk scratch: stream order makes reuse safe
  std::vector<size_t> mark() const {
    std::vector<size_t> m;
    for (auto& s : slabs) m.push_back(s.used);
    return m;
  }
  void release_to(const std::vector<size_t>& m) {
    for (size_t i = 0; i < slabs.size(); ++i) slabs[i].used = i < m.size() ? m[i] : 0;
  }
  void reset() {
    // coalesce to one slab when fragmented so steady state is a single range
    if (slabs.size() > 1) {
      size_t total = 0;
      for (auto& s : slabs) {
        total += s.cap;
        cudaFree(s.base);
      }
      slabs.clear();
      char* p = nullptr;
      OAR_CUDA(cudaMalloc(&p, total));
      slabs.push_back(Slab{p, total, 0});
    }
    for (auto& s : slabs) s.used = 0;
  }
  void release() {
    for (auto& s : slabs) cudaFree(s.base);
    slabs.clear();
  }
};

struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
  double flops, bytes;
};

}  // namespace oar

struct oar_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  oar::Arena arena;
  // second launch stream with its own arena (capi.cu: two-stream recognition).  Every launch helper reads ctx->stream /
  // ctx->arena, so "the other lane" is entered by swapping both pairs; outside that bracket they are never visible.
  cudaStream_t stream_aux = nullptr;
  oar::Arena arena_aux;
  cudaStream_t stream_copy = nullptr;  // page uploads of a pipeline call: overlap the detector (capi.cu: stage_upload)
  std::mutex mu;
  bool capturing = false;       // a model_forward walk is being recorded into a CUDA graph (engine.cu)
  long long captured = 0;       // kernels recorded by the capture in progress
  void* graph_cache = nullptr;  // engine.cu: captured layer lists per (model, lane, shape)
  bool profile = false;
  std::vector<oar::ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  size_t event_next = 0;
  int sm_count = 148;
  cudaEvent_t timer0 = nullptr, timer1 = nullptr;  // oar_timer_start/stop
  void* flush_buf = nullptr;                       // oar_l2_flush scratch (256 MiB)
  // pinned host staging (bump allocated, reset at the start of every API call)
  struct PinnedSlab {
    char* base;
    size_t cap, used;
  };
  std::vector<PinnedSlab> pinned;
  void* pinned_get(size_t bytes);
  void pinned_reset();
  cudaEvent_t next_event();
  void begin_call();  // set device, reset arenas and profile records
};

namespace oar {

// RAII launch bracket: counts the launch and, when profiling, times it.
struct Launch {
  oar_ctx* ctx;
  int idx = -1;
  Launch(oar_ctx* c, const char* name, double flops = 0, double bytes = 0) : ctx(c) {
    if (c->capturing) {
      ++c->captured;  // recorded, not launched: counted when (and every time) the graph is
    } else {
      ++g_launches;
      ++g_submits;
    }
    if (c->profile) {
      ProfRec r{name, c->next_event(), c->next_event(), flops, bytes};
      cudaEventRecord(r.e0, c->stream);
      c->prof.push_back(r);
      idx = (int)c->prof.size() - 1;
    }
  }
  ~Launch() {
    if (idx >= 0) cudaEventRecord(ctx->prof[idx].e1, ctx->stream);
  }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace oar
