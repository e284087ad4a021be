// dbpost.cu -- DB detector post-processing on device (compiled with -fmad=false).
//
// Replaces DBPostProcess::apply (oar-ocr-core/src/processors/db_postprocess.rs:100-221),
// boxes_from_bitmap (db_bitmap.rs:84-150) and everything it calls:
//   threshold_to_mask (db_postprocess.rs:185-221), imageproc::find_contours (Suzuki-Abe border
//   following, db_bitmap.rs:100), get_mini_boxes_from_contour/points (db_bitmap.rs:153-202,
//   253-277), simplify_chain_points (:207-239), get_min_area_rect_from_points
//   (geometry.rs:310-441), box_score_fast + process_scanline (db_score.rs:34-134,
//   geometry.rs:1087-1164), unclip via Clipper2 round-join offsetting (db_bitmap.rs:279-368).
//
// The reference does this serially per image on one host thread.  Here:
//   1. threshold + union-find connected-component labelling (8-connectivity): the root of a
//      component is its raster-first pixel.  Border following only ever reads zero/non-zero,
//      and the visited marks of one component never influence another, so
//   2. one warp per component replays the raster scan restricted to that component's bounding
//      box and follows its outer and hole borders exactly as the sequential algorithm would
//      (8 lanes probe the 8-neighbourhood per step), appending points to a global pool;
//   3. contour records are radix-sorted by (image, start pixel) = discovery order, which also
//      implements the `take(max_candidates)` cut;
//   4. one warp per candidate does the geometry in the reference's own f32 operation order: lane 0 walks the
//      strictly sequential pieces (chain simplification, hull march, unclip), the lanes share box_score_fast
//      (whole-row partial sums, folded in row order), the Graham-scan sort (stable rank sort) and the
//      min-area-rectangle search (first strictly smaller edge wins);
//   5. survivors are compacted in discovery order.
#include <cub/device/device_radix_sort.cuh>

#include "prepost.cuh"

namespace oar {

// ---------------------------------------------------------------------------
// 1. threshold + CCL
// ---------------------------------------------------------------------------
__global__ void db_init_kernel(const float* __restrict__ pred, float thresh, uint8_t* __restrict__ state,
                               int32_t* __restrict__ lab, size_t total, int HW, int W) {
  // Launched with whole warps over the linear pixel index: each lane labels its pixel with the first pixel of its
  // horizontal run *within the warp's 32-pixel segment*, so the union pass only has to stitch segment and row
  // boundaries instead of every pixel pair.
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool fg = false;
  int li = 0, x = 0;
  if (i < total) {
    fg = pred[i] > thresh;  // strict, db_postprocess.rs:202
    li = (int)(i % HW);
    x = li % W;
  }
  const unsigned fgm = __ballot_sync(0xffffffffu, fg);
  const bool prev_fg = lane > 0 && ((fgm >> (lane - 1)) & 1u);
  const bool is_start = fg && (!prev_fg || x == 0 || li == 0);
  const unsigned starts = __ballot_sync(0xffffffffu, is_start);
  if (i >= total) return;
  state[i] = fg ? 1 : 0;
  if (fg) {
    const unsigned below = starts & (0xffffffffu >> (31 - lane));
    const int sl = 31 - __clz(below);
    lab[i] = li - (lane - sl);
  }
  // (background labels are never read: every reader tests the state byte first -- 4 bytes per pixel less to write)
}

__device__ __forceinline__ int32_t uf_find(const int32_t* lab, int32_t x) {
  int32_t p = lab[x];
  while (p != x) {
    x = p;
    p = lab[x];
  }
  return x;
}

__device__ __forceinline__ void uf_union(int32_t* lab, int32_t a, int32_t b) {
  bool done = false;
  while (!done) {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a == b) return;
    if (a > b) {
      int32_t t = a;
      a = b;
      b = t;
    }
    int32_t old = atomicMin(&lab[b], a);
    done = (old == b);
    b = old;
  }
}

// The three labelling passes below visit foreground pixels only; a thread takes FOUR consecutive pixels with one
// 32-bit load of their state bytes and leaves at once when all four are background (most of a page): a quarter of
// the threads of the one-pixel-per-thread form, which spent its time launching and retiring them.
template <typename F>
__device__ __forceinline__ void db_for_fg4(const uint8_t* __restrict__ state, size_t total, F&& body) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= total) return;
  if (i4 + 4 <= total && ((reinterpret_cast<uintptr_t>(state) + i4) & 3) == 0) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(state + i4);
    if (w == 0) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((w >> (8 * k)) & 0xffu) body(i4 + k);
  } else {
    for (size_t i = i4; i < total && i < i4 + 4; ++i)
      if (state[i]) body(i);
  }
}

__device__ __forceinline__ void db_merge_px(const uint8_t* __restrict__ state, int32_t* __restrict__ lab, int H, int W,
                                            size_t i) {
  int HW = H * W;
  int b = (int)(i / HW);
  int li = (int)(i - (size_t)b * HW);
  int y = li / W, x = li - y * W;
  const uint8_t* st = state + (size_t)b * HW;
  int32_t* L = lab + (size_t)b * HW;
  const bool left = x > 0 && st[li - 1];
  // a run that continues across the 32-pixel segment boundary of the init pass (or an image boundary in the linear
  // index): one union per segment instead of one per pixel
  if (left && (i & 31) == 0) uf_union(L, li, li - 1);  // lane 0 of the init warp: its run start was cut at the segment
  if (y > 0) {
    const bool up = st[li - W];
    const bool ul = x > 0 && st[li - W - 1];
    const bool ur = x + 1 < W && st[li - W + 1];
    if (up) {
      // redundant when the left neighbour already joined the same upper run through (x-1, y-1)
      if (!(left && ul)) uf_union(L, li, li - W);
    } else {
      if (ul && !left) uf_union(L, li, li - W - 1);
      if (ur) uf_union(L, li, li - W + 1);
    }
  }
}

__global__ void db_merge_kernel(const uint8_t* __restrict__ state, int32_t* __restrict__ lab, int B, int H, int W) {
  db_for_fg4(state, (size_t)B * H * W, [&](size_t i) { db_merge_px(state, lab, H, W, i); });
}

struct Comp {
  int img, root;
  int ymax, xmin, xmax;
};

__device__ __forceinline__ void db_flatten_px(int32_t* __restrict__ lab, int32_t* __restrict__ slot_of,
                                              Comp* __restrict__ comps, int* __restrict__ n_comps, int comp_cap, int H,
                                              int W, size_t i) {
  int HW = H * W;
  int b = (int)(i / HW);
  int li = (int)(i - (size_t)b * HW);
  int32_t* L = lab + (size_t)b * HW;
  int32_t r = uf_find(L, li);
  L[li] = r;  // benign race: every writer stores an ancestor-or-root on the same chain
  if (r == li) {
    int s = atomicAdd(n_comps, 1);
    if (s < comp_cap) {
      int y = li / W, x = li - y * W;
      comps[s] = Comp{b, li, y, x, x};
      slot_of[i] = s;
    } else {
      slot_of[i] = -1;
    }
  }
}

__global__ void db_flatten_kernel(const uint8_t* __restrict__ state, int32_t* __restrict__ lab,
                                  int32_t* __restrict__ slot_of, Comp* __restrict__ comps, int* __restrict__ n_comps,
                                  int comp_cap, int B, int H, int W) {
  db_for_fg4(state, (size_t)B * H * W,
             [&](size_t i) { db_flatten_px(lab, slot_of, comps, n_comps, comp_cap, H, W, i); });
}

__device__ __forceinline__ void db_bbox_px(const uint8_t* __restrict__ state, const int32_t* __restrict__ lab,
                                           const int32_t* __restrict__ slot_of, Comp* __restrict__ comps, int H, int W,
                                           size_t i) {
  int HW = H * W;
  int b = (int)(i / HW);
  int li = (int)(i - (size_t)b * HW);
  int y = li / W, x = li - y * W;
  bool wz = (x == 0) || !state[i - 1];
  bool ez = (x + 1 == W) || !state[i + 1];
  if (!wz && !ez) return;
  int32_t r = uf_find(lab + (size_t)b * HW, li);
  int s = slot_of[(size_t)b * HW + r];
  if (s < 0) return;
  atomicMax(&comps[s].ymax, y);
  if (wz) atomicMin(&comps[s].xmin, x);
  if (ez) atomicMax(&comps[s].xmax, x);
}

__global__ void db_bbox_kernel(const uint8_t* __restrict__ state, const int32_t* __restrict__ lab,
                               const int32_t* __restrict__ slot_of, Comp* __restrict__ comps, int B, int H, int W) {
  db_for_fg4(state, (size_t)B * H * W, [&](size_t i) { db_bbox_px(state, lab, slot_of, comps, H, W, i); });
}

// ---------------------------------------------------------------------------
// 2. border following, one warp per component
// ---------------------------------------------------------------------------
struct ContourRec {
  int img, start;   // start = y*W + x
  long long off;    // into the point pool
  int len;
};

// The 8-neighbour ring, clockwise from West: RX = {-1,-1,0,1,1,1,0,-1}, RY = {0,-1,-1,-1,0,1,1,1}, and its inverse
// DIR[(dx+1) + 3*(dy+1)] = {1,2,3,0,-,4,7,6,5}.  Packed into immediates (2 / 2 / 3 bits per entry) and read with a
// shift: as __constant__ arrays the eight probing lanes index them with eight different values, which the constant
// cache serialises -- twice per border pixel, on the one serial chain of the whole post-process.
__device__ __forceinline__ int ring_dx(int d) { return (int)((0x1a90u >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int ring_dy(int d) { return (int)((0xa901u >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int ring_of(int dx, int dy) { return (int)((0x5de00d1u >> (3 * ((dx + 1) + 3 * (dy + 1)))) & 7u); }

// The window a border is followed in: either the global state map of the image (one byte per pixel) or the block's
// shared-memory copy of one component, two bits per pixel (0 background, 1 component, 2 / 3 the visited marks), 16
// pixels per word -- four times the pixels of a byte map in the same 16 KB, so a page-wide text line still walks in
// shared memory.  Only lane 0 writes; the probing lanes only ask "non-zero?", which a mark never changes.
struct MapU8 {
  static constexpr bool kInside = false;  // the walk may step to the image's edge: probes are bounds-checked
  uint8_t* p;
  int stride;
  __device__ __forceinline__ int get(int x, int y) const { return p[(size_t)y * stride + x]; }
  __device__ __forceinline__ void set(int x, int y, int v) const { p[(size_t)y * stride + x] = (uint8_t)v; }
  // a pixel's storage unit, fetched early (next to the probes) and written back with a new value later
  __device__ __forceinline__ uint32_t fetch(int x, int y) const { return p[(size_t)y * stride + x]; }
  __device__ __forceinline__ int value(uint32_t unit, int) const { return (int)unit; }
  __device__ __forceinline__ void store(int x, int y, uint32_t, int v) const { p[(size_t)y * stride + x] = (uint8_t)v; }
};
struct MapPacked {
  static constexpr bool kInside = true;  // one ring of background around the component: its pixels' neighbours exist
  uint32_t* p;
  int wpr;  // words per row
  __device__ __forceinline__ int get(int x, int y) const { return (int)((p[y * wpr + (x >> 4)] >> ((x & 15) * 2)) & 3u); }
  __device__ __forceinline__ void set(int x, int y, int v) const {
    uint32_t* w = p + y * wpr + (x >> 4);
    const int sh = (x & 15) * 2;
    *w = (*w & ~(3u << sh)) | ((uint32_t)v << sh);
  }
  __device__ __forceinline__ uint32_t fetch(int x, int y) const { return p[y * wpr + (x >> 4)]; }
  __device__ __forceinline__ int value(uint32_t unit, int x) const { return (int)((unit >> ((x & 15) * 2)) & 3u); }
  __device__ __forceinline__ void store(int x, int y, uint32_t unit, int v) const {
    const int sh = (x & 15) * 2;
    p[y * wpr + (x >> 4)] = (unit & ~(3u << sh)) | ((uint32_t)v << sh);
  }
};

// Follows one border from (sx,sy) inside the window `st` (bw x bh pixels, window origin (ox,oy) in image coordinates;
// Wimg = image width for the right-edge rule).  start_dir = ring index of the adjacent
// zero pixel.  All 32 lanes execute; lanes 0..7 probe.  When pts != nullptr, writes points (image coordinates) and
// the visited marks (2 = +nbd, 3 = -nbd); only the first `cap` points are stored.  Returns the number of points.
// (The marks are idempotent: walking a border again leaves them as they are.)
template <class Map>
__device__ int follow_border(const Map st, int bw, int bh, int ox, int oy, int Wimg, int sx, int sy, int start_dir,
                             short2* pts, int cap, int lane) {
  // (Measured slower, round 2: every lane loading all eight neighbours and the walk as arithmetic on one 8-bit mask,
  // without the vote -- 0.63 against 0.50 ms: eight loads and their assembly are a longer chain than one probe + vote.)
  auto nonzero = [&](int x, int y) -> bool {
    if (Map::kInside) return st.get(x, y) != 0;
    return x >= 0 && y >= 0 && x < bw && y < bh && st.get(x, y) != 0;
  };
  int k = lane & 7;
  int d = (start_dir + k) & 7;
  bool hit = (lane < 8) && nonzero(sx + ring_dx(d), sy + ring_dy(d));
  unsigned mask = __ballot_sync(0xffffffffu, hit) & 0xffu;
  if (!mask) {
    if (pts && lane == 0) {
      if (cap > 0) pts[0] = make_short2((short)(sx + ox), (short)(sy + oy));
      st.set(sx, sy, 3);
    }
    return 1;
  }
  int k1 = __ffs(mask) - 1;
  int d1 = (start_dir + k1) & 7;
  int p1x = sx + ring_dx(d1), p1y = sy + ring_dy(d1);
  int p2x = p1x, p2y = p1y, p3x = sx, p3y = sy;
  int n = 0;
  for (;;) {
    if (pts && lane == 0 && n < cap) pts[n] = make_short2((short)(p3x + ox), (short)(p3y + oy));
    ++n;
    // lane 0 marks this pixel below; its storage unit is requested here, next to the probes, so that the mark adds
    // no load latency of its own to the step (only lane 0 writes the map: the unit cannot change in between)
    const uint32_t unit = (pts && lane == 0) ? st.fetch(p3x, p3y) : 0u;
    int front = ring_of(p2x - p3x, p2y - p3y);
    int dk = (front - 1 - k + 16) & 7;  // k = 7 -> front itself (examined last)
    bool h2 = (lane < 8) && nonzero(p3x + ring_dx(dk), p3y + ring_dy(dk));
    unsigned m2 = __ballot_sync(0xffffffffu, h2) & 0xffu;
    int k4 = __ffs(m2) - 1;  // never -1: p2 is non-zero
    int d4 = (front - 1 - k4 + 16) & 7;
    int p4x = p3x + ring_dx(d4), p4y = p3y + ring_dy(d4);
    int kE = (front - 5 + 16) & 7;  // the probe index that looks East
    bool right_edge = kE < k4;
    if (pts && lane == 0) {
      if (p3x + ox + 1 == Wimg || right_edge)
        st.store(p3x, p3y, unit, 3);
      else if (st.value(unit, p3x) == 1)
        st.store(p3x, p3y, unit, 2);
    }
    if (p4x == sx && p4y == sy && p3x == p1x && p3y == p1y) break;
    p2x = p3x, p2y = p3y;
    p3x = p4x, p3y = p4y;
  }
  return n;
}

// One block per component (grid-stride when there are more components than blocks).  All four warps stage the
// component's bounding box plus a one-pixel ring into shared memory as a membership mask (1 = pixel of THIS
// component, everything else 0 -- other components are never 8-adjacent, so the borders are unchanged); warp 0 then
// replays the raster scan and walks the borders in shared memory, where every probe costs ~30 cycles instead of an
// L2 round trip.  The visited marks live only in that copy.  The copy holds two bits per pixel (MapPacked): windows
// of up to 65536 pixels -- a page-wide text line -- fit the 16 KB tile; larger ones walk the global state map directly
// (warp 0 only).
constexpr int TRACE_TILE_BYTES = 16 * 1024;  // ~14 blocks per SM: the walk is serial per component, concurrency is what counts (48 KB tiles measured 1.6x slower)
constexpr int TRACE_THREADS = 128;
// Every border is walked ONCE: the walk is the serial chain of the post-process (one dependent probe per border pixel),
// and its length is only known at its end, so the points go to a per-block staging run first and are copied into the
// pool -- whose offset needs the length -- by the whole warp afterwards.  A border longer than the staging run (rare:
// > 8192 points) is walked a second time straight into the pool.
constexpr int TRACE_STAGE_POINTS = 8192;

// Where the walks' points and records go (one per kernel launch).
struct TraceOut {
  short2* pool;
  unsigned long long* pool_used;
  unsigned long long pool_cap;
  ContourRec* recs;
  int* n_recs;
  int rec_cap;
  int* err;
};

// A raster-scan candidate (cx, y) (image coordinates; a component pixel with background to its West or East): start an
// outer or a hole border there if the visited marks say so, walk it once, publish its points and its record.  Called by
// the whole (converged) warp; returns false when a capacity ran out.
template <class Map>
__device__ bool start_border(const Map st, const Comp& c, int ox, int oy, int ww, int wh, int W, int cx, int y,
                             const TraceOut& T, short2* my_stage, int lane) {
  const int s = st.get(cx - ox, y - oy);
  int start_dir = -1;
  if (s == 1 && cx > 0 && st.get(cx - ox - 1, y - oy) == 0)
    start_dir = 0;  // outer border, adjacent = West
  else if ((s == 1 || s == 2) && cx + 1 < W && st.get(cx - ox + 1, y - oy) == 0)
    start_dir = 4;  // hole border, adjacent = East
  if (start_dir < 0) return true;
  const int n = follow_border(st, ww, wh, ox, oy, W, cx - ox, y - oy, start_dir, my_stage, TRACE_STAGE_POINTS, lane);
  unsigned long long off = 0;
  int ri = -1;
  if (lane == 0) {
    off = atomicAdd(T.pool_used, (unsigned long long)n);
    ri = atomicAdd(T.n_recs, 1);
  }
  off = __shfl_sync(0xffffffffu, off, 0);
  ri = __shfl_sync(0xffffffffu, ri, 0);
  if (off + n > T.pool_cap || ri >= T.rec_cap) {
    if (lane == 0) atomicExch(T.err, 1);
    return false;
  }
  if (n <= TRACE_STAGE_POINTS) {
    __syncwarp();  // lane 0's staged points are visible to the warp
    for (int i = lane; i < n; i += 32) T.pool[off + i] = my_stage[i];
  } else {
    follow_border(st, ww, wh, ox, oy, W, cx - ox, y - oy, start_dir, T.pool + off, 0x7fffffff, lane);
  }
  if (lane == 0) T.recs[ri] = ContourRec{c.img, (int)((size_t)y * W + cx), (long long)off, n};
  __syncwarp();
  return true;
}

// Raster scan of one component in the global state map (a window too large for the tile; warp 0 of the block): 32
// pixels per step, one probe each.  member(x, y): does this non-zero pixel belong to THIS component (label comparison).
template <class Member>
__device__ void trace_component_global(const MapU8 st, Member member, const Comp& c, int y0, int H, int W, const TraceOut& T,
                                       short2* my_stage, int lane) {
  bool ok = true;
  for (int y = y0; y <= c.ymax && ok; ++y) {
    for (int xb = c.xmin; xb <= c.xmax && ok; xb += 32) {
      const int x = xb + lane;
      bool cand = false;
      if (x <= c.xmax && st.get(x, y) != 0 && member(x, y)) {
        const bool wz = (x > 0) && st.get(x - 1, y) == 0;
        const bool ez = (x + 1 < W) && st.get(x + 1, y) == 0;
        cand = wz || ez;
      }
      unsigned cm = __ballot_sync(0xffffffffu, cand);
      while (cm && ok) {
        const int l = __ffs(cm) - 1;
        cm &= cm - 1;
        ok = start_border(st, c, 0, 0, W, H, W, xb + l, y, T, my_stage, lane);
      }
    }
  }
}

// the non-zero-ness of the 16 two-bit pixels of a packed word as 16 bits
__device__ __forceinline__ uint32_t packed_nz16(uint32_t w) {
  uint32_t t = (w | (w >> 1)) & 0x55555555u;
  t = (t | (t >> 1)) & 0x33333333u;
  t = (t | (t >> 2)) & 0x0f0f0f0fu;
  t = (t | (t >> 4)) & 0x00ff00ffu;
  t = (t | (t >> 8)) & 0x0000ffffu;
  return t;
}

// Raster scan of one component in its packed tile (warp 0 of the block).  Whether a pixel is a candidate depends on
// membership only, never on the visited marks, so a row's candidates are found with word arithmetic -- lane j owns the
// 32-pixel chunk j of the row: member bits, shifted copies for the West / East neighbours, the image-edge rules -- and
// only chunks that hold a candidate are visited, in order.  (One probe per pixel made the scan of a 900 x 35 window
// ~1000 dependent steps before the first walk could even start.)
__device__ void trace_component_packed(const MapPacked st, const Comp& c, int y0, int ox, int oy, int bw, int bh, int W,
                                       const TraceOut& T, short2* my_stage, int lane) {
  const int chunks = st.wpr >> 1;
  const int lx_w = -ox;         // window column of image column 0: no West test there (x > 0)
  const int lx_e = W - 1 - ox;  // window column of the image's last column: no East test there (x + 1 < W)
  bool ok = true;
  for (int y = y0; y <= c.ymax && ok; ++y) {
    const uint32_t* row = st.p + (y - oy) * st.wpr;
    for (int g0 = 0; g0 < chunks && ok; g0 += 32) {
      const int ch = g0 + lane;
      uint32_t mask = 0;
      if (ch < chunks) {
        const uint32_t M = packed_nz16(row[2 * ch]) | (packed_nz16(row[2 * ch + 1]) << 16);
        const uint32_t pw = ch > 0 ? (packed_nz16(row[2 * ch - 1]) >> 15) & 1u : 0u;
        const uint32_t ne = 2 * ch + 2 < st.wpr ? packed_nz16(row[2 * ch + 2]) & 1u : 0u;
        uint32_t wz = M & ~((M << 1) | pw), ez = M & ~((M >> 1) | (ne << 31));
        if (lx_w >= 32 * ch && lx_w < 32 * ch + 32) wz &= ~(1u << (lx_w - 32 * ch));
        if (lx_e >= 32 * ch && lx_e < 32 * ch + 32) ez &= ~(1u << (lx_e - 32 * ch));
        mask = wz | ez;
      }
      unsigned any = __ballot_sync(0xffffffffu, mask != 0);
      while (any && ok) {
        const int j = __ffs(any) - 1;
        any &= any - 1;
        uint32_t cm = __shfl_sync(0xffffffffu, mask, j);
        while (cm && ok) {
          const int l = __ffs(cm) - 1;
          cm &= cm - 1;
          ok = start_border(st, c, ox, oy, bw, bh, W, ox + 32 * (g0 + j) + l, y, T, my_stage, lane);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(TRACE_THREADS) db_trace_kernel(
    uint8_t* __restrict__ state, const int32_t* __restrict__ lab, const Comp* __restrict__ comps,
    const int* __restrict__ n_comps, int comp_cap, int H, int W, short2* __restrict__ pool,
    unsigned long long* __restrict__ pool_used, unsigned long long pool_cap, ContourRec* __restrict__ recs,
    int* __restrict__ n_recs, int rec_cap, int* __restrict__ err, short2* __restrict__ stage) {
  extern __shared__ uint32_t tile[];
  short2* my_stage = stage + (size_t)blockIdx.x * TRACE_STAGE_POINTS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = min(*n_comps, comp_cap);
  for (int comp = blockIdx.x; comp < nc; comp += gridDim.x) {
    const Comp c = comps[comp];
    const size_t HW = (size_t)H * W;
    uint8_t* img = state + (size_t)c.img * HW;
    const int32_t* L = lab + (size_t)c.img * HW;
    const int y0 = c.root / W;
    const int bw = c.xmax - c.xmin + 3, bh = c.ymax - y0 + 3;  // bounding box + ring
    const int wpr = (bw + 31) >> 5 << 1;                       // 16 pixels per word, rows padded to 32 pixels
    const bool staged = (size_t)wpr * bh * 4 <= (size_t)TRACE_TILE_BYTES;
    const int ox = staged ? c.xmin - 1 : 0, oy = staged ? y0 - 1 : 0;
    if (staged) {
      // one warp per row, 32 pixels per step: the membership bits of a ballot, spread to two bits per pixel.  Four steps
      // at a time with the state byte and the label loaded side by side (the label of a background pixel is never
      // written -- db_init_kernel -- and never used: the state byte gates it), so eight loads are in flight per lane
      // instead of two dependent ones.
      for (int ly = warp; ly < bh; ly += TRACE_THREADS / 32) {
        const int gy = oy + ly;
        const bool row_in = gy >= 0 && gy < H;
        for (int lx0 = 0; lx0 < 16 * wpr; lx0 += 128) {
          uint8_t sb[4];
          int32_t lb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int lx = lx0 + 32 * u + lane, gx = ox + lx;
            sb[u] = 0, lb[u] = 0;
            if (row_in && lx < bw && gx >= 0 && gx < W) {
              const size_t o = (size_t)gy * W + gx;
              sb[u] = img[o], lb[u] = L[o];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (lx0 + 32 * u >= 16 * wpr) break;  // uniform
            const unsigned m = __ballot_sync(0xffffffffu, sb[u] != 0 && lb[u] == c.root);
            if ((lane & 15) == 0) {
              uint32_t b = (m >> lane) & 0xffffu;  // 16 membership bits -> even bit positions
              b = (b | (b << 8)) & 0x00ff00ffu;
              b = (b | (b << 4)) & 0x0f0f0f0fu;
              b = (b | (b << 2)) & 0x33333333u;
              b = (b | (b << 1)) & 0x55555555u;
              tile[ly * wpr + ((lx0 + 32 * u + lane) >> 4)] = b;
            }
          }
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      const TraceOut T{pool, pool_used, pool_cap, recs, n_recs, rec_cap, err};
      if (staged)
        trace_component_packed(MapPacked{tile, wpr}, c, y0, ox, oy, bw, bh, W, T, my_stage, lane);
      else
        trace_component_global(MapU8{img, W}, [&](int x, int y) { return L[(size_t)y * W + x] == c.root; }, c, y0, H, W, T,
                               my_stage, lane);
    }
    __syncthreads();  // the tile is reused by this block's next component
  }
}

// ---------------------------------------------------------------------------
// 3. ordering helpers
// ---------------------------------------------------------------------------
__global__ void db_keys_kernel(const ContourRec* __restrict__ recs, const int* __restrict__ n_recs, int rec_cap,
                               unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rec_cap) return;
  int n = min(*n_recs, rec_cap);
  if (i < n) {
    keys[i] = ((unsigned long long)(unsigned)recs[i].img << 32) | (unsigned)recs[i].start;
    vals[i] = i;
  } else {
    keys[i] = ~0ull;
    vals[i] = -1;
  }
}

__global__ void db_img_first_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ n_recs,
                                    int rec_cap, int B, int* __restrict__ first) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  int n = min(*n_recs, rec_cap);
  unsigned long long key = (unsigned long long)(unsigned)b << 32;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys[mid] < key)
      lo = mid + 1;
    else
      hi = mid;
  }
  first[b] = lo;
}

// ---------------------------------------------------------------------------
// 4. per-candidate geometry (one thread each)
// ---------------------------------------------------------------------------
struct MinRect {
  float cx, cy, w, h, angle;
};

__device__ __forceinline__ int total_cmp_f32(float a, float b) {
  int ia = __float_as_int(a), ib = __float_as_int(b);
  ia ^= (int)(((unsigned)(ia >> 31)) >> 1);
  ib ^= (int)(((unsigned)(ib >> 31)) >> 1);
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

// min-area rect over a convex hull given in the reference's hull order (geometry.rs:355-440)
__device__ MinRect min_rect_from_hull(const float* hx, const float* hy, int n) {
  MinRect best{0.f, 0.f, 0.f, 0.f, 0.f};
  float min_area = 3.402823466e+38f;
  const float PI_F = 3.14159265358979323846f;
  for (int i = 0; i < n; ++i) {
    int j = (i + 1) % n;
    float ex = hx[j] - hx[i], ey = hy[j] - hy[i];
    float l2 = ex * ex + ey * ey;
    if (l2 < 1.1920929e-7f) continue;
    float inv = 1.0f / sqrtf(l2);
    float nx = ex * inv, ny = ey * inv;
    float px = -ny, py = nx;
    float hix = hx[i], hiy = hy[i];
    float min_n = 3.402823466e+38f, max_n = -3.402823466e+38f, min_p = 3.402823466e+38f, max_p = -3.402823466e+38f;
    for (int q = 0; q < n; ++q) {
      float dx = hx[q] - hix, dy = hy[q] - hiy;
      float pn = nx * dx + ny * dy;
      float pp = px * dx + py * dy;
      if (pn < min_n) min_n = pn;
      if (pn > max_n) max_n = pn;
      if (pp < min_p) min_p = pp;
      if (pp > max_p) max_p = pp;
    }
    float width = max_n - min_n, height = max_p - min_p;
    float area = width * height;
    if (area < min_area) {
      min_area = area;
      float cn = (min_n + max_n) * 0.5f, cp = (min_p + max_p) * 0.5f;
      best.cx = hix + cn * nx + cp * px;
      best.cy = hiy + cn * ny + cp * py;
      best.w = width;
      best.h = height;
      best.angle = atan2f(ny, nx) * 180.0f / PI_F;
    }
  }
  return best;
}

// The same search with the hull edges spread over a warp.  The sequential loop keeps the FIRST edge whose area is
// strictly smaller (geometry.rs:419); each lane walks its edges in ascending order and the butterfly prefers the
// smaller area, then the smaller edge index, which selects the same edge.  Per-edge arithmetic is unchanged.
__device__ MinRect min_rect_from_hull_warp(const float* hx, const float* hy, int n, int lane) {
  MinRect best{0.f, 0.f, 0.f, 0.f, 0.f};
  float min_area = 3.402823466e+38f;
  int best_i = 0x7fffffff;
  const float PI_F = 3.14159265358979323846f;
  for (int i = lane; i < n; i += 32) {
    int j = (i + 1) % n;
    float ex = hx[j] - hx[i], ey = hy[j] - hy[i];
    float l2 = ex * ex + ey * ey;
    if (l2 < 1.1920929e-7f) continue;
    float inv = 1.0f / sqrtf(l2);
    float nx = ex * inv, ny = ey * inv;
    float px = -ny, py = nx;
    float hix = hx[i], hiy = hy[i];
    float min_n = 3.402823466e+38f, max_n = -3.402823466e+38f, min_p = 3.402823466e+38f, max_p = -3.402823466e+38f;
    for (int q = 0; q < n; ++q) {
      float dx = hx[q] - hix, dy = hy[q] - hiy;
      float pn = nx * dx + ny * dy;
      float pp = px * dx + py * dy;
      if (pn < min_n) min_n = pn;
      if (pn > max_n) max_n = pn;
      if (pp < min_p) min_p = pp;
      if (pp > max_p) max_p = pp;
    }
    float width = max_n - min_n, height = max_p - min_p;
    float area = width * height;
    if (area < min_area) {
      min_area = area;
      best_i = i;
      float cn = (min_n + max_n) * 0.5f, cp = (min_p + max_p) * 0.5f;
      best.cx = hix + cn * nx + cp * px;
      best.cy = hiy + cn * ny + cp * py;
      best.w = width;
      best.h = height;
      best.angle = atan2f(ny, nx) * 180.0f / PI_F;
    }
  }
  for (int o = 16; o; o >>= 1) {
    float oa = __shfl_xor_sync(0xffffffffu, min_area, o);
    int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    MinRect ob;
    ob.cx = __shfl_xor_sync(0xffffffffu, best.cx, o);
    ob.cy = __shfl_xor_sync(0xffffffffu, best.cy, o);
    ob.w = __shfl_xor_sync(0xffffffffu, best.w, o);
    ob.h = __shfl_xor_sync(0xffffffffu, best.h, o);
    ob.angle = __shfl_xor_sync(0xffffffffu, best.angle, o);
    if (oa < min_area || (oa == min_area && oi < best_i)) min_area = oa, best_i = oi, best = ob;
  }
  return best;
}

// degenerate hull (< 3 points) branch, geometry.rs:326-353
__device__ MinRect aabb_rect(const float* x, const float* y, int n) {
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = 0; i < n; ++i) {
    if (x[i] < mnx) mnx = x[i];
    if (x[i] > mxx) mxx = x[i];
    if (y[i] < mny) mny = y[i];
    if (y[i] > mxy) mxy = y[i];
  }
  if (!isfinite(mnx)) return MinRect{0.f, 0.f, 0.f, 0.f, 0.f};
  return MinRect{(mnx + mxx) * 0.5f, (mny + mxy) * 0.5f, mxx - mnx, mxy - mny, 0.0f};
}

// box_points_without_reorder + paddlex_order_mini_box_points (db_bitmap.rs:187-202, 253-277)
__device__ bool rect_to_ordered_box(const MinRect& r, float bx[4], float by[4], float* min_side) {
  float ms = fminf(r.w, r.h);
  if (!isfinite(ms) || ms <= 0.0f) return false;
  const float PI_F = 3.14159265358979323846f;
  float rad = r.angle * PI_F / 180.0f;
  // libm cosf/sinf are evaluated in double and rounded once; do the same
  float ca = (float)cos((double)rad), sa = (float)sin((double)rad);
  float w2 = r.w / 2.0f, h2 = r.h / 2.0f;
  float cxs[4] = {-w2, w2, w2, -w2}, cys[4] = {-h2, -h2, h2, h2};
  float px[4], py[4];
  for (int i = 0; i < 4; ++i) {
    px[i] = cxs[i] * ca - cys[i] * sa + r.cx;
    py[i] = cxs[i] * sa + cys[i] * ca + r.cy;
  }
  for (int a = 1; a < 4; ++a) {  // stable sort by x
    float kx = px[a], ky = py[a];
    int b = a - 1;
    while (b >= 0 && px[b] > kx) {
      px[b + 1] = px[b], py[b + 1] = py[b];
      --b;
    }
    px[b + 1] = kx, py[b + 1] = ky;
  }
  int i1, i4, i2, i3;
  if (py[1] > py[0]) i1 = 0, i4 = 1; else i1 = 1, i4 = 0;
  if (py[3] > py[2]) i2 = 2, i3 = 3; else i2 = 3, i3 = 2;
  bx[0] = px[i1], by[0] = py[i1];
  bx[1] = px[i2], by[1] = py[i2];
  bx[2] = px[i3], by[2] = py[i3];
  bx[3] = px[i4], by[3] = py[i4];
  *min_side = ms;
  return true;
}

__device__ __forceinline__ long long sat_usize_dev(float v) {
  if (!(v > 0.0f)) return 0;
  if (v >= 9.2e18f) return 0x7fffffffffffffffLL;
  return (long long)v;
}

constexpr int UNCLIP_CAP = 768;

// unclip (db_bitmap.rs:279-368): Clipper2 ClipperOffset, JoinType::Round, EndType::Polygon,
// precision 2.  Writes the offset polygon's vertices as f32.  Returns the vertex count,
// 0 for "empty", -1 on capacity overflow.
__device__ int unclip_dev(const float bx[4], const float by[4], float ratio, float* ox, float* oy) {
  double pxs[4], pys[4];
  for (int i = 0; i < 4; ++i) pxs[i] = (double)bx[i], pys[i] = (double)by[i];
  double a = 0.0;
  {
    double prx = pxs[3], pry = pys[3];
    for (int i = 0; i < 4; ++i) {
      a += (pry + pys[i]) * (prx - pxs[i]);
      prx = pxs[i], pry = pys[i];
    }
    a *= 0.5;
  }
  double area = fabs(a);
  if (area <= 2.220446049250313e-16) return 0;
  double per = 0.0;
  {
    double p1x = pxs[0], p1y = pys[0];
    for (int i = 1; i < 4; ++i) {
      per += hypot(pxs[i] - p1x, pys[i] - p1y);
      p1x = pxs[i], p1y = pys[i];
    }
    per += hypot(pxs[0] - p1x, pys[0] - p1y);
  }
  if (per <= 2.220446049250313e-16) return 0;
  double delta = area * (double)ratio / per;
  if (fabs(delta) <= 2.220446049250313e-16) return 0;
  const double scale = 100.0;
  long long ix[4], iy[4];
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    long long qx = llround(pxs[i] * scale), qy = llround(pys[i] * scale);
    if (n > 0 && ix[n - 1] == qx && iy[n - 1] == qy) continue;
    ix[n] = qx, iy[n] = qy, ++n;
  }
  if (n > 1 && ix[0] == ix[n - 1] && iy[0] == iy[n - 1]) --n;
  if (n < 3) return 0;
  double d = delta * scale;
  double ia = 0.0;
  {
    long long prx = ix[n - 1], pry = iy[n - 1];
    for (int i = 0; i < n; ++i) {
      ia += (double)(pry + iy[i]) * (double)(prx - ix[i]);
      prx = ix[i], pry = iy[i];
    }
  }
  double gd = ia < 0.0 ? -d : d;
  double ad = fabs(gd);
  const double PI_D = 3.141592653589793238;
  double arc_tol = log10(2.0 + ad) * 0.25;
  double steps_per_360 = fmin(PI_D / acos(1.0 - arc_tol / ad), ad * PI_D);
  double step_sin = sin(2.0 * PI_D / steps_per_360);
  double step_cos = cos(2.0 * PI_D / steps_per_360);
  if (gd < 0.0) step_sin = -step_sin;
  double steps_per_rad = steps_per_360 / (2.0 * PI_D);
  double nxs[4], nys[4];
  for (int i = 0; i < n; ++i) {
    int j = (i + 1) % n;
    double dx = (double)(ix[j] - ix[i]), dy = (double)(iy[j] - iy[i]);
    if (dx == 0 && dy == 0) {
      nxs[i] = 0, nys[i] = 0;
      continue;
    }
    double inv = 1.0 / hypot(dx, dy);
    dx *= inv, dy *= inv;
    nxs[i] = dy, nys[i] = -dx;
  }
  int m = 0;
  bool overflow = false;
  auto push = [&](double x, double y) {
    if (m >= UNCLIP_CAP) {
      overflow = true;
      return;
    }
    long long qx = llround(x), qy = llround(y);
    ox[m] = (float)((double)qx * (1.0 / scale));
    oy[m] = (float)((double)qy * (1.0 / scale));
    ++m;
  };
  for (int j = 0, k = n - 1; j < n; k = j, ++j) {
    double sin_a = nys[j] * nxs[k] - nys[k] * nxs[j];
    double cos_a = nxs[j] * nxs[k] + nys[j] * nys[k];
    if (sin_a > 1.0) sin_a = 1.0; else if (sin_a < -1.0) sin_a = -1.0;
    double ptx = (double)ix[j], pty = (double)iy[j];
    if (cos_a > -0.999 && (sin_a * gd < 0)) {
      push(ptx + nxs[k] * gd, pty + nys[k] * gd);
      push(ptx, pty);
      push(ptx + nxs[j] * gd, pty + nys[j] * gd);
    } else {
      double ang = atan2(sin_a, cos_a);
      double vx = nxs[k] * gd, vy = nys[k] * gd;
      push(ptx + vx, pty + vy);
      int steps = (int)ceil(steps_per_rad * fabs(ang));
      for (int i = 1; i < steps; ++i) {
        double tx = vx * step_cos - step_sin * vy;
        double ty = vx * step_sin + vy * step_cos;
        vx = tx, vy = ty;
        push(ptx + vx, pty + vy);
      }
      push(ptx + nxs[j] * gd, pty + nys[j] * gd);
    }
  }
  if (overflow) return -1;
  if (m > 1 && fabsf(ox[0] - ox[m - 1]) < 1.1920929e-7f && fabsf(oy[0] - oy[m - 1]) < 1.1920929e-7f) --m;
  if (m < 3) return 0;
  return m;
}

struct Cand {
  int valid;
  float box[8];
  float score;
};

// box_score_fast with the rows of the box spread over a warp.  The reference sums each scanline left to right
// from 0.0 and then adds the row partials in row order (db_score.rs:90-133); every lane therefore produces whole-row
// partials into `part` and lane 0 folds them in row order, which keeps the f32 association bit-identical.
__device__ float box_score_fast_warp(const float* __restrict__ pred, int W, int H, const float bx[4], const float by[4],
                                     float* part, int lane) {
  constexpr int CHUNK = 256;  // rows per pass (size of `part`)
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = 0; i < 4; ++i) {
    if (bx[i] < mnx) mnx = bx[i];
    if (bx[i] > mxx) mxx = bx[i];
    if (by[i] < mny) mny = by[i];
    if (by[i] > mxy) mxy = by[i];
  }
  float fminx = fminf(fmaxf(floorf(mnx), 0.0f), (float)W - 1.0f);
  float fmaxx = fminf(fmaxf(ceilf(mxx), 0.0f), (float)W - 1.0f);
  float fminy = fminf(fmaxf(floorf(mny), 0.0f), (float)H - 1.0f);
  float fmaxy = fminf(fmaxf(ceilf(mxy), 0.0f), (float)H - 1.0f);
  long long start_y = sat_usize_dev(fminy), end_y = sat_usize_dev(fmaxy) + 1;
  long long start_x = sat_usize_dev(fminx), end_x = sat_usize_dev(fmaxx) + 1;
  float total = 0.0f;
  long long total_px = 0;
  for (long long y0 = start_y; y0 < end_y; y0 += CHUNK) {
    long long y1 = y0 + CHUNK < end_y ? y0 + CHUNK : end_y;
    long long my_px = 0;
    for (long long yy = y0 + lane; yy < y1; yy += 32) {
      float y = (float)yy + 0.5f;
      float xs[4];
      int nx = 0;
      for (int i = 0; i < 4; ++i) {
        int j = (i + 1) & 3;
        float p1x = bx[i], p1y = by[i], p2x = bx[j], p2y = by[j];
        if (((p1y <= y && y < p2y) || (p2y <= y && y < p1y)) && fabsf(p2y - p1y) > 1.1920929e-7f) {
          xs[nx++] = p1x + (y - p1y) * (p2x - p1x) / (p2y - p1y);
        }
      }
      for (int a = 1; a < nx; ++a) {  // stable insertion sort
        float kx = xs[a];
        int b = a - 1;
        while (b >= 0 && xs[b] > kx) {
          xs[b + 1] = xs[b];
          --b;
        }
        xs[b + 1] = kx;
      }
      float line = 0.0f;
      long long yi = sat_usize_dev(y);
      if (yi < H) {
        const float* row = pred + (size_t)yi * W;
        for (int k = 0; k + 1 < nx; k += 2) {
          long long x1 = sat_usize_dev(fmaxf(xs[k], (float)start_x));
          long long x2 = sat_usize_dev(fminf(xs[k + 1], (float)end_x));
          if (x1 < x2 && x1 >= start_x && x2 <= end_x) {
            long long xe = x2 < W ? x2 : W;
            if (x1 < xe) {
              for (long long x = x1; x < xe; ++x) line += row[x];  // strictly left to right
              my_px += xe - x1;
            }
          }
        }
      }
      part[yy - y0] = line;
    }
    __syncwarp();
    for (int o = 16; o; o >>= 1) my_px += __shfl_xor_sync(0xffffffffu, my_px, o);
    total_px += my_px;
    if (lane == 0)
      for (long long yy = y0; yy < y1; ++yy) total += part[yy - y0];  // row order
    __syncwarp();
  }
  total = __shfl_sync(0xffffffffu, total, 0);
  return total_px > 0 ? total / (float)total_px : 0.0f;
}

// simplify_chain_points (db_bitmap.rs:207-239) by the whole warp: a point stays when the direction of the chain changes
// at it.  Every lane tests the points lane, lane + 32, ... against their two neighbours; a ballot and a prefix count
// keep the survivors in chain order, so `scratch` receives exactly the sequence the sequential form writes.  (As one
// lane's loop this was the longest serial piece of the geometry kernel: a dependent L2 round trip per contour point,
// ~2000 points on a page-wide text line.)  Returns the number of survivors.
__device__ int simplify_chain_warp(const ContourRec& rec, const short2* __restrict__ pool, short2* __restrict__ scratch,
                                   int lane) {
  const short2* pts = pool + rec.off;
  short2* simp = scratch + rec.off;
  const int n = rec.len;
  int ns = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool keep = false;
    short2 cur = make_short2(0, 0);
    if (i < n) {
      const short2 prev = pts[i == 0 ? n - 1 : i - 1], nxt = pts[i + 1 == n ? 0 : i + 1];
      cur = pts[i];
      const int a0 = (cur.x > prev.x) - (cur.x < prev.x), a1 = (cur.y > prev.y) - (cur.y < prev.y);
      const int b0 = (nxt.x > cur.x) - (nxt.x < cur.x), b1 = (nxt.y > cur.y) - (nxt.y < cur.y);
      keep = a0 != b0 || a1 != b1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) simp[ns + __popc(m & ((1u << lane) - 1u))] = cur;
    ns += __popc(m);
  }
  __syncwarp();  // the survivors are read back by lane 0
  return ns;
}

// Convex hull of the chain by the whole warp.  All cross products are exact integers, so the reference's Graham scan
// yields the unique strict hull, starting at the lowest-y (then lowest-x) point in increasing polar angle; a Jarvis
// march reproduces that sequence.  Each step is an arg-max over the points under "more clockwise seen from the
// current vertex, farther among collinear" -- a strict weak order, because the current vertex is an extreme point and
// every other point lies in a wedge of less than 180 degrees -- so the lanes reduce strided subsets and the partial
// winners with the same rule and arrive at the coordinates the sequential scan finds.  (One lane's scan of ~1000
// simplified points per hull vertex was the longest serial piece of the kernel after the simplification.)
// hx / hy (shared, UNCLIP_CAP floats each) receive the hull; returns its size, or -1 when it does not fit.
__device__ __forceinline__ bool hull_better(int cx, int cy, int bx, int by, int qx, int qy) {
  const long long cr = (long long)(bx - cx) * (qy - cy) - (long long)(by - cy) * (qx - cx);
  if (cr < 0) return true;
  if (cr == 0) {
    const long long d1 = (long long)(bx - cx) * (bx - cx) + (long long)(by - cy) * (by - cy);
    const long long d2 = (long long)(qx - cx) * (qx - cx) + (long long)(qy - cy) * (qy - cy);
    const long long dot = (long long)(bx - cx) * (qx - cx) + (long long)(by - cy) * (qy - cy);
    return dot > 0 && d2 > d1;
  }
  return false;
}

__device__ int hull_march_warp(const short2* P, int np, float* hx, float* hy, int* err, int lane) {
  const unsigned FULL = 0xffffffffu;
  int sx = 0x7fffffff, sy = 0x7fffffff;
  for (int i = lane; i < np; i += 32) {
    const short2 q = P[i];
    if (q.y < sy || (q.y == sy && q.x < sx)) sx = q.x, sy = q.y;
  }
  for (int off = 16; off; off >>= 1) {
    const int ox = __shfl_xor_sync(FULL, sx, off), oy = __shfl_xor_sync(FULL, sy, off);
    if (oy < sy || (oy == sy && ox < sx)) sx = ox, sy = oy;
  }
  int cx = sx, cy = sy, pcx = 0, pcy = 0, nh = 0;
  for (;;) {
    if (nh >= UNCLIP_CAP) {
      if (lane == 0) atomicExch(err, 2);
      return -1;
    }
    if (lane == 0) hx[nh] = (float)cx, hy[nh] = (float)cy;
    ++nh;
    int bx = cx, by = cy;
    bool have = false;
    for (int i = lane; i < np; i += 32) {
      const short2 q = P[i];
      if (q.x == cx && q.y == cy) continue;
      if (!have || hull_better(cx, cy, bx, by, q.x, q.y)) bx = q.x, by = q.y, have = true;
    }
    for (int off = 16; off; off >>= 1) {
      const int ox = __shfl_xor_sync(FULL, bx, off), oy = __shfl_xor_sync(FULL, by, off);
      const bool oh = __shfl_xor_sync(FULL, have ? 1 : 0, off) != 0;
      if (oh && (!have || hull_better(cx, cy, bx, by, ox, oy))) bx = ox, by = oy, have = true;
    }
    // (duplicated points can leave different lanes with different copies of the same coordinates: harmless)
    if (!have) break;
    if (bx == sx && by == sy) break;
    if (nh >= 2 && bx == pcx && by == pcy) break;  // collinear degenerate set: the march would bounce between the extremes
    pcx = cx, pcy = cy;
    cx = bx, cy = by;
  }
  __syncwarp();
  return nh;
}

// phase 1 tail (one lane): hull (nh points in hx / hy) -> min-area rect -> ordered mini box
__device__ bool cand_mini_box(const short2* P, int np, int nh, const float* hx, const float* hy, float bx[4], float by[4],
                              float* min_side) {
  MinRect r;
  if (nh < 3) {
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;  // degenerate: AABB of the points
    for (int i = 0; i < np; ++i) {
      float fx = (float)P[i].x, fy = (float)P[i].y;
      if (fx < mnx) mnx = fx;
      if (fx > mxx) mxx = fx;
      if (fy < mny) mny = fy;
      if (fy > mxy) mxy = fy;
    }
    r = MinRect{(mnx + mxx) * 0.5f, (mny + mxy) * 0.5f, mxx - mnx, mxy - mny, 0.0f};
  } else {
    r = min_rect_from_hull(hx, hy, nh);
  }
  return rect_to_ordered_box(r, bx, by, min_side);
}

// phase 2 (warp): unclip (lane 0) -> Graham hull with a stable rank sort spread over the lanes -> min-area rect
// (edges spread over the lanes) -> ordered box -> rescale, clamp.  ws = per-warp shared scratch, 6 x UNCLIP_CAP floats.
__device__ bool cand_unclip_box_warp(const float bx[4], const float by[4], float unclip_ratio, float min_size, float* ws,
                                     unsigned dw, unsigned dh, int W, int H, float* out_box, int* err, int lane) {
  float* ux = ws;
  float* uy = ws + UNCLIP_CAP;
  float* tx = ws + 2 * UNCLIP_CAP;
  float* ty = ws + 3 * UNCLIP_CAP;
  float* ang = ws + 4 * UNCLIP_CAP;
  float* dist = ws + 5 * UNCLIP_CAP;
  int nu = 0;
  if (lane == 0) nu = unclip_dev(bx, by, unclip_ratio, ux, uy);
  nu = __shfl_sync(0xffffffffu, nu, 0);
  if (nu < 0) {
    if (lane == 0) atomicExch(err, 3);
    return false;
  }
  if (nu == 0) return false;
  __syncwarp();
  // convex_hull_from_points (geometry.rs:226-274): start = lowest y then lowest x, swapped to the front
  if (lane == 0) {
    int s = 0;
    for (int i = 1; i < nu; ++i)
      if (uy[i] < uy[s] || (uy[i] == uy[s] && ux[i] < ux[s])) s = i;
    float t0 = ux[0], t1 = uy[0];
    ux[0] = ux[s], uy[0] = uy[s];
    ux[s] = t0, uy[s] = t1;
  }
  __syncwarp();
  const float spx = ux[0], spy = uy[0];
  for (int i = 1 + lane; i < nu; i += 32) {
    float dx = ux[i] - spx, dy = uy[i] - spy;
    ang[i] = atan2f(dy, dx);
    dist[i] = dx * dx + dy * dy;
  }
  __syncwarp();
  // stable sort of points 1.. by (angle, distance) under total_cmp: rank = number of elements that sort before
  for (int i = 1 + lane; i < nu; i += 32) {
    const float ka = ang[i], kd = dist[i];
    int rank = 1;
    for (int j = 1; j < nu; ++j) {
      int c = total_cmp_f32(ang[j], ka);
      if (c == 0) c = total_cmp_f32(dist[j], kd);
      if (c < 0 || (c == 0 && j < i)) ++rank;
    }
    tx[rank] = ux[i];
    ty[rank] = uy[i];
  }
  if (lane == 0) tx[0] = spx, ty[0] = spy;
  __syncwarp();
  int h2 = 0;
  if (lane == 0) {  // Graham scan: pop while cross <= 0
    for (int i = 0; i < nu; ++i) {
      float px = tx[i], py = ty[i];
      while (h2 > 1) {
        float cr = (ux[h2 - 1] - ux[h2 - 2]) * (py - uy[h2 - 2]) - (uy[h2 - 1] - uy[h2 - 2]) * (px - ux[h2 - 2]);
        if (cr <= 0.0f) --h2; else break;
      }
      ux[h2] = px, uy[h2] = py;
      ++h2;
    }
  }
  h2 = __shfl_sync(0xffffffffu, h2, 0);
  __syncwarp();
  MinRect r2 = h2 < 3 ? aabb_rect(tx, ty, nu) : min_rect_from_hull_warp(ux, uy, h2, lane);
  float qx[4], qy[4], sside;
  if (!rect_to_ordered_box(r2, qx, qy, &sside)) return false;
  if (sside < min_size + 2.0f) return false;
  if (lane == 0) {
    float width_scale = (float)dw / (float)W, height_scale = (float)dh / (float)H;
    for (int i = 0; i < 4; ++i) {
      out_box[2 * i] = fminf(fmaxf(roundf(qx[i] * width_scale), 0.0f), (float)dw);
      out_box[2 * i + 1] = fminf(fmaxf(roundf(qy[i] * height_scale), 0.0f), (float)dh);
    }
  }
  return true;
}

// one warp per candidate: lane 0 walks the strictly sequential pieces, the lanes share box scoring, the hull sort
// and the min-area-rectangle search
constexpr int GEO_WARPS = 2;
__global__ void __launch_bounds__(GEO_WARPS * 32) db_geometry_kernel(
    const float* __restrict__ pred, int H, int W, const ContourRec* __restrict__ recs, const int* __restrict__ order,
    const int* __restrict__ first, int B, int max_cand, const short2* __restrict__ pool, short2* __restrict__ scratch,
    const int32_t* __restrict__ dest_wh, float box_thresh, float unclip_ratio, float min_size,
    Cand* __restrict__ cands, int* __restrict__ err) {
  __shared__ float s_part[GEO_WARPS][256];
  __shared__ float s_ws[GEO_WARPS][6 * UNCLIP_CAP];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x * GEO_WARPS + wib;
  const int b = blockIdx.y;
  const int cnt = min(first[b + 1] - first[b], max_cand);
  if (rank >= cnt) return;  // warp-uniform
  Cand& out = cands[(size_t)b * max_cand + rank];
  float* ws = s_ws[wib];
  float bx[4] = {0.f, 0.f, 0.f, 0.f}, by[4] = {0.f, 0.f, 0.f, 0.f}, min_side = 0.0f;
  int ok = 0;
  const ContourRec rec = recs[order[first[b] + rank]];
  // get_mini_boxes_from_points: < 3 points -> None; a chain that simplifies to < 3 points is used raw
  const short2* P = pool + rec.off;
  int np = rec.len, nh = -1;
  if (rec.len >= 3) {
    const int ns = simplify_chain_warp(rec, pool, scratch, lane);
    if (ns >= 3) P = scratch + rec.off, np = ns;
    nh = hull_march_warp(P, np, ws, ws + UNCLIP_CAP, err, lane);
  }
  if (lane == 0) {
    out.valid = 0;
    ok = (nh >= 0 && cand_mini_box(P, np, nh, ws, ws + UNCLIP_CAP, bx, by, &min_side)) ? 1 : 0;
    if (ok && min_side < min_size) ok = 0;
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) return;
  for (int i = 0; i < 4; ++i) {
    bx[i] = __shfl_sync(0xffffffffu, bx[i], 0);
    by[i] = __shfl_sync(0xffffffffu, by[i], 0);
  }
  const float score = box_score_fast_warp(pred + (size_t)b * H * W, W, H, bx, by, s_part[wib], lane);
  if (score < box_thresh) return;  // uniform: every lane holds the same score
  const unsigned dw = (unsigned)dest_wh[2 * b], dh = (unsigned)dest_wh[2 * b + 1];
  if (!cand_unclip_box_warp(bx, by, unclip_ratio, min_size, ws, dw, dh, W, H, out.box, err, lane)) return;
  if (lane == 0) {
    out.score = score;
    out.valid = 1;
  }
}

// one warp per image: survivors keep their discovery order (ballot + prefix count per group of 32 candidates).  The
// one-thread-per-image form walked up to max_candidates records through dependent global loads: 0.10 ms on the step's
// critical path for a few KB of output.
__global__ void db_compact_kernel(const Cand* __restrict__ cands, const int* __restrict__ first, int B, int max_cand,
                                  float* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ counts) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int cnt = min(first[b + 1] - first[b], max_cand);
  int n = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    const Cand* c = i < cnt ? &cands[(size_t)b * max_cand + i] : nullptr;
    const bool keep = c && c->valid;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      for (int k = 0; k < 8; ++k) boxes[((size_t)b * max_cand + pos) * 8 + k] = c->box[k];
      scores[(size_t)b * max_cand + pos] = c->score;
    }
    n += __popc(m);
  }
  if (lane == 0) counts[b] = n;
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
DbPostStatus db_postprocess_device(oar_ctx* ctx, const float* pred, int B, int H, int W, const int32_t* h_src_h,
                                   const int32_t* h_src_w, const oar_det_config& cfg, DbPostOut out,
                                   int launch_comps_hint) {
  DbPostStatus status{};
  if (B == 0) return status;
  if (H > 32767 || W > 32767) OAR_FAIL(OAR_E_UNSUPPORTED, "prediction map %dx%d too large", H, W);
  cudaStream_t st = ctx->stream;
  Arena& A = ctx->arena;
  size_t total = (size_t)B * H * W;
  int HW = H * W;
  uint8_t* state = A.get<uint8_t>(total);
  int32_t* lab = A.get<int32_t>(total);
  int32_t* slot_of = A.get<int32_t>(total);
  int comp_cap = (int)std::min<size_t>(total / 4 + 1024, (size_t)64 << 20);
  int rec_cap = comp_cap;
  Comp* comps = A.get<Comp>(comp_cap);
  ContourRec* recs = A.get<ContourRec>(rec_cap);
  unsigned long long pool_cap = total + 4096;
  short2* pool = A.get<short2>(pool_cap);
  short2* scratch = A.get<short2>(pool_cap);
  // counters: [0] n_comps, [1] n_recs, [2] err, [4..5] pool_used (u64)
  int* counters = A.get<int>(8);
  OAR_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(int), st));
  int* n_comps = counters;
  int* n_recs = counters + 1;
  int* err = counters + 2;
  unsigned long long* pool_used = reinterpret_cast<unsigned long long*>(counters + 4);
  int32_t* dest_wh = A.get<int32_t>(2 * B);
  {
    std::vector<int32_t> h(2 * B);
    for (int b = 0; b < B; ++b) h[2 * b] = h_src_w[b], h[2 * b + 1] = h_src_h[b];
    int32_t* staged = (int32_t*)ctx->pinned_get(h.size() * sizeof(int32_t));
    memcpy(staged, h.data(), h.size() * sizeof(int32_t));
    OAR_CUDA(cudaMemcpyAsync(dest_wh, staged, h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  int nb = cdiv(total, 256), nb4 = cdiv(total, 1024);  // one / four pixels per thread
  {
    Launch l(ctx, "db_threshold_init", (double)total, 9.0 * total);
    db_init_kernel<<<nb, 256, 0, st>>>(pred, cfg.thresh, state, lab, total, HW, W);
  }
  {
    Launch l(ctx, "db_ccl_merge", 0, 5.0 * total);
    db_merge_kernel<<<nb4, 256, 0, st>>>(state, lab, B, H, W);
  }
  {
    Launch l(ctx, "db_ccl_flatten", 0, 5.0 * total);
    db_flatten_kernel<<<nb4, 256, 0, st>>>(state, lab, slot_of, comps, n_comps, comp_cap, B, H, W);
  }
  {
    Launch l(ctx, "db_ccl_bbox", 0, 1.0 * total);
    db_bbox_kernel<<<nb4, 256, 0, st>>>(state, lab, slot_of, comps, B, H, W);
  }
  // The component count lives on the device; launch for a generous bound and let
  // surplus warps exit.  (Typical pages have tens of components per image.)
  int launch_comps = std::min(comp_cap, std::max(launch_comps_hint, std::max(4096, B * 2048)));
  {
    const int trace_blocks = std::min(launch_comps, 8192);
    short2* stage = A.get<short2>((size_t)trace_blocks * TRACE_STAGE_POINTS);
    Launch l(ctx, "db_trace_borders", 0, 0);
    db_trace_kernel<<<trace_blocks, TRACE_THREADS, TRACE_TILE_BYTES, st>>>(
        state, lab, comps, n_comps, launch_comps, H, W, pool, pool_used, pool_cap, recs, n_recs, rec_cap, err, stage);
  }
  // sort contour records into discovery order
  int sort_n = std::min(rec_cap, launch_comps * 2);
  unsigned long long* keys = A.get<unsigned long long>(sort_n);
  unsigned long long* keys2 = A.get<unsigned long long>(sort_n);
  int* vals = A.get<int>(sort_n);
  int* vals2 = A.get<int>(sort_n);
  {
    Launch l(ctx, "db_contour_keys");
    db_keys_kernel<<<cdiv(sort_n, 256), 256, 0, st>>>(recs, n_recs, sort_n, keys, vals);
  }
  {
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, sort_n, 0, 64, st);
    void* tmp = A.alloc(tmp_bytes);
    Launch l(ctx, "db_contour_sort");
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, sort_n, 0, 64, st);
  }
  int* first = A.get<int>(B + 1);
  {
    Launch l(ctx, "db_img_first");
    db_img_first_kernel<<<cdiv(B + 1, 64), 64, 0, st>>>(keys2, n_recs, sort_n, B, first);
  }
  Cand* cands = A.get<Cand>((size_t)B * cfg.max_candidates);
  {
    Launch l(ctx, "db_box_geometry");
    db_geometry_kernel<<<dim3(cdiv(cfg.max_candidates, GEO_WARPS), B), GEO_WARPS * 32, 0, st>>>(
        pred, H, W, recs, vals2, first, B, cfg.max_candidates, pool, scratch, dest_wh, cfg.box_thresh,
        cfg.unclip_ratio, cfg.min_size, cands, err);
  }
  {
    Launch l(ctx, "db_compact");
    db_compact_kernel<<<cdiv(B, 4), 128, 0, st>>>(cands, first, B, cfg.max_candidates, out.boxes, out.scores,
                                                   out.counts);
  }
  // counters come back with the results; the caller checks them after its sync
  status.h_counters = (int*)ctx->pinned_get(sizeof(int) * 4);
  status.launch_comps = launch_comps;
  status.sort_n = sort_n;
  OAR_CUDA(cudaMemcpyAsync(status.h_counters, counters, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  return status;
}

// after the stream is synchronised: 0 ok, 1 = re-run with a larger launch bound, throws on hard errors
int db_postprocess_check(const DbPostStatus& s, int* need_comps) {
  if (!s.h_counters) return 0;
  int n_comps = s.h_counters[0], n_recs = s.h_counters[1], err = s.h_counters[2];
  if (err) OAR_FAIL(OAR_E_CAPACITY, "DB post-process scratch capacity exceeded (code %d)", err);
  if (n_comps > s.launch_comps || n_recs > s.sort_n) {
    *need_comps = std::max(n_comps, (n_recs + 1) / 2) + 1024;
    return 1;
  }
  return 0;
}

}  // namespace oar
