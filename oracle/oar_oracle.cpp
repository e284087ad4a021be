// oar_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the oar-ocr (GreatV/oar-ocr @ v0.9.3) det+rec hot path's
// pre/post-processing, written from the reference's Rust sources and, where the
// arithmetic lives in an un-vendored crate, from that crate's published
// algorithm (each function says which).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library;
// the product (oar_ocr_b200/) never links, imports or calls it.
//
// Parity status (see DESIGN.md "Oracle"):
//   * pinned by the reference's own unit-test vectors (tests/test_oracle_kat.py):
//     normalize (normalization.rs:498-709, simd.rs:356-429), CTC argmax/decode
//     (decode.rs:692-758, simd.rs:389-403), mini-box ordering / min side /
//     chain simplification (db_bitmap.rs:375-423), sort_quad_boxes
//     (sorting.rs tests), strict axis-aligned crop predicate
//     (transform.rs:699-716), singular homography error (transform.rs:718-728).
//   * PARITY UNPINNED (third-party crates absent from /root/reference, no
//     golden vectors in the reference): imageproc::find_contours (Suzuki-Abe),
//     clipper2 inflate_paths_d, image::imageops::resize(Triangle),
//     nalgebra LU / 3x3 inverse.  Restated from the published algorithms.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC
//        (f32 semantics must match Rust: no FMA contraction, true f32 ops).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <vector>

namespace {

struct Pt {
  float x, y;
};

// Rust `as u32` / `as usize` from f32: saturating, NaN -> 0.
inline int64_t sat_u(float v) {
  if (!(v > 0.0f)) return 0;
  if (v >= 4294967296.0f) return 4294967295LL;
  return (int64_t)v;
}
inline int64_t sat_usize(float v) {
  if (!(v > 0.0f)) return 0;
  if (v >= 9.2e18f) return INT64_MAX;
  return (int64_t)v;
}

// ---------------------------------------------------------------------------
// resize_detection.rs:243-319  resize_image_type0 (dims only) + ratios
// ---------------------------------------------------------------------------
void det_resize_dims(uint32_t h, uint32_t w, uint32_t limit_side_len, int limit_type /*0 max,1 min,2 long*/,
                     uint32_t max_side_limit, uint32_t* out_h, uint32_t* out_w) {
  float ratio;
  uint32_t mx = std::max(h, w), mn = std::min(h, w);
  if (limit_type == 0) {
    ratio = (mx > limit_side_len) ? (float)limit_side_len / (float)mx : 1.0f;
  } else if (limit_type == 1) {
    ratio = (mn < limit_side_len) ? (float)limit_side_len / (float)mn : 1.0f;
  } else {
    ratio = (float)limit_side_len / (float)mx;
  }
  uint32_t rh = (uint32_t)sat_u((float)h * ratio);
  uint32_t rw = (uint32_t)sat_u((float)w * ratio);
  if (std::max(rh, rw) > max_side_limit) {
    float lr = (float)max_side_limit / (float)std::max(rh, rw);
    rh = (uint32_t)sat_u((float)rh * lr);
    rw = (uint32_t)sat_u((float)rw * lr);
  }
  rh = std::max((rh + 16) / 32 * 32, 32u);
  rw = std::max((rw + 16) / 32 * 32, 32u);
  *out_h = rh;
  *out_w = rw;
}

// ---------------------------------------------------------------------------
// image 0.25 imageops::resize(.., FilterType::Triangle) -- crate absent; the
// published algorithm (sample.rs: vertical_sample then horizontal_sample).
// Call sites: crnn.rs:104-109, resize_detection.rs:314.  PARITY UNPINNED.
// ---------------------------------------------------------------------------
inline float tri_kernel(float x) {
  float a = std::fabs(x);
  return a < 1.0f ? 1.0f - a : 0.0f;
}

// The other filters the reference uses (FilterType::CatmullRom for the PP-DocLayout detectors, Lanczos3 for PicoDet:
// scale_aware_detector.rs:52-80): same two-pass sampler, other kernel and support (image 0.25 sample.rs, EXT).
inline float bc_cubic_spline(float x, float b, float c) {
  float a = std::fabs(x);
  float k;
  if (a < 1.0f)
    k = (12.0f - 9.0f * b - 6.0f * c) * (a * a * a) + (-18.0f + 12.0f * b + 6.0f * c) * (a * a) + (6.0f - 2.0f * b);
  else if (a < 2.0f)
    k = (-b - 6.0f * c) * (a * a * a) + (6.0f * b + 30.0f * c) * (a * a) + (-12.0f * b - 48.0f * c) * a +
        (8.0f * b + 24.0f * c);
  else
    k = 0.0f;
  return k / 6.0f;
}
inline float sinc_f(float t) {
  float a = t * 3.14159265358979323846f;
  return t == 0.0f ? 1.0f : std::sin(a) / a;
}
inline float filter_kernel(int filter, float x) {
  if (filter == 1) return bc_cubic_spline(x, 0.0f, 0.5f);                          // CatmullRom
  if (filter == 2) return std::fabs(x) < 3.0f ? sinc_f(x) * sinc_f(x / 3.0f) : 0.0f;  // Lanczos3
  return tri_kernel(x);
}
inline float filter_support(int filter) { return filter == 1 ? 2.0f : (filter == 2 ? 3.0f : 1.0f); }

void resize_filter_rgb(const uint8_t* src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, uint8_t* dst, int filter);
void resize_triangle_rgb(const uint8_t* src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, uint8_t* dst) {
  resize_filter_rgb(src, w, h, nw, nh, dst, 0);
}

void resize_filter_rgb(const uint8_t* src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, uint8_t* dst, int filter) {
  if (nw == w && nh == h) {
    std::memcpy(dst, src, (size_t)w * h * 3);
    return;
  }
  // vertical pass -> f32 [nh][w][3]
  std::vector<float> tmp((size_t)nh * w * 3);
  {
    float ratio = (float)h / (float)nh;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = filter_support(filter) * sratio;
    std::vector<float> ws;
    for (uint32_t oy = 0; oy < nh; ++oy) {
      float in = ((float)oy + 0.5f) * ratio;
      int64_t left = (int64_t)std::floor(in - support);
      left = std::clamp<int64_t>(left, 0, (int64_t)h - 1);
      int64_t right = (int64_t)std::ceil(in + support);
      right = std::clamp<int64_t>(right, left + 1, (int64_t)h);
      in = in - 0.5f;
      ws.clear();
      float sum = 0.0f;
      for (int64_t i = left; i < right; ++i) {
        float wgt = filter_kernel(filter, ((float)i - in) / sratio);
        ws.push_back(wgt);
        sum += wgt;
      }
      for (auto& v : ws) v /= sum;
      for (uint32_t x = 0; x < w; ++x) {
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
        for (size_t i = 0; i < ws.size(); ++i) {
          const uint8_t* p = src + ((size_t)(left + (int64_t)i) * w + x) * 3;
          t0 += (float)p[0] * ws[i];
          t1 += (float)p[1] * ws[i];
          t2 += (float)p[2] * ws[i];
        }
        float* o = &tmp[((size_t)oy * w + x) * 3];
        o[0] = t0;
        o[1] = t1;
        o[2] = t2;
      }
    }
  }
  // horizontal pass -> u8 [nh][nw][3]
  {
    float ratio = (float)w / (float)nw;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = filter_support(filter) * sratio;
    std::vector<float> ws;
    for (uint32_t ox = 0; ox < nw; ++ox) {
      float in = ((float)ox + 0.5f) * ratio;
      int64_t left = (int64_t)std::floor(in - support);
      left = std::clamp<int64_t>(left, 0, (int64_t)w - 1);
      int64_t right = (int64_t)std::ceil(in + support);
      right = std::clamp<int64_t>(right, left + 1, (int64_t)w);
      in = in - 0.5f;
      ws.clear();
      float sum = 0.0f;
      for (int64_t i = left; i < right; ++i) {
        float wgt = filter_kernel(filter, ((float)i - in) / sratio);
        ws.push_back(wgt);
        sum += wgt;
      }
      for (auto& v : ws) v /= sum;
      for (uint32_t y = 0; y < nh; ++y) {
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
        for (size_t i = 0; i < ws.size(); ++i) {
          const float* p = &tmp[((size_t)y * w + (size_t)(left + (int64_t)i)) * 3];
          t0 += p[0] * ws[i];
          t1 += p[1] * ws[i];
          t2 += p[2] * ws[i];
        }
        uint8_t* o = dst + ((size_t)y * nw + ox) * 3;
        float c0 = std::fmin(std::fmax(t0, 0.0f), 255.0f), c1 = std::fmin(std::fmax(t1, 0.0f), 255.0f),
              c2 = std::fmin(std::fmax(t2, 0.0f), 255.0f);
        o[0] = (uint8_t)std::round(c0);
        o[1] = (uint8_t)std::round(c1);
        o[2] = (uint8_t)std::round(c2);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// imageproc::contours::find_contours::<u32> (Suzuki-Abe border following) --
// crate absent; restated from the published algorithm as used at
// db_bitmap.rs:100.  Returns contours in discovery order.  PARITY UNPINNED.
// ---------------------------------------------------------------------------
struct Contour {
  std::vector<int32_t> xy;  // x0,y0,x1,y1...
  int border_type;          // 0 outer, 1 hole
};

void find_contours(const uint8_t* mask, int width, int height, std::vector<Contour>& out) {
  std::vector<int32_t> v((size_t)width * height);
  for (size_t i = 0; i < v.size(); ++i) v[i] = mask[i] > 0 ? 1 : 0;
  auto at = [&](int x, int y) -> int32_t& { return v[(size_t)y * width + x]; };
  // clockwise ring starting at W
  static const int RX[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
  static const int RY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  auto dir_index = [&](int dx, int dy) {
    for (int i = 0; i < 8; ++i)
      if (RX[i] == dx && RY[i] == dy) return i;
    return 0;
  };
  auto nonzero = [&](int x, int y) { return x >= 0 && y >= 0 && x < width && y < height && at(x, y) != 0; };
  int32_t curr_border_num = 1;
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      if (at(x, y) == 0) continue;
      int adjx = 0, adjy = 0, btype = -1;
      if (at(x, y) == 1 && x > 0 && at(x - 1, y) == 0) {
        adjx = x - 1, adjy = y, btype = 0;
      } else if (at(x, y) > 0 && x + 1 < width && at(x + 1, y) == 0) {
        adjx = x + 1, adjy = y, btype = 1;
      }
      if (btype < 0) continue;
      curr_border_num += 1;
      Contour c;
      c.border_type = btype;
      int start = dir_index(adjx - x, adjy - y);
      // first search: clockwise from the adjacent (zero) pixel direction
      int p1x = 0, p1y = 0;
      bool found = false;
      for (int k = 0; k < 8; ++k) {
        int d = (start + k) & 7;
        if (nonzero(x + RX[d], y + RY[d])) {
          p1x = x + RX[d], p1y = y + RY[d];
          found = true;
          break;
        }
      }
      if (found) {
        int p2x = p1x, p2y = p1y, p3x = x, p3y = y;
        for (;;) {
          c.xy.push_back(p3x);
          c.xy.push_back(p3y);
          int front = dir_index(p2x - p3x, p2y - p3y);
          // counter-clockwise search: ring reversed, starting just before `front`
          int p4x = 0, p4y = 0, d4 = -1;
          for (int k = 1; k <= 8; ++k) {
            int d = (front - k + 16) & 7;  // k=8 -> front itself (last)
            if (nonzero(p3x + RX[d], p3y + RY[d])) {
              p4x = p3x + RX[d], p4y = p3y + RY[d], d4 = d;
              break;
            }
          }
          bool is_right_edge = false;
          for (int k = 1; k <= 8; ++k) {
            int d = (front - k + 16) & 7;
            if (d == d4) break;
            if (RX[d] == 1 && RY[d] == 0) {
              is_right_edge = true;
              break;
            }
          }
          if (p3x + 1 == width || is_right_edge) {
            at(p3x, p3y) = -curr_border_num;
          } else if (at(p3x, p3y) == 1) {
            at(p3x, p3y) = curr_border_num;
          }
          if (p4x == x && p4y == y && p3x == p1x && p3y == p1y) break;
          p2x = p3x, p2y = p3y;
          p3x = p4x, p3y = p4y;
        }
      } else {
        c.xy.push_back(x);
        c.xy.push_back(y);
        at(x, y) = -curr_border_num;
      }
      out.push_back(std::move(c));
    }
  }
}

// ---------------------------------------------------------------------------
// geometry.rs:226-274 convex_hull_from_points (Graham scan)
// ---------------------------------------------------------------------------
inline int total_cmp_f32(float a, float b) {
  int32_t ia, ib;
  std::memcpy(&ia, &a, 4);
  std::memcpy(&ib, &b, 4);
  ia ^= (int32_t)(((uint32_t)(ia >> 31)) >> 1);
  ib ^= (int32_t)(((uint32_t)(ib >> 31)) >> 1);
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

inline float cross3(const Pt& p1, const Pt& p2, const Pt& p3) {
  return (p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x);
}

std::vector<Pt> convex_hull(const std::vector<Pt>& src) {
  if (src.size() < 3) return src;
  std::vector<Pt> pts = src;
  size_t s = 0;
  for (size_t i = 1; i < pts.size(); ++i)
    if (pts[i].y < pts[s].y || (pts[i].y == pts[s].y && pts[i].x < pts[s].x)) s = i;
  std::swap(pts[0], pts[s]);
  Pt sp = pts[0];
  std::stable_sort(pts.begin() + 1, pts.end(), [&](const Pt& a, const Pt& b) {
    float aa = std::atan2(a.y - sp.y, a.x - sp.x);
    float ab = std::atan2(b.y - sp.y, b.x - sp.x);
    int c = total_cmp_f32(aa, ab);
    if (c != 0) return c < 0;
    float dax = a.x - sp.x, day = a.y - sp.y, dbx = b.x - sp.x, dby = b.y - sp.y;
    float da = dax * dax + day * day;
    float db = dbx * dbx + dby * dby;
    return total_cmp_f32(da, db) < 0;
  });
  std::vector<Pt> hull;
  hull.reserve(pts.size());
  for (const Pt& p : pts) {
    while (hull.size() > 1 && cross3(hull[hull.size() - 2], hull[hull.size() - 1], p) <= 0.0f) hull.pop_back();
    hull.push_back(p);
  }
  return hull;
}

// geometry.rs:310-441 get_min_area_rect_from_points
struct MinRect {
  float cx, cy, w, h, angle;
};

MinRect min_area_rect(const std::vector<Pt>& src) {
  MinRect zero{0, 0, 0, 0, 0};
  if (src.size() < 3) return zero;
  std::vector<Pt> hull = convex_hull(src);
  if (hull.size() < 3) {
    float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (const Pt& p : src) {
      if (p.x < mnx) mnx = p.x;
      if (p.x > mxx) mxx = p.x;
      if (p.y < mny) mny = p.y;
      if (p.y > mxy) mxy = p.y;
    }
    if (!std::isfinite(mnx)) return zero;
    return MinRect{(mnx + mxx) * 0.5f, (mny + mxy) * 0.5f, mxx - mnx, mxy - mny, 0.0f};
  }
  const float PI_F = 3.14159265358979323846f;
  float min_area = std::numeric_limits<float>::max();
  MinRect best = zero;
  size_t n = hull.size();
  for (size_t i = 0; i < n; ++i) {
    size_t j = (i + 1) % n;
    float ex = hull[j].x - hull[i].x, ey = hull[j].y - hull[i].y;
    float l2 = ex * ex + ey * ey;
    if (l2 < std::numeric_limits<float>::epsilon()) continue;
    float inv = 1.0f / std::sqrt(l2);
    float nx = ex * inv, ny = ey * inv;
    float px = -ny, py = nx;
    float hix = hull[i].x, hiy = hull[i].y;
    float min_n = std::numeric_limits<float>::max(), max_n = std::numeric_limits<float>::lowest();
    float min_p = std::numeric_limits<float>::max(), max_p = std::numeric_limits<float>::lowest();
    for (const Pt& p : hull) {
      float dx = p.x - hix, dy = p.y - hiy;
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
      float cx = hix + cn * nx + cp * px;
      float cy = hiy + cn * ny + cp * py;
      float ang = std::atan2(ny, nx) * 180.0f / PI_F;
      best = MinRect{cx, cy, width, height, ang};
    }
  }
  return best;
}

// db_bitmap.rs:187-202 box_points_without_reorder + :253-277 paddlex order
void order_mini_box(Pt p[4]) {
  std::stable_sort(p, p + 4, [](const Pt& a, const Pt& b) { return a.x < b.x; });
  int i1, i4, i2, i3;
  if (p[1].y > p[0].y) {
    i1 = 0, i4 = 1;
  } else {
    i1 = 1, i4 = 0;
  }
  if (p[3].y > p[2].y) {
    i2 = 2, i3 = 3;
  } else {
    i2 = 3, i3 = 2;
  }
  Pt o[4] = {p[i1], p[i2], p[i3], p[i4]};
  std::memcpy(p, o, sizeof(o));
}

// db_bitmap.rs:168-185 get_mini_boxes_from_points
bool mini_boxes_from_points(const std::vector<Pt>& pts, Pt out[4], float* min_side) {
  if (pts.size() < 3) return false;
  MinRect r = min_area_rect(pts);
  float ms = std::fmin(r.w, r.h);
  if (!std::isfinite(ms) || ms <= 0.0f) return false;
  const float PI_F = 3.14159265358979323846f;
  float ca = std::cos(r.angle * PI_F / 180.0f);
  float sa = std::sin(r.angle * PI_F / 180.0f);
  float w2 = r.w / 2.0f, h2 = r.h / 2.0f;
  float cxs[4] = {-w2, w2, w2, -w2};
  float cys[4] = {-h2, -h2, h2, h2};
  for (int i = 0; i < 4; ++i) {
    out[i].x = cxs[i] * ca - cys[i] * sa + r.cx;
    out[i].y = cxs[i] * sa + cys[i] * ca + r.cy;
  }
  order_mini_box(out);
  *min_side = ms;
  return true;
}

// db_bitmap.rs:207-249 simplify_chain_points
inline int sign_step(float v) { return v > 0.0f ? 1 : (v < 0.0f ? -1 : 0); }

std::vector<Pt> simplify_chain(const std::vector<Pt>& pts) {
  if (pts.size() <= 2) return pts;
  std::vector<Pt> s;
  size_t n = pts.size();
  for (size_t i = 0; i < n; ++i) {
    Pt prev = pts[(i + n - 1) % n], cur = pts[i], next = pts[(i + 1) % n];
    int a0 = sign_step(cur.x - prev.x), a1 = sign_step(cur.y - prev.y);
    int b0 = sign_step(next.x - cur.x), b1 = sign_step(next.y - cur.y);
    if (a0 != b0 || a1 != b1) s.push_back(cur);
  }
  if (s.size() < 3) return pts;
  return s;
}

// db_bitmap.rs:153-165 get_mini_boxes_from_contour
bool mini_boxes_from_contour(const Contour& c, Pt out[4], float* min_side) {
  std::vector<Pt> pts(c.xy.size() / 2);
  for (size_t i = 0; i < pts.size(); ++i) pts[i] = Pt{(float)c.xy[2 * i], (float)c.xy[2 * i + 1]};
  std::vector<Pt> s = simplify_chain(pts);
  if (s.size() >= 3) return mini_boxes_from_points(s, out, min_side);
  return mini_boxes_from_points(pts, out, min_side);
}

// ---------------------------------------------------------------------------
// db_score.rs:34-134 box_score_fast + geometry.rs:1087-1164 process_scanline
// ---------------------------------------------------------------------------
float box_score_fast(const float* pred, int width, int height, const Pt* box, int npts) {
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  if (npts == 0) mnx = mny = mxx = mxy = 0.0f;
  for (int i = 0; i < npts; ++i) {
    if (box[i].x < mnx) mnx = box[i].x;
    if (box[i].x > mxx) mxx = box[i].x;
    if (box[i].y < mny) mny = box[i].y;
    if (box[i].y > mxy) mxy = box[i].y;
  }
  float fminx = std::fmin(std::fmax(std::floor(mnx), 0.0f), (float)width - 1.0f);
  float fmaxx = std::fmin(std::fmax(std::ceil(mxx), 0.0f), (float)width - 1.0f);
  float fminy = std::fmin(std::fmax(std::floor(mny), 0.0f), (float)height - 1.0f);
  float fmaxy = std::fmin(std::fmax(std::ceil(mxy), 0.0f), (float)height - 1.0f);
  int64_t start_y = sat_usize(fminy), end_y = sat_usize(fmaxy) + 1;
  int64_t start_x = sat_usize(fminx), end_x = sat_usize(fmaxx) + 1;
  float total = 0.0f;
  int64_t total_px = 0;
  std::vector<float> xs;
  for (int64_t yy = start_y; yy < end_y; ++yy) {
    float y = (float)yy + 0.5f;
    xs.clear();
    for (int i = 0; i < npts; ++i) {
      int j = (i + 1) % npts;
      const Pt &p1 = box[i], &p2 = box[j];
      if (((p1.y <= y && y < p2.y) || (p2.y <= y && y < p1.y)) &&
          std::fabs(p2.y - p1.y) > std::numeric_limits<float>::epsilon()) {
        float x = p1.x + (y - p1.y) * (p2.x - p1.x) / (p2.y - p1.y);
        xs.push_back(x);
      }
    }
    std::stable_sort(xs.begin(), xs.end(), [](float a, float b) { return a < b; });
    float line = 0.0f;
    int64_t line_px = 0;
    int64_t yi = sat_usize(y);
    if (yi < height) {
      const float* row = pred + (size_t)yi * width;
      for (size_t k = 0; k + 1 < xs.size(); k += 2) {
        int64_t x1 = sat_usize(std::fmax(xs[k], (float)start_x));
        int64_t x2 = sat_usize(std::fmin(xs[k + 1], (float)end_x));
        if (x1 < x2 && x1 >= start_x && x2 <= end_x) {
          int64_t xe = std::min<int64_t>(x2, width);
          if (x1 < xe) {
            for (int64_t x = x1; x < xe; ++x) line += row[x];
            line_px += xe - x1;
          }
        }
      }
    }
    total += line;
    total_px += line_px;
  }
  return total_px > 0 ? total / (float)total_px : 0.0f;
}

// ---------------------------------------------------------------------------
// db_bitmap.rs:279-368 unclip.  clipper2-rust 1.0.3 (port of Clipper2's
// ClipperOffset) is absent; restated from Clipper2's published offsetting
// algorithm (InflatePaths -> ClipperOffset::DoGroupOffset/OffsetPolygon/
// OffsetPoint/DoRound, precision 2 => x100 integer grid).  The final
// Clipper64 union (positive fill) of a simple outward offset of a convex
// polygon only drops collinear/duplicate vertices and may rotate the start
// vertex; the caller feeds the vertices to a convex hull, which is invariant
// to both, so the union is not restated.  PARITY UNPINNED.
// ---------------------------------------------------------------------------
struct PD {
  double x, y;
};

bool unclip(const Pt* box, int npts, float unclip_ratio, std::vector<Pt>& out) {
  out.clear();
  if (npts < 3) {
    for (int i = 0; i < npts; ++i) out.push_back(box[i]);
    return true;
  }
  std::vector<PD> path(npts);
  for (int i = 0; i < npts; ++i) path[i] = PD{(double)box[i].x, (double)box[i].y};
  // clipper2 Area(): a += (prev.y + cur.y) * (prev.x - cur.x); a * 0.5
  double a = 0.0;
  {
    PD prev = path[npts - 1];
    for (int i = 0; i < npts; ++i) {
      a += (prev.y + path[i].y) * (prev.x - path[i].x);
      prev = path[i];
    }
    a *= 0.5;
  }
  double polygon_area = std::fabs(a);
  if (polygon_area <= std::numeric_limits<double>::epsilon()) return false;
  double perimeter = 0.0;
  {
    const PD* p1 = &path[0];
    for (int i = 1; i < npts; ++i) {
      perimeter += std::hypot(path[i].x - p1->x, path[i].y - p1->y);
      p1 = &path[i];
    }
    perimeter += std::hypot(path[0].x - p1->x, path[0].y - p1->y);
  }
  if (perimeter <= std::numeric_limits<double>::epsilon()) return false;
  double delta = polygon_area * (double)unclip_ratio / perimeter;
  if (std::fabs(delta) <= std::numeric_limits<double>::epsilon()) return false;

  const double scale = 100.0;  // precision = 2
  struct P64 {
    int64_t x, y;
  };
  std::vector<P64> ip;
  ip.reserve(npts);
  for (int i = 0; i < npts; ++i) {
    P64 q{(int64_t)std::llround(path[i].x * scale), (int64_t)std::llround(path[i].y * scale)};
    if (!ip.empty() && ip.back().x == q.x && ip.back().y == q.y) continue;  // StripDuplicates (closed)
    ip.push_back(q);
  }
  if (ip.size() > 1 && ip.front().x == ip.back().x && ip.front().y == ip.back().y) ip.pop_back();
  size_t n = ip.size();
  if (n < 3) return false;
  double d = delta * scale;
  // orientation of the (single, lowest) path decides the sign of group_delta
  double ia = 0.0;
  {
    P64 prev = ip[n - 1];
    for (size_t i = 0; i < n; ++i) {
      ia += (double)(prev.y + ip[i].y) * (double)(prev.x - ip[i].x);
      prev = ip[i];
    }
  }
  bool reversed = ia < 0.0;
  double group_delta = reversed ? -d : d;
  double abs_delta = std::fabs(group_delta);
  const double PI_D = 3.141592653589793238;
  double arc_tol = std::log10(2.0 + abs_delta) * 0.25;
  double steps_per_360 = std::min(PI_D / std::acos(1.0 - arc_tol / abs_delta), abs_delta * PI_D);
  double step_sin = std::sin(2.0 * PI_D / steps_per_360);
  double step_cos = std::cos(2.0 * PI_D / steps_per_360);
  if (group_delta < 0.0) step_sin = -step_sin;
  double steps_per_rad = steps_per_360 / (2.0 * PI_D);
  // normals
  std::vector<PD> norms(n);
  for (size_t i = 0; i < n; ++i) {
    const P64 &p1 = ip[i], &p2 = ip[(i + 1) % n];
    double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
    if (dx == 0 && dy == 0) {
      norms[i] = PD{0, 0};
      continue;
    }
    double inv = 1.0 / std::hypot(dx, dy);
    dx *= inv;
    dy *= inv;
    norms[i] = PD{dy, -dx};
  }
  std::vector<P64> po;
  auto push_d = [&](double x, double y) { po.push_back(P64{(int64_t)std::llround(x), (int64_t)std::llround(y)}); };
  for (size_t j = 0, k = n - 1; j < n; k = j, ++j) {
    // Clipper2 offset: CrossProduct(v1, v2) = v1.y*v2.x - v2.y*v1.x, called as (norms[j], norms[k])
    double sin_a = norms[j].y * norms[k].x - norms[k].y * norms[j].x;
    double cos_a = norms[j].x * norms[k].x + norms[j].y * norms[k].y;
    if (sin_a > 1.0) sin_a = 1.0;
    else if (sin_a < -1.0) sin_a = -1.0;
    const P64& pt = ip[j];
    if (cos_a > -0.999 && (sin_a * group_delta < 0)) {
      // concave join
      push_d((double)pt.x + norms[k].x * group_delta, (double)pt.y + norms[k].y * group_delta);
      po.push_back(pt);
      push_d((double)pt.x + norms[j].x * group_delta, (double)pt.y + norms[j].y * group_delta);
    } else {
      // JoinType::Round -> DoRound(path, j, k, atan2(sin_a, cos_a))
      double angle = std::atan2(sin_a, cos_a);
      double ox = norms[k].x * group_delta, oy = norms[k].y * group_delta;
      push_d((double)pt.x + ox, (double)pt.y + oy);
      int steps = (int)std::ceil(steps_per_rad * std::fabs(angle));
      for (int i = 1; i < steps; ++i) {
        double nx2 = ox * step_cos - step_sin * oy;
        double ny2 = ox * step_sin + oy * step_cos;
        ox = nx2;
        oy = ny2;
        push_d((double)pt.x + ox, (double)pt.y + oy);
      }
      push_d((double)pt.x + norms[j].x * group_delta, (double)pt.y + norms[j].y * group_delta);
    }
  }
  // back to f64 /100 then f32 (db_bitmap.rs:349-352)
  out.reserve(po.size());
  for (const P64& q : po) {
    double x = (double)q.x * (1.0 / scale), y = (double)q.y * (1.0 / scale);
    out.push_back(Pt{(float)x, (float)y});
  }
  if (out.size() > 1) {
    const Pt &f = out.front(), &l = out.back();
    if (std::fabs(f.x - l.x) < std::numeric_limits<float>::epsilon() &&
        std::fabs(f.y - l.y) < std::numeric_limits<float>::epsilon())
      out.pop_back();
  }
  if (out.size() < 3) {
    out.clear();
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------
// transform.rs:212-283 get_perspective_transform (nalgebra f32 LU, partial
// pivoting) + :312-316 Matrix3::try_inverse.  nalgebra absent; restated from
// its published LU/solve/inverse routines.  PARITY UNPINNED.
// ---------------------------------------------------------------------------
bool perspective_transform(const Pt src[4], const Pt dst[4], float M[9]) {
  float a[8][8];
  float b[8];
  std::memset(a, 0, sizeof(a));
  for (int i = 0; i < 4; ++i) {
    float sx = src[i].x, sy = src[i].y, dx = dst[i].x, dy = dst[i].y;
    float r0[8] = {sx, sy, 1.0f, 0.0f, 0.0f, 0.0f, -sx * dx, -sy * dx};
    float r1[8] = {0.0f, 0.0f, 0.0f, sx, sy, 1.0f, -sx * dy, -sy * dy};
    std::memcpy(a[2 * i], r0, sizeof(r0));
    std::memcpy(a[2 * i + 1], r1, sizeof(r1));
    b[2 * i] = dx;
    b[2 * i + 1] = dy;
  }
  int perm_a[8], perm_b[8], nperm = 0;
  for (int i = 0; i < 8; ++i) {
    int piv = i;
    float mx = std::fabs(a[i][i]);
    for (int r = i + 1; r < 8; ++r) {
      float v = std::fabs(a[r][i]);
      if (v > mx) {
        mx = v;
        piv = r;
      }
    }
    float diag = a[piv][i];
    if (diag == 0.0f) continue;
    if (piv != i) {
      perm_a[nperm] = i;
      perm_b[nperm] = piv;
      ++nperm;
      for (int c = 0; c < 8; ++c) std::swap(a[i][c], a[piv][c]);
    }
    float inv_diag = 1.0f / diag;
    for (int r = i + 1; r < 8; ++r) a[r][i] *= inv_diag;
    for (int c = i + 1; c < 8; ++c) {
      float pv = -a[i][c];
      for (int r = i + 1; r < 8; ++r) a[r][c] = pv * a[r][i] + a[r][c];
    }
  }
  for (int k = 0; k < nperm; ++k) std::swap(b[perm_a[k]], b[perm_b[k]]);
  // L (unit diag) forward substitution, column oriented
  for (int i = 0; i < 7; ++i) {
    float coeff = -b[i];
    for (int r = i + 1; r < 8; ++r) b[r] = coeff * a[r][i] + b[r];
  }
  // U back substitution, column oriented
  for (int i = 7; i >= 0; --i) {
    float diag = a[i][i];
    if (diag == 0.0f) return false;
    float coeff = b[i] / diag;
    b[i] = coeff;
    float nc = -coeff;
    for (int r = 0; r < i; ++r) b[r] = nc * a[r][i] + b[r];
  }
  for (int i = 0; i < 8; ++i) M[i] = b[i];
  M[8] = 1.0f;
  return true;
}

bool invert3(const float m[9], float inv[9]) {
  float m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
  float minor_m12_m23 = m22 * m33 - m32 * m23;
  float minor_m11_m23 = m21 * m33 - m31 * m23;
  float minor_m11_m22 = m21 * m32 - m31 * m22;
  float det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
  if (det == 0.0f) return false;
  inv[0] = minor_m12_m23 / det;
  inv[1] = (m13 * m32 - m33 * m12) / det;
  inv[2] = (m12 * m23 - m22 * m13) / det;
  inv[3] = -minor_m11_m23 / det;
  inv[4] = (m11 * m33 - m31 * m13) / det;
  inv[5] = (m13 * m21 - m23 * m11) / det;
  inv[6] = minor_m11_m22 / det;
  inv[7] = (m12 * m31 - m32 * m11) / det;
  inv[8] = (m11 * m22 - m21 * m12) / det;
  return true;
}

// transform.rs:411-502
inline float cubic_kernel(float t) {
  const float A = -0.5f;
  float ta = std::fabs(t);
  if (ta <= 1.0f) return (A + 2.0f) * ta * ta * ta - (A + 3.0f) * ta * ta + 1.0f;
  if (ta < 2.0f) return A * ta * ta * ta - 5.0f * A * ta * ta + 8.0f * A * ta - 4.0f * A;
  return 0.0f;
}

void bicubic(const uint8_t* raw, int w, int h, float x, float y, uint8_t out[3]) {
  float fx = std::floor(x), fy = std::floor(y);
  // Rust `as i32` saturates
  int xi = (fx >= 2147483648.0f) ? INT32_MAX : (fx <= -2147483648.0f ? INT32_MIN : (int)fx);
  int yi = (fy >= 2147483648.0f) ? INT32_MAX : (fy <= -2147483648.0f ? INT32_MIN : (int)fy);
  if (fx != fx) xi = 0;
  if (fy != fy) yi = 0;
  float dx = x - (float)xi, dy = y - (float)yi;
  float wx[4] = {cubic_kernel(dx + 1.0f), cubic_kernel(dx), cubic_kernel(dx - 1.0f), cubic_kernel(dx - 2.0f)};
  float wy[4] = {cubic_kernel(dy + 1.0f), cubic_kernel(dy), cubic_kernel(dy - 1.0f), cubic_kernel(dy - 2.0f)};
  size_t stride = (size_t)w * 3;
  auto cl = [](int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); };
  size_t cx[4], cy[4];
  for (int i = 0; i < 4; ++i) {
    cx[i] = (size_t)cl((int64_t)xi - 1 + i, 0, w - 1) * 3;
    cy[i] = (size_t)cl((int64_t)yi - 1 + i, 0, h - 1) * stride;
  }
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
  for (int j = 0; j < 4; ++j) {
    for (int i = 0; i < 4; ++i) {
      float wgt = wx[i] * wy[j];
      size_t idx = cy[j] + cx[i];
      r0 += wgt * (float)raw[idx];
      r1 += wgt * (float)raw[idx + 1];
      r2 += wgt * (float)raw[idx + 2];
    }
  }
  out[0] = (uint8_t)std::fmin(std::fmax(std::round(r0), 0.0f), 255.0f);
  out[1] = (uint8_t)std::fmin(std::fmax(std::round(r1), 0.0f), 255.0f);
  out[2] = (uint8_t)std::fmin(std::fmax(std::round(r2), 0.0f), 255.0f);
}

}  // namespace

extern "C" {

// resize_detection.rs:243-319
void oracle_det_resize_dims(uint32_t h, uint32_t w, uint32_t limit, int limit_type, uint32_t max_side, uint32_t* oh,
                            uint32_t* ow) {
  det_resize_dims(h, w, limit, limit_type, max_side, oh, ow);
}

// filter: 0 Triangle, 1 CatmullRom, 2 Lanczos3
void oracle_resize_filter(const uint8_t* src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, uint8_t* dst, int filter) {
  resize_filter_rgb(src, w, h, nw, nh, dst, filter);
}

void oracle_resize_triangle(const uint8_t* src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, uint8_t* dst) {
  resize_triangle_rgb(src, w, h, nw, nh, dst);
}

// simd.rs:87-104 normalize_chw_scalar (== the SIMD path bit for bit, simd.rs:11-14)
void oracle_normalize_chw(const uint8_t* rgb, int width, int height, const int src_ch[3], const float alpha[3],
                          const float beta[3], float* out) {
  size_t plane = (size_t)width * height;
  for (int c = 0; c < 3; ++c) {
    float a = alpha[c], b = beta[c];
    int sc = src_ch[c];
    float* dst = out + c * plane;
    for (size_t p = 0; p < plane; ++p) dst[p] = (float)rgb[p * 3 + sc] * a + b;
  }
}

// simd.rs:107-123 normalize_hwc_scalar
void oracle_normalize_hwc(const uint8_t* rgb, int width, int height, const int src_ch[3], const float alpha[3],
                          const float beta[3], float* out) {
  size_t plane = (size_t)width * height;
  for (size_t p = 0; p < plane; ++p)
    for (int c = 0; c < 3; ++c) out[p * 3 + c] = (float)rgb[p * 3 + src_ch[c]] * alpha[c] + beta[c];
}

// normalization.rs:142-143
void oracle_norm_coeffs(float scale, const float mean[3], const float stdv[3], float alpha[3], float beta[3]) {
  for (int i = 0; i < 3; ++i) {
    alpha[i] = scale / stdv[i];
    beta[i] = -mean[i] / stdv[i];
  }
}

// db_postprocess.rs:185-221
void oracle_threshold_mask(const float* pred, int n, float thresh, uint8_t* mask) {
  for (int i = 0; i < n; ++i) mask[i] = pred[i] > thresh ? 255 : 0;
}

// find_contours -> flattened.  Returns number of contours; offsets has n+1 entries
// (in points).  If xy == nullptr only counts are returned through *total_pts.
int oracle_find_contours(const uint8_t* mask, int width, int height, int32_t* xy, int64_t xy_cap, int32_t* offsets,
                         int32_t* border_types, int max_contours, int64_t* total_pts) {
  std::vector<Contour> cs;
  find_contours(mask, width, height, cs);
  int64_t tot = 0;
  for (auto& c : cs) tot += (int64_t)c.xy.size() / 2;
  *total_pts = tot;
  if (!xy) return (int)cs.size();
  int n = std::min<int>((int)cs.size(), max_contours);
  int64_t off = 0;
  for (int i = 0; i < n; ++i) {
    offsets[i] = (int32_t)off;
    border_types[i] = cs[i].border_type;
    int64_t np = (int64_t)cs[i].xy.size() / 2;
    if ((off + np) * 2 > xy_cap) {
      n = i;
      break;
    }
    std::memcpy(xy + off * 2, cs[i].xy.data(), cs[i].xy.size() * sizeof(int32_t));
    off += np;
  }
  offsets[n] = (int32_t)off;
  return n;
}

// db_bitmap.rs:207-239 (KAT access)
int oracle_simplify_chain(const float* xy, int n, float* out) {
  std::vector<Pt> p(n);
  for (int i = 0; i < n; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  auto s = simplify_chain(p);
  for (size_t i = 0; i < s.size(); ++i) {
    out[2 * i] = s[i].x;
    out[2 * i + 1] = s[i].y;
  }
  return (int)s.size();
}

// db_bitmap.rs:253-277 (KAT access)
void oracle_order_mini_box(float* xy) {
  Pt p[4];
  for (int i = 0; i < 4; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  order_mini_box(p);
  for (int i = 0; i < 4; ++i) {
    xy[2 * i] = p[i].x;
    xy[2 * i + 1] = p[i].y;
  }
}

// db_bitmap.rs:168-185 (KAT access).  returns 1 on Some
int oracle_mini_boxes_from_points(const float* xy, int n, float* out_xy, float* min_side) {
  std::vector<Pt> p(n);
  for (int i = 0; i < n; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  Pt o[4];
  if (!mini_boxes_from_points(p, o, min_side)) return 0;
  for (int i = 0; i < 4; ++i) {
    out_xy[2 * i] = o[i].x;
    out_xy[2 * i + 1] = o[i].y;
  }
  return 1;
}

// geometry.rs:310-441 (test access): out = cx,cy,w,h,angle
void oracle_min_area_rect(const float* xy, int n, float* out) {
  std::vector<Pt> p(n);
  for (int i = 0; i < n; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  MinRect r = min_area_rect(p);
  out[0] = r.cx, out[1] = r.cy, out[2] = r.w, out[3] = r.h, out[4] = r.angle;
}

float oracle_box_score_fast(const float* pred, int width, int height, const float* xy, int n) {
  std::vector<Pt> p(n);
  for (int i = 0; i < n; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  return box_score_fast(pred, width, height, p.data(), n);
}

// db_bitmap.rs:279-368.  returns number of points (0 => empty)
int oracle_unclip(const float* xy, int n, float ratio, float* out, int cap) {
  std::vector<Pt> p(n), o;
  for (int i = 0; i < n; ++i) p[i] = Pt{xy[2 * i], xy[2 * i + 1]};
  if (!unclip(p.data(), n, ratio, o)) return 0;
  int m = std::min<int>((int)o.size(), cap);
  for (int i = 0; i < m; ++i) {
    out[2 * i] = o[i].x;
    out[2 * i + 1] = o[i].y;
  }
  return m;
}

// db_bitmap.rs:84-150 boxes_from_bitmap (+ db_postprocess.rs:134-179 process with
// use_dilation=false, ScoreMode::Fast, BoxType::Quad as both adapters configure,
// text_detection_adapter.rs:165-173).  boxes: [n][4][2] rounded+clamped coords;
// raw (optional): the same coordinates before round/clamp, for tie analysis.
int oracle_db_postprocess(const float* pred, int width, int height, uint32_t dest_w, uint32_t dest_h, float thresh,
                          float box_thresh, float unclip_ratio, int max_candidates, float min_size, float* boxes,
                          float* scores, float* raw, int cap) {
  std::vector<uint8_t> mask((size_t)width * height);
  oracle_threshold_mask(pred, width * height, thresh, mask.data());
  std::vector<Contour> cs;
  find_contours(mask.data(), width, height, cs);
  float width_scale = (float)dest_w / (float)width;
  float height_scale = (float)dest_h / (float)height;
  float dwf = (float)dest_w, dhf = (float)dest_h;
  int n = 0;
  size_t lim = std::min<size_t>(cs.size(), (size_t)max_candidates);
  for (size_t ci = 0; ci < lim && n < cap; ++ci) {
    Pt mb[4];
    float min_side;
    if (!mini_boxes_from_contour(cs[ci], mb, &min_side)) continue;
    if (min_side < min_size) continue;
    float score = box_score_fast(pred, width, height, mb, 4);
    if (score < box_thresh) continue;
    std::vector<Pt> un;
    if (!unclip(mb, 4, unclip_ratio, un)) continue;
    if (un.empty()) continue;
    Pt bp[4];
    float sside;
    if (!mini_boxes_from_points(un, bp, &sside)) continue;
    if (sside < min_size + 2.0f) continue;
    for (int i = 0; i < 4; ++i) {
      float rx = bp[i].x * width_scale, ry = bp[i].y * height_scale;
      if (raw) {
        raw[(size_t)n * 8 + 2 * i] = rx;
        raw[(size_t)n * 8 + 2 * i + 1] = ry;
      }
      float x = std::fmin(std::fmax(std::round(rx), 0.0f), dwf);
      float y = std::fmin(std::fmax(std::round(ry), 0.0f), dhf);
      boxes[(size_t)n * 8 + 2 * i] = x;
      boxes[(size_t)n * 8 + 2 * i + 1] = y;
    }
    scores[n] = score;
    ++n;
  }
  return n;
}

// sorting.rs:35-84.  boxes [n][4][2] in place; order[] receives the permutation.
void oracle_sort_quad_boxes(float* boxes, int n, int32_t* order) {
  struct B {
    float p[8];
    int idx;
    float ymin, xmin;
  };
  std::vector<B> v(n);
  for (int i = 0; i < n; ++i) {
    std::memcpy(v[i].p, boxes + (size_t)i * 8, 32);
    v[i].idx = i;
    float ym = INFINITY, xm = INFINITY;
    for (int k = 0; k < 4; ++k) {
      if (v[i].p[2 * k] < xm) xm = v[i].p[2 * k];
      if (v[i].p[2 * k + 1] < ym) ym = v[i].p[2 * k + 1];
    }
    v[i].ymin = ym;
    v[i].xmin = xm;
  }
  std::stable_sort(v.begin(), v.end(), [](const B& a, const B& b) {
    if (a.ymin < b.ymin) return true;
    if (a.ymin > b.ymin) return false;
    if (a.ymin == b.ymin) return a.xmin < b.xmin;
    return false;
  });
  for (int i = 0; i + 1 < n; ++i) {
    for (int j = i; j >= 0; --j) {
      if (j + 1 >= n) break;
      if (std::fabs(v[j + 1].ymin - v[j].ymin) < 10.0f && v[j + 1].xmin < v[j].xmin) {
        std::swap(v[j], v[j + 1]);
      } else {
        break;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    std::memcpy(boxes + (size_t)i * 8, v[i].p, 32);
    if (order) order[i] = v[i].idx;
  }
}

// transform.rs:27-38 (KAT access)
int oracle_is_exact_axis_aligned(const float* xy, uint32_t width, uint32_t height) {
  float w = (float)width, h = (float)height;
  return xy[0] == 0.0f && xy[1] == 0.0f && xy[2] == w && xy[3] == 0.0f && xy[4] == w && xy[5] == h && xy[6] == 0.0f &&
         xy[7] == h;
}

// transform.rs:212-283 (KAT access) returns 1 on success
int oracle_perspective_transform(const float* src, const float* dst, float* M) {
  Pt s[4], d[4];
  for (int i = 0; i < 4; ++i) {
    s[i] = Pt{src[2 * i], src[2 * i + 1]};
    d[i] = Pt{dst[2 * i], dst[2 * i + 1]};
  }
  return perspective_transform(s, d, M) ? 1 : 0;
}

// transform.rs:439-502 (KAT access)
void oracle_bicubic(const uint8_t* raw, int w, int h, float x, float y, uint8_t* out) { bicubic(raw, w, h, x, y, out); }

// transform.rs:76-191 get_rotate_crop_image.  Two-phase: call with out==nullptr to
// get dims (returns 0 ok / nonzero = Err), then with a buffer of ow*oh*3.
int oracle_rotate_crop(const uint8_t* img, int W, int H, const float* quad, uint8_t* out, int* ow, int* oh) {
  float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
  for (int i = 0; i < 4; ++i) {
    mnx = std::fmin(mnx, quad[2 * i]);
    mxx = std::fmax(mxx, quad[2 * i]);
    mny = std::fmin(mny, quad[2 * i + 1]);
    mxy = std::fmax(mxy, quad[2 * i + 1]);
  }
  uint32_t left = (uint32_t)sat_u(std::fmax(mnx, 0.0f));
  uint32_t top = (uint32_t)sat_u(std::fmax(mny, 0.0f));
  uint32_t right = (uint32_t)sat_u(std::fmin(mxx, (float)W));
  uint32_t bottom = (uint32_t)sat_u(std::fmin(mxy, (float)H));
  if (right <= left || bottom <= top) return 1;
  uint32_t cw = right - left, ch = bottom - top;
  Pt pts[4];
  for (int i = 0; i < 4; ++i) pts[i] = Pt{quad[2 * i] - (float)left, quad[2 * i + 1] - (float)top};
  Pt sorted[4];
  std::memcpy(sorted, pts, sizeof(pts));
  std::stable_sort(sorted, sorted + 4, [](const Pt& a, const Pt& b) { return a.x < b.x; });
  int ia = 0, id = 1;
  if (sorted[1].y < sorted[0].y) ia = 1, id = 0;
  int ib = 2, ic = 3;
  if (sorted[3].y < sorted[2].y) ib = 3, ic = 2;
  Pt ord[4] = {sorted[ia], sorted[ib], sorted[ic], sorted[id]};
  float oxy[8] = {ord[0].x, ord[0].y, ord[1].x, ord[1].y, ord[2].x, ord[2].y, ord[3].x, ord[3].y};
  std::vector<uint8_t> res;
  uint32_t rw, rh;
  if (oracle_is_exact_axis_aligned(oxy, cw, ch)) {
    rw = cw, rh = ch;
    if (out || true) {
      res.resize((size_t)rw * rh * 3);
      for (uint32_t y = 0; y < ch; ++y)
        std::memcpy(&res[(size_t)y * cw * 3], img + ((size_t)(top + y) * W + left) * 3, (size_t)cw * 3);
    }
  } else {
    auto dist = [](const Pt& a, const Pt& b) { return std::hypot(a.x - b.x, a.y - b.y); };
    float w1 = dist(ord[0], ord[1]), w2 = dist(ord[2], ord[3]);
    uint32_t icw = (uint32_t)sat_u(std::round(std::fmax(w1, w2)));
    float h1 = dist(ord[0], ord[3]), h2 = dist(ord[1], ord[2]);
    uint32_t ich = (uint32_t)sat_u(std::round(std::fmax(h1, h2)));
    if (icw == 0 || ich == 0) return 2;
    Pt stdp[4] = {{0.0f, 0.0f}, {(float)icw, 0.0f}, {(float)icw, (float)ich}, {0.0f, (float)ich}};
    float M[9], inv[9];
    if (!perspective_transform(ord, stdp, M)) return 3;
    if (!invert3(M, inv)) return 4;
    rw = icw, rh = ich;
    res.resize((size_t)rw * rh * 3);
    // source = the cropped sub-image (crop_imm .to_image()): replicate borders of the CROP
    std::vector<uint8_t> crop((size_t)cw * ch * 3);
    for (uint32_t y = 0; y < ch; ++y)
      std::memcpy(&crop[(size_t)y * cw * 3], img + ((size_t)(top + y) * W + left) * 3, (size_t)cw * 3);
    for (uint32_t dy = 0; dy < rh; ++dy) {
      for (uint32_t dx = 0; dx < rw; ++dx) {
        float fx = (float)dx, fy = (float)dy;
        // nalgebra gemv: column-major axpy accumulation, v = (x, y, 1)
        float sx = inv[0] * fx;
        sx = inv[1] * fy + sx;
        sx = inv[2] * 1.0f + sx;
        float sy = inv[3] * fx;
        sy = inv[4] * fy + sy;
        sy = inv[5] * 1.0f + sy;
        float sz = inv[6] * fx;
        sz = inv[7] * fy + sz;
        sz = inv[8] * 1.0f + sz;
        uint8_t* o = &res[((size_t)dy * rw + dx) * 3];
        if (std::fabs(sz) > std::numeric_limits<float>::epsilon()) {
          bicubic(crop.data(), (int)cw, (int)ch, sx / sz, sy / sz, o);
        } else {
          o[0] = crop[0], o[1] = crop[1], o[2] = crop[2];
        }
      }
    }
  }
  // orient_vertical_crop transform.rs:40-51 (imageops::rotate270)
  if ((float)rh >= (float)rw * 1.5f) {
    *ow = (int)rh;
    *oh = (int)rw;
    if (out) {
      for (uint32_t y = 0; y < rh; ++y)
        for (uint32_t x = 0; x < rw; ++x) {
          // out.put_pixel(y, rw-1-x, p)
          size_t o = ((size_t)(rw - 1 - x) * rh + y) * 3;
          const uint8_t* s = &res[((size_t)y * rw + x) * 3];
          out[o] = s[0], out[o + 1] = s[1], out[o + 2] = s[2];
        }
    }
  } else {
    *ow = (int)rw;
    *oh = (int)rh;
    if (out) std::memcpy(out, res.data(), res.size());
  }
  return 0;
}

// crnn.rs:79-87 tensor width for a chunk
int oracle_crnn_tensor_width(const int32_t* ws, const int32_t* hs, int n, int img_h, int img_w, int max_img_w) {
  float base = (float)img_w / (float)std::max(img_h, 1);
  float mx = base;
  for (int i = 0; i < n; ++i) {
    float r = (float)ws[i] / (float)std::max(hs[i], 1);
    mx = std::fmax(mx, r);
  }
  int64_t tw = sat_usize((float)img_h * mx);
  return (int)std::min<int64_t>(tw, max_img_w);
}

// crnn.rs:100-120 + simd.rs:248-308: one crop into its [3][img_h][tensor_w] slice (pre-zeroed)
int oracle_crnn_preprocess_one(const uint8_t* crop, int w, int h, int img_h, int tensor_w, float* dst) {
  float ratio = (float)w / (float)h;
  int64_t rw = std::min<int64_t>(sat_usize(std::ceil((float)img_h * ratio)), tensor_w);
  std::vector<uint8_t> rs((size_t)rw * img_h * 3);
  if (rw > 0) resize_triangle_rgb(crop, (uint32_t)w, (uint32_t)h, (uint32_t)rw, (uint32_t)img_h, rs.data());
  size_t plane = (size_t)img_h * tensor_w;
  const int SRC[3] = {2, 1, 0};
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < img_h; ++y)
      for (int64_t x = 0; x < rw; ++x)
        dst[c * plane + (size_t)y * tensor_w + x] = ((float)rs[((size_t)y * rw + x) * 3 + SRC[c]] / 255.0f - 0.5f) / 0.5f;
  return (int)rw;
}

// simd.rs:128-133 / 190-205 argmax: LAST maximal index wins
void oracle_ctc_argmax(const float* pred, int64_t rows, int vocab, int32_t* idx, float* prob) {
  for (int64_t r = 0; r < rows; ++r) {
    const float* row = pred + r * vocab;
    if (vocab == 0) {
      idx[r] = 0;
      prob[r] = 0.0f;
      continue;
    }
    int best = 0;
    float bv = row[0];
    for (int i = 1; i < vocab; ++i) {
      if (row[i] >= bv) {  // max_by keeps the last of equal maxima
        bv = row[i];
        best = i;
      }
    }
    idx[r] = best;
    prob[r] = bv;
  }
}

// decode.rs:505-614 decode_argmax(_with_positions).  n_chars = len(character list incl. blank).
// out_idx [B][T] kept class indices, out_cols [B][T] timesteps, out_len [B], out_score [B]
void oracle_ctc_decode(const int32_t* idx, const float* prob, int B, int T, int n_chars, int32_t* out_idx,
                       int32_t* out_cols, int32_t* out_len, float* out_score) {
  for (int b = 0; b < B; ++b) {
    int prev = 0, n = 0;
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      int k = idx[(size_t)b * T + t];
      if (k != 0 && k != prev && k >= 0 && k < n_chars) {
        out_idx[(size_t)b * T + n] = k;
        out_cols[(size_t)b * T + n] = t;
        sum += prob[(size_t)b * T + t];
        ++n;
      }
      prev = k;
    }
    out_len[b] = n;
    out_score[b] = n ? sum / (float)n : 0.0f;
  }
}

// ---- text-line orientation stage ----------------------------------------------------------
// image::imageops::rotate180 (image 0.25, EXT): out(w-1-x, h-1-y) = in(x, y); used by
// OAROCR::classify_line_orientations (src/oarocr/ocr.rs:755-792) on crops whose top class is 1.
void oracle_rotate180(const uint8_t* src, int w, int h, uint8_t* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const uint8_t* s = src + ((size_t)y * w + x) * 3;
      uint8_t* d = dst + ((size_t)(h - 1 - y) * w + (w - 1 - x)) * 3;
      d[0] = s[0], d[1] = s[1], d[2] = s[2];
    }
}

// Topk::extract_topk_from_prediction (oar-ocr-core/src/utils/topk.rs): (index, score) pairs, STABLE sort by score
// descending with partial_cmp (incomparable = Equal), first k.  Returns the number written.
int oracle_topk(const float* pred, int n, int k, int32_t* idx, float* score) {
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pred[a] > pred[b]; });
  int m = std::min(k, n);
  for (int i = 0; i < m; ++i) idx[i] = order[i], score[i] = pred[order[i]];
  return m;
}

}  // extern "C"

// ---- layout detection post-process (host half) ----------------------------------------------
// LayoutDetectionAdapter::postprocess_pp_doclayout and its helpers
// (oar-ocr-core/src/domain/adapters/layout_detection_adapter.rs:631-1116) + unclip_boxes
// (oar-ocr-core/src/processors/layout_postprocess.rs:636-681), restated step by step in f32.
// A box is BoundingBox::from_coords(x1, y1, x2, y2); x_min()/x_max() are the compare-and-keep loops of
// geometry.rs:179-209, 569-599 over its four corners (so a NaN coordinate falls back to the other corner's).
namespace {
struct LBox {
  float x1, y1, x2, y2;
  float xmin() const { float m = INFINITY; if (x1 < m) m = x1; if (x2 < m) m = x2; return m; }
  float xmax() const { float m = -INFINITY; if (x1 > m) m = x1; if (x2 > m) m = x2; return m; }
  float ymin() const { float m = INFINITY; if (y1 < m) m = y1; if (y2 < m) m = y2; return m; }
  float ymax() const { float m = -INFINITY; if (y1 > m) m = y1; if (y2 > m) m = y2; return m; }
};
inline float rmin(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }  // f32::min
inline float rmax(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }  // f32::max
inline float rclamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // f32::clamp (NaN stays)

// paddlex_iou, layout_detection_adapter.rs:935-951 (the "+ 1" pixel convention of PaddleX)
float layout_iou(const LBox& a, const LBox& b) {
  float x1 = a.xmin(), y1 = a.ymin(), x2 = a.xmax(), y2 = a.ymax();
  float x1p = b.xmin(), y1p = b.ymin(), x2p = b.xmax(), y2p = b.ymax();
  float iw = rmax(rmin(x2, x2p) - rmax(x1, x1p) + 1.0f, 0.0f);
  float ih = rmax(rmin(y2, y2p) - rmax(y1, y1p) + 1.0f, 0.0f);
  float inter = iw * ih;
  float area1 = (x2 - x1 + 1.0f) * (y2 - y1 + 1.0f);
  float area2 = (x2p - x1p + 1.0f) * (y2p - y1p + 1.0f);
  float uni = area1 + area2 - inter;
  return uni > 0.0f ? inter / uni : 0.0f;
}

// stable sort by score descending with partial_cmp (incomparable = Equal), :889-894
std::vector<int> score_order(const float* scores, int n) {
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  // descending, NaN last, ties by index: a total order (the reference's partial_cmp().unwrap_or(Equal) leaves the
  // order unspecified when a score is NaN; std::stable_sort needs a strict weak ordering)
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
    const float sa = scores[a], sb = scores[b];
    if (sa != sa) return false;
    if (sb != sb) return true;
    return sa > sb;
  });
  return idx;
}

// paddlex_layout_nms, :884-933: same class suppresses at IoU >= 0.6, other classes at >= 0.98, NaN IoU suppressed
std::vector<int> layout_nms(const std::vector<LBox>& boxes, const std::vector<int>& classes,
                            const std::vector<float>& scores) {
  std::vector<int> indices = score_order(scores.data(), (int)boxes.size());
  std::vector<char> suppressed(indices.size(), 0);
  std::vector<int> selected;
  for (size_t pos = 0; pos < indices.size(); ++pos) {
    if (suppressed[pos]) continue;
    int cur = indices[pos];
    selected.push_back(cur);
    for (size_t np = pos + 1; np < indices.size(); ++np) {
      if (suppressed[np]) continue;
      int idx = indices[np];
      float thr = classes[idx] == classes[cur] ? 0.6f : 0.98f;
      float iou = layout_iou(boxes[cur], boxes[idx]);
      if (iou >= thr || std::isnan(iou)) suppressed[np] = 1;
    }
  }
  return selected;
}

// is_contained, :1085-1106: intersection / area(inner) >= 0.9
bool layout_is_contained(const LBox& in, const LBox& out) {
  float x1 = in.xmin(), y1 = in.ymin(), x2 = in.xmax(), y2 = in.ymax();
  float x1p = out.xmin(), y1p = out.ymin(), x2p = out.xmax(), y2p = out.ymax();
  float area = (x2 - x1) * (y2 - y1);
  if (area <= 0.0f) return false;
  float iw = rmax(rmin(x2, x2p) - rmax(x1, x1p), 0.0f);
  float ih = rmax(rmin(y2, y2p) - rmax(y1, y1p), 0.0f);
  return iw * ih / area >= 0.9f;
}
}  // namespace

extern "C" {

int oracle_layout_nms(const float* boxes, const int32_t* classes, const float* scores, int n, int32_t* keep) {
  std::vector<LBox> b(n);
  for (int i = 0; i < n; ++i) b[i] = LBox{boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]};
  std::vector<int> sel = layout_nms(b, std::vector<int>(classes, classes + n), std::vector<float>(scores, scores + n));
  for (size_t i = 0; i < sel.size(); ++i) keep[i] = sel[i];
  return (int)sel.size();
}

// the reference's own test oracle, compacting_nms_reference (layout_detection_adapter.rs:1667-1697): rebuilds the
// candidate list after every selection and keeps idx while iou < threshold
int oracle_layout_nms_compacting(const float* boxes, const int32_t* classes, const float* scores, int n, int32_t* keep) {
  std::vector<LBox> b(n);
  for (int i = 0; i < n; ++i) b[i] = LBox{boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]};
  std::vector<int> indices = score_order(scores, n);
  int m = 0;
  while (!indices.empty()) {
    int cur = indices[0];
    keep[m++] = cur;
    std::vector<int> next;
    for (size_t k = 1; k < indices.size(); ++k) {
      int idx = indices[k];
      float thr = classes[idx] == classes[cur] ? 0.6f : 0.98f;
      if (layout_iou(b[cur], b[idx]) < thr) next.push_back(idx);
    }
    indices.swap(next);
  }
  return m;
}

// postprocess_pp_doclayout for ONE image (:674-840).  pred [num_boxes][feature_dim] rows
// [class_id, score, x1, y1, x2, y2, (order...)]; class_thresholds [num_classes] (NaN = not configured) or null;
// merge_modes [num_classes] (-1 not configured, 0 Large, 1 Small, 2 Union) or null; unclip_mode 0 none,
// 1 uniform/separate (unclip_w, unclip_h), 2 per class (class_unclip [num_classes][2], NaN = (1, 1)).
// Outputs up to max_elements rows; returns the count.
int oracle_layout_postprocess(const float* pred, int num_boxes, int feature_dim, float orig_w, float orig_h,
                              float score_threshold, int max_elements, int layout_nms_on, int num_classes,
                              const float* class_thresholds, const int32_t* merge_modes, int image_class_id,
                              int formula_class_id, int unclip_mode, float unclip_w, float unclip_h,
                              const float* class_unclip, float* out_boxes, int32_t* out_classes, float* out_scores) {
  const int order_mode = feature_dim == 8 ? 2 : (feature_dim == 7 ? 3 : 0);
  std::vector<LBox> boxes;
  std::vector<int> classes;
  std::vector<float> scores;
  std::vector<std::pair<float, float>> order;
  for (int i = 0; i < num_boxes; ++i) {
    const float* r = pred + (size_t)i * feature_dim;
    // `as i32` saturates and maps NaN to 0
    float cf = r[0];
    int class_id = std::isnan(cf) ? 0 : (cf >= 2147483648.0f ? INT32_MAX : (cf <= -2147483648.0f ? INT32_MIN : (int)cf));
    float score = r[1];
    if (class_id < 0 || class_id >= num_classes) continue;
    float thr = rmax(score_threshold, 0.0f);
    if (class_thresholds && !std::isnan(class_thresholds[class_id])) thr = class_thresholds[class_id];
    if (score < thr) continue;
    float x1 = r[2], y1 = r[3], x2 = r[4], y2 = r[5];
    // convert_bbox_coords, :848-878
    bool normalized = x2 <= 1.05f && y2 <= 1.05f && x1 >= -0.05f && y1 >= -0.05f && orig_w > 0.0f && orig_h > 0.0f;
    float sx1, sy1, sx2, sy2;
    if (normalized) {
      sx1 = rclamp(x1, 0.0f, 1.0f) * orig_w, sy1 = rclamp(y1, 0.0f, 1.0f) * orig_h;
      sx2 = rclamp(x2, 0.0f, 1.0f) * orig_w, sy2 = rclamp(y2, 0.0f, 1.0f) * orig_h;
    } else {
      sx1 = rclamp(x1, 0.0f, orig_w), sy1 = rclamp(y1, 0.0f, orig_h);
      sx2 = rclamp(x2, 0.0f, orig_w), sy2 = rclamp(y2, 0.0f, orig_h);
    }
    if (!(sx2 > sx1 && sy2 > sy1 && std::isfinite(sx1) && std::isfinite(sy1) && std::isfinite(sx2) && std::isfinite(sy2)))
      continue;  // is_valid_box, :880-882
    boxes.push_back(LBox{sx1, sy1, sx2, sy2});
    classes.push_back(class_id);
    scores.push_back(score);
    order.push_back(order_mode == 2 ? std::make_pair(r[6], r[7])
                                    : (order_mode == 3 ? std::make_pair(r[6], 0.0f) : std::make_pair(0.0f, 0.0f)));
  }
  auto select = [&](const std::vector<int>& keep) {
    std::vector<LBox> b;
    std::vector<int> c;
    std::vector<float> s;
    std::vector<std::pair<float, float>> o;
    for (int k : keep) b.push_back(boxes[k]), c.push_back(classes[k]), s.push_back(scores[k]), o.push_back(order[k]);
    boxes.swap(b), classes.swap(c), scores.swap(s), order.swap(o);
  };
  if (layout_nms_on && !boxes.empty()) select(layout_nms(boxes, classes, scores));
  // filter_large_image_boxes, :953-992
  if (image_class_id >= 0 && boxes.size() > 1) {
    float area_thres = orig_w > orig_h ? 0.82f : 0.93f;
    float img_area = orig_w * orig_h;
    std::vector<int> keep;
    for (size_t i = 0; i < boxes.size(); ++i) {
      if (classes[i] != image_class_id) {
        keep.push_back((int)i);
        continue;
      }
      float xmin = rmax(boxes[i].xmin(), 0.0f), ymin = rmax(boxes[i].ymin(), 0.0f);
      float xmax = rmin(boxes[i].xmax(), orig_w), ymax = rmin(boxes[i].ymax(), orig_h);
      float area = (xmax - xmin) * (ymax - ymin);
      if (area <= area_thres * img_area) keep.push_back((int)i);
    }
    if (!keep.empty()) select(keep);
  }
  // apply_paddlex_merge_modes, :994-1037 (+ check_containment :1039-1083)
  bool any_mode = false;
  if (merge_modes)
    for (int c = 0; c < num_classes; ++c) any_mode = any_mode || merge_modes[c] >= 0;
  if (any_mode && !boxes.empty()) {
    const int n = (int)boxes.size();
    std::vector<char> keep_mask(n, 1);
    for (int cls = 0; cls < num_classes; ++cls) {
      int mode = merge_modes[cls];
      if (mode != 0 && mode != 1) continue;  // Union (and unset) do nothing
      std::vector<int> contains(n, 0), contained(n, 0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          if (i == j) continue;
          if (formula_class_id >= 0 && classes[i] == formula_class_id && classes[j] != formula_class_id) continue;
          if (mode == 0 && classes[j] == cls && layout_is_contained(boxes[i], boxes[j])) contained[i] = 1, contains[j] = 1;
          if (mode == 1 && classes[i] == cls && layout_is_contained(boxes[i], boxes[j])) contained[i] = 1, contains[j] = 1;
        }
      for (int i = 0; i < n; ++i) {
        if (mode == 0 && contained[i] == 1) keep_mask[i] = 0;
        if (mode == 1 && !(contains[i] == 0 || contained[i] == 1)) keep_mask[i] = 0;
      }
    }
    std::vector<int> keep;
    for (int i = 0; i < n; ++i)
      if (keep_mask[i]) keep.push_back(i);
    select(keep);
  }
  // reading-order sort, :787-812: stable, total_cmp on (column, row) for V2, on the single key for V3
  if (order_mode != 0 && !boxes.empty()) {
    auto total_less = [](float a, float b) {  // f32::total_cmp
      int32_t x, y;
      memcpy(&x, &a, 4), memcpy(&y, &b, 4);
      x ^= (int32_t)(((uint32_t)(x >> 31)) >> 1);
      y ^= (int32_t)(((uint32_t)(y >> 31)) >> 1);
      return x < y;
    };
    std::vector<int> idx(boxes.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int i, int j) {
      if (total_less(order[i].first, order[j].first)) return true;
      if (total_less(order[j].first, order[i].first)) return false;
      return order_mode == 2 && total_less(order[i].second, order[j].second);
    });
    select(idx);
  }
  // unclip_boxes, layout_postprocess.rs:636-681
  if (unclip_mode != 0) {
    for (size_t i = 0; i < boxes.size(); ++i) {
      float wr = unclip_mode == 2 ? 1.0f : unclip_w, hr = unclip_mode == 2 ? 1.0f : unclip_h;
      if (unclip_mode == 2 && class_unclip && !std::isnan(class_unclip[2 * classes[i]]))
        wr = class_unclip[2 * classes[i]], hr = class_unclip[2 * classes[i] + 1];
      if (std::fabs(wr - 1.0f) < 1e-6f && std::fabs(hr - 1.0f) < 1e-6f) continue;
      float x_min = boxes[i].xmin(), y_min = boxes[i].ymin(), x_max = boxes[i].xmax(), y_max = boxes[i].ymax();
      float width = x_max - x_min, height = y_max - y_min;
      float cx = x_min + width * 0.5f, cy = y_min + height * 0.5f;
      float hw = width * wr * 0.5f, hh = height * hr * 0.5f;
      boxes[i] = LBox{cx - hw, cy - hh, cx + hw, cy + hh};
    }
  }
  int m = 0;
  for (size_t i = 0; i < boxes.size() && m < max_elements; ++i, ++m) {
    out_boxes[4 * m] = boxes[i].x1, out_boxes[4 * m + 1] = boxes[i].y1;
    out_boxes[4 * m + 2] = boxes[i].x2, out_boxes[4 * m + 3] = boxes[i].y2;
    out_classes[m] = classes[i];
    out_scores[m] = scores[i];
  }
  return m;
}

}  // extern "C"
