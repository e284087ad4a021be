// layout.cu -- host half of the layout-detection row: the PP-DocLayout post-process.
//
// Replaces LayoutDetectionAdapter::postprocess_pp_doclayout and its helpers
// (oar-ocr-core/src/domain/adapters/layout_detection_adapter.rs:631-1116: class/score filter, convert_bbox_coords,
// paddlex_layout_nms, filter_large_image_boxes, apply_paddlex_merge_modes / check_containment, the reading-order sort)
// and unclip_boxes (oar-ocr-core/src/processors/layout_postprocess.rs:636-681).
//
// The detector emits at most a few hundred rows per page (300 queries for RT-DETR-L), so this stage is host work in
// the reference and stays host work here -- like sort_quad_boxes it runs on the few KB the network hands back, never
// touches the device and needs no context.  The network itself (HGNetV2-L + hybrid encoder + deformable decoder) is
// the part of the row that is still to be built on the conv engine.
//
// Everything is f32 in the reference's operation order, so the kept boxes, their order and their coordinates are
// identical to the reference's (checked against the oracle restatement, bit for bit, in tests/test_layout_post.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace {

using namespace oar;

// f32::min / f32::max: a NaN operand yields the other one
inline float fmin_rs(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }
inline float fmax_rs(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }
inline float clamp_rs(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// f32::total_cmp as a strict-weak "less"
inline bool total_less(float a, float b) {
  int32_t x, y;
  memcpy(&x, &a, 4);
  memcpy(&y, &b, 4);
  x ^= (int32_t)(((uint32_t)(x >> 31)) >> 1);
  y ^= (int32_t)(((uint32_t)(y >> 31)) >> 1);
  return x < y;
}

// Candidates of one page, structure-of-arrays.  lo/hi are BoundingBox::x_min() ... y_max() of
// from_coords(x1, y1, x2, y2): compare-and-keep over the corners (geometry.rs:179-209, 569-599).
struct Candidates {
  std::vector<float> x1, y1, x2, y2, score, key0, key1;
  std::vector<int32_t> cls;
  size_t size() const { return cls.size(); }
  void push(float a, float b, float c, float d, int32_t k, float s, float o0, float o1) {
    x1.push_back(a), y1.push_back(b), x2.push_back(c), y2.push_back(d);
    cls.push_back(k), score.push_back(s), key0.push_back(o0), key1.push_back(o1);
  }
  static float lo(float a, float b) {
    float m = INFINITY;
    if (a < m) m = a;
    if (b < m) m = b;
    return m;
  }
  static float hi(float a, float b) {
    float m = -INFINITY;
    if (a > m) m = a;
    if (b > m) m = b;
    return m;
  }
  float xmin(size_t i) const { return lo(x1[i], x2[i]); }
  float xmax(size_t i) const { return hi(x1[i], x2[i]); }
  float ymin(size_t i) const { return lo(y1[i], y2[i]); }
  float ymax(size_t i) const { return hi(y1[i], y2[i]); }
  // keep rows `order[0..]` in that order (select_by_indices / select_by_mask)
  void gather(const std::vector<int>& order) {
    Candidates out;
    for (int k : order) out.push(x1[k], y1[k], x2[k], y2[k], cls[k], score[k], key0[k], key1[k]);
    *this = std::move(out);
  }
};

// paddlex_layout_nms (:884-933).  Axis-aligned extents and the PaddleX "+1" areas are computed once per box; the
// pairwise loop is the reference's: candidates in stable score order, a selected box suppresses every later one whose
// IoU reaches 0.6 (same class) / 0.98 (other class), NaN IoUs are suppressed.
std::vector<int> nms(const Candidates& c) {
  const int n = (int)c.size();
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  // The reference sorts with partial_cmp(..).unwrap_or(Equal) (layout_detection_adapter.rs:890-894): memory-safe in
  // Rust, but with a NaN score it is not a strict weak ordering, which is undefined behaviour for std::stable_sort.
  // A total order is used instead: descending score, NaN after every number, ties by index (stable).  For NaN-free
  // input -- the only case in which the reference's order is specified -- the result is identical.
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const float sa = c.score[a], sb = c.score[b];
    if (sa != sa) return false;
    if (sb != sb) return true;
    return sa > sb;
  });
  std::vector<float> bx1(n), by1(n), bx2(n), by2(n), area(n);
  for (int i = 0; i < n; ++i) {
    bx1[i] = c.xmin(i), by1[i] = c.ymin(i), bx2[i] = c.xmax(i), by2[i] = c.ymax(i);
    area[i] = (bx2[i] - bx1[i] + 1.0f) * (by2[i] - by1[i] + 1.0f);
  }
  std::vector<char> dead(n, 0);
  std::vector<int> keep;
  for (int p = 0; p < n; ++p) {
    if (dead[p]) continue;
    const int a = order[p];
    keep.push_back(a);
    for (int q = p + 1; q < n; ++q) {
      if (dead[q]) continue;
      const int b = order[q];
      const float iw = fmax_rs(fmin_rs(bx2[a], bx2[b]) - fmax_rs(bx1[a], bx1[b]) + 1.0f, 0.0f);
      const float ih = fmax_rs(fmin_rs(by2[a], by2[b]) - fmax_rs(by1[a], by1[b]) + 1.0f, 0.0f);
      const float inter = iw * ih;
      const float uni = area[a] + area[b] - inter;
      const float iou = uni > 0.0f ? inter / uni : 0.0f;
      const float thr = c.cls[b] == c.cls[a] ? 0.6f : 0.98f;
      if (iou >= thr || std::isnan(iou)) dead[q] = 1;
    }
  }
  return keep;
}

// is_contained (:1085-1106): at least 90 % of `in` lies inside `out`
bool contained_in(const Candidates& c, int in, int out) {
  const float x1 = c.xmin(in), y1 = c.ymin(in), x2 = c.xmax(in), y2 = c.ymax(in);
  const float area = (x2 - x1) * (y2 - y1);
  if (area <= 0.0f) return false;
  const float iw = fmax_rs(fmin_rs(x2, c.xmax(out)) - fmax_rs(x1, c.xmin(out)), 0.0f);
  const float ih = fmax_rs(fmin_rs(y2, c.ymax(out)) - fmax_rs(y1, c.ymin(out)), 0.0f);
  return iw * ih / area >= 0.9f;
}

int postprocess_page(const float* pred, int num_boxes, int fdim, float W, float H, const oar_layout_config& cfg,
                     float* out_boxes, int32_t* out_cls, float* out_scores) {
  const int order_mode = fdim == 8 ? 2 : (fdim == 7 ? 1 : 0);  // 2: (column, row) keys, 1: one key, 0: none
  Candidates c;
  for (int i = 0; i < num_boxes; ++i) {
    const float* r = pred + (size_t)i * fdim;
    const float cf = r[0];  // `as i32`: saturating, NaN -> 0
    const int32_t k = std::isnan(cf) ? 0 : (cf >= 2147483648.0f ? INT32_MAX : (cf <= -2147483648.0f ? INT32_MIN : (int32_t)cf));
    if (k < 0 || k >= cfg.num_classes) continue;
    float thr = fmax_rs(cfg.score_threshold, 0.0f);
    if (cfg.class_thresholds && !std::isnan(cfg.class_thresholds[k])) thr = cfg.class_thresholds[k];
    if (r[1] < thr) continue;
    // convert_bbox_coords (:848-878): normalised outputs are scaled to the source page, pixel outputs clamped to it
    const bool normalised = r[4] <= 1.05f && r[5] <= 1.05f && r[2] >= -0.05f && r[3] >= -0.05f && W > 0.0f && H > 0.0f;
    float x1, y1, x2, y2;
    if (normalised) {
      x1 = clamp_rs(r[2], 0.0f, 1.0f) * W, y1 = clamp_rs(r[3], 0.0f, 1.0f) * H;
      x2 = clamp_rs(r[4], 0.0f, 1.0f) * W, y2 = clamp_rs(r[5], 0.0f, 1.0f) * H;
    } else {
      x1 = clamp_rs(r[2], 0.0f, W), y1 = clamp_rs(r[3], 0.0f, H);
      x2 = clamp_rs(r[4], 0.0f, W), y2 = clamp_rs(r[5], 0.0f, H);
    }
    if (!(x2 > x1 && y2 > y1 && std::isfinite(x1) && std::isfinite(y1) && std::isfinite(x2) && std::isfinite(y2)))
      continue;
    c.push(x1, y1, x2, y2, k, r[1], order_mode ? r[6] : 0.0f, order_mode == 2 ? r[7] : 0.0f);
  }
  if (cfg.layout_nms && c.size()) c.gather(nms(c));
  // filter_large_image_boxes (:953-992): an "image" box covering (almost) the whole page is dropped
  if (cfg.image_class_id >= 0 && c.size() > 1) {
    const float limit = (W > H ? 0.82f : 0.93f) * (W * H);
    std::vector<int> keep;
    for (int i = 0; i < (int)c.size(); ++i) {
      bool ok = c.cls[i] != cfg.image_class_id;
      if (!ok) {
        const float xa = fmax_rs(c.xmin(i), 0.0f), ya = fmax_rs(c.ymin(i), 0.0f);
        const float xb = fmin_rs(c.xmax(i), W), yb = fmin_rs(c.ymax(i), H);
        ok = (xb - xa) * (yb - ya) <= limit;
      }
      if (ok) keep.push_back(i);
    }
    if (!keep.empty()) c.gather(keep);
  }
  // apply_paddlex_merge_modes (:994-1083).  Large: drop a box mostly inside a box of that class; Small: keep a box of
  // that class only if it contains nothing or is itself contained.  A formula is never tested against non-formulas.
  if (cfg.class_merge_modes && c.size()) {
    const int n = (int)c.size();
    bool any = false;
    for (int k = 0; k < cfg.num_classes; ++k) any = any || cfg.class_merge_modes[k] >= 0;
    if (any) {
      // containment is a property of the pair, not of the class under consideration: evaluate each pair once
      std::vector<char> inside((size_t)n * n, 0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          if (i == j) continue;
          if (cfg.formula_class_id >= 0 && c.cls[i] == cfg.formula_class_id && c.cls[j] != cfg.formula_class_id) continue;
          inside[(size_t)i * n + j] = contained_in(c, i, j);
        }
      std::vector<char> keep_mask(n, 1);
      for (int k = 0; k < cfg.num_classes; ++k) {
        const int mode = cfg.class_merge_modes[k];
        if (mode != OAR_MERGE_LARGE && mode != OAR_MERGE_SMALL) continue;
        std::vector<char> contains(n, 0), contained(n, 0);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            if (!inside[(size_t)i * n + j]) continue;
            if ((mode == OAR_MERGE_LARGE ? c.cls[j] : c.cls[i]) == k) contained[i] = 1, contains[j] = 1;
          }
        for (int i = 0; i < n; ++i) {
          if (mode == OAR_MERGE_LARGE && contained[i]) keep_mask[i] = 0;
          if (mode == OAR_MERGE_SMALL && !(contains[i] == 0 || contained[i] == 1)) keep_mask[i] = 0;
        }
      }
      std::vector<int> keep;
      for (int i = 0; i < n; ++i)
        if (keep_mask[i]) keep.push_back(i);
      c.gather(keep);
    }
  }
  // reading order the network predicted (:787-812): stable sort on total_cmp keys
  if (order_mode && c.size()) {
    std::vector<int> order(c.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int i, int j) {
      if (total_less(c.key0[i], c.key0[j])) return true;
      if (total_less(c.key0[j], c.key0[i])) return false;
      return order_mode == 2 && total_less(c.key1[i], c.key1[j]);
    });
    c.gather(order);
  }
  // unclip_boxes (layout_postprocess.rs:636-681): scale about the centre
  if (cfg.unclip_mode != OAR_UNCLIP_NONE) {  // OAR_UNCLIP_RATIO or OAR_UNCLIP_PER_CLASS
    for (size_t i = 0; i < c.size(); ++i) {
      float wr = cfg.unclip_w, hr = cfg.unclip_h;
      if (cfg.unclip_mode == OAR_UNCLIP_PER_CLASS) {
        wr = hr = 1.0f;
        if (cfg.class_unclip && !std::isnan(cfg.class_unclip[2 * c.cls[i]]))
          wr = cfg.class_unclip[2 * c.cls[i]], hr = cfg.class_unclip[2 * c.cls[i] + 1];
      }
      if (std::fabs(wr - 1.0f) < 1e-6f && std::fabs(hr - 1.0f) < 1e-6f) continue;
      const float xa = c.xmin(i), ya = c.ymin(i), w = c.xmax(i) - xa, h = c.ymax(i) - ya;
      const float cx = xa + w * 0.5f, cy = ya + h * 0.5f;
      const float hw = w * wr * 0.5f, hh = h * hr * 0.5f;
      c.x1[i] = cx - hw, c.y1[i] = cy - hh, c.x2[i] = cx + hw, c.y2[i] = cy + hh;
    }
  }
  const int m = (int)std::min<size_t>(c.size(), (size_t)std::max(cfg.max_elements, 1));
  for (int i = 0; i < m; ++i) {
    out_boxes[4 * i] = c.x1[i], out_boxes[4 * i + 1] = c.y1[i], out_boxes[4 * i + 2] = c.x2[i], out_boxes[4 * i + 3] = c.y2[i];
    out_cls[i] = c.cls[i];
    out_scores[i] = c.score[i];
  }
  return m;
}

}  // namespace

extern "C" {

void oar_layout_config_default(oar_layout_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->score_threshold = 0.5f;  // LayoutDetectionConfig::default, tasks/layout_detection.rs:88-99
  cfg->max_elements = 100;
  cfg->layout_nms = 1;
  cfg->num_classes = 23;        // LayoutModelConfig::pp_doclayout_l, layout_detection_adapter.rs:314-348
  cfg->image_class_id = 1;      // "image"
  cfg->formula_class_id = 7;    // "formula"
  cfg->unclip_mode = OAR_UNCLIP_NONE;
  cfg->unclip_w = cfg->unclip_h = 1.0f;
}

int32_t oar_layout_postprocess(const float* pred, int32_t batch, int32_t num_boxes, int32_t feature_dim,
                               const float* src_w, const float* src_h, const oar_layout_config* cfg, float* boxes,
                               int32_t* classes, float* scores, int32_t* counts) {
  try {
    if (!cfg || !counts) OAR_FAIL(OAR_E_INVALID, "null argument");
    if (batch <= 0) return OAR_OK;
    if (num_boxes < 0 || feature_dim < 6) OAR_FAIL(OAR_E_INVALID, "predictions must be [batch][boxes][>= 6]");
    if (cfg->num_classes <= 0 || cfg->max_elements < 1)  // #[validate(min = 1)] on max_elements
      OAR_FAIL(OAR_E_INVALID, "num_classes and max_elements must be positive");
    if ((num_boxes > 0 && !pred) || !src_w || !src_h || !boxes || !classes || !scores)
      OAR_FAIL(OAR_E_INVALID, "null argument");
    if (cfg->unclip_mode < OAR_UNCLIP_NONE || cfg->unclip_mode > OAR_UNCLIP_PER_CLASS)
      OAR_FAIL(OAR_E_INVALID, "unknown unclip mode %d", cfg->unclip_mode);
    const size_t me = (size_t)cfg->max_elements;
    for (int b = 0; b < batch; ++b)
      counts[b] = postprocess_page(pred + (size_t)b * num_boxes * feature_dim, num_boxes, feature_dim, src_w[b], src_h[b],
                                   *cfg, boxes + (size_t)b * me * 4, classes + (size_t)b * me, scores + (size_t)b * me);
  } catch (const oar::OarError& e) {
    return e.code;
  } catch (const std::exception& e) {
    oar::set_error("internal error: %s", e.what());
    return OAR_E_CUDA;
  }
  return OAR_OK;
}

}  // extern "C"
