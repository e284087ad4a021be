// oar_ocr.hpp -- header-only C++ mirror of the reference's Rust API for the det+rec path, over the C ABI
// (include/oar_b200.h).  Same names, argument meaning and error behaviour as the crate:
//   OAROCRBuilder / OAROCR::predict       src/oarocr/ocr.rs:66-417, 518-659
//   TextDetectionPredictor                oar-ocr-core/src/predictors/text_detection.rs:23-112
//   TextRecognitionPredictor              oar-ocr-core/src/predictors/text_recognition.rs:19-110
//   TextLineOrientationPredictor          oar-ocr-core/src/predictors/text_line_orientation.rs:18-105
//   LayoutDetectionConfig / postprocess   oar-ocr-core/src/domain/tasks/layout_detection.rs:45-100,
//                                         domain/adapters/layout_detection_adapter.rs:631-846 (host half of the row)
//   OCRError                              oar-ocr-core/src/core/errors/types.rs:110-214
// Errors are thrown as oar::OCRError (Rust returns Result<_, OCRError>).  No CPU fallback exists.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "oar_b200.h"

namespace oar {

struct OCRError : std::runtime_error {
  std::string kind;  // variant name: InvalidInput, ConfigError, Inference, ModelLoad
  int code;
  OCRError(std::string k, const std::string& msg, int c = 0)
      : std::runtime_error(k + ": " + msg), kind(std::move(k)), code(c) {}
};

inline void check(int32_t rc) {
  if (rc == OAR_OK) return;
  const char* kind = rc == OAR_E_INVALID ? "InvalidInput" : rc == OAR_E_MODEL ? "ModelLoad"
                     : rc == OAR_E_UNSUPPORTED ? "ConfigError" : "Inference";
  throw OCRError(kind, oar_last_error(), rc);
}

// The recogniser's character table from the dictionary file's UTF-8 content, as the reference derives it:
// `content.lines()` (ocr.rs:386 -- '\n' ends a line, a '\r' directly before it belongs to the terminator, no final
// empty line) then CTCLabelDecode::from_string_list(.., use_space_char = true, has_explicit_blank = false)
// (decode.rs:118-141, 392-423): blank '\0', the FIRST code point of every non-empty line, ' '.  `n_chars` of the C ABI
// is the size of this table; text = table[label index].
inline std::vector<char32_t> character_list(const std::string& dict_utf8) {
  std::vector<char32_t> chars{U'\0'};
  auto first_code_point = [](const std::string& s, size_t b, size_t e, char32_t* cp) {
    if (b >= e) return false;
    const unsigned char c0 = (unsigned char)s[b];
    int n = c0 < 0x80 ? 1 : (c0 >> 5) == 6 ? 2 : (c0 >> 4) == 14 ? 3 : (c0 >> 3) == 30 ? 4 : 1;
    char32_t v = n == 1 ? c0 : (c0 & (0xFF >> (n + 1)));
    for (int i = 1; i < n && b + i < e; ++i) v = (v << 6) | ((unsigned char)s[b + i] & 0x3F);
    *cp = v;
    return true;
  };
  size_t pos = 0;
  while (pos < dict_utf8.size()) {
    size_t nl = dict_utf8.find('\n', pos);
    size_t end = nl == std::string::npos ? dict_utf8.size() : nl;
    size_t line_end = (nl != std::string::npos && end > pos && dict_utf8[end - 1] == '\r') ? end - 1 : end;
    char32_t cp;
    if (first_code_point(dict_utf8, pos, line_end, &cp)) chars.push_back(cp);
    pos = nl == std::string::npos ? dict_utf8.size() : nl + 1;
  }
  chars.push_back(U' ');
  return chars;
}

struct RgbImage {  // image::RgbImage: u8 HWC, row-major
  const uint8_t* data = nullptr;
  int32_t height = 0, width = 0;
};
struct Point { float x, y; };
struct BoundingBox { std::vector<Point> points; };
struct Detection { BoundingBox bbox; float score; };
struct TextDetectionResult { std::vector<std::vector<Detection>> detections; };
struct TextRecognitionResult {
  std::vector<std::vector<int32_t>> label_indices;  // text = character[index] (decode.rs:392-423)
  std::vector<float> scores;
  std::vector<std::vector<int32_t>> char_col_indices;
  std::vector<size_t> sequence_lengths;
};
struct TextRegion {
  BoundingBox bounding_box;
  std::vector<int32_t> label_indices;
  float confidence = 0.0f;
  int32_t detection_index = 0;
  // return_word_box (ocr.rs:241): CTC column of every emitted character, the sequence length of the region's
  // recognition batch, its crop's w/h ratio and the batch's chunk_max_wh_ratio -- the inputs of ctc_word_boxes()
  std::vector<int32_t> char_col_indices;
  int32_t sequence_length = 0;
  float wh_ratio = 0.0f, max_wh_ratio = 0.0f;
  // line-orientation stage (ocr.rs:755-792): Some(0.0) / Some(180.0) when a classifier is attached, None otherwise
  bool has_orientation_angle = false;
  float orientation_angle = 0.0f;
};
struct Classification {  // domain/tasks: Classification{class_id, label, score}
  size_t class_id;
  std::string label;
  float score;
};
struct TextLineOrientationResult { std::vector<std::vector<Classification>> orientations; };

// Topk::extract_topk_from_prediction (oar-ocr-core/src/utils/topk.rs): stable sort by score descending, first k
inline std::vector<size_t> topk_indices(const float* pred, size_t n, size_t k) {
  if (k == 0) throw OCRError("InvalidInput", "k must be greater than 0", OAR_E_INVALID);
  std::vector<size_t> order(n);
  std::iota(order.begin(), order.end(), (size_t)0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return pred[a] > pred[b]; });
  order.resize(std::min(k, n));
  return order;
}

// OAROCR::is_cjk (src/oarocr/ocr.rs:1065-1084)
inline bool is_cjk(char32_t u) {
  return (u >= 0x4E00 && u <= 0x9FFF) || (u >= 0x3400 && u <= 0x4DBF) || (u >= 0x20000 && u <= 0x2A6DF) ||
         (u >= 0x2A700 && u <= 0x2B73F) || (u >= 0x2B740 && u <= 0x2B81F);
}

// OAROCR::ctc_word_boxes (src/oarocr/ocr.rs:949-1022): one box per character from its CTC timestep.  `text` holds the
// region's characters as code points (the dictionary lives above the ABI).  f32 arithmetic in the reference's order.
inline std::vector<BoundingBox> ctc_word_boxes(const BoundingBox& line, const std::u32string& text,
                                               const std::vector<int32_t>& cols, size_t seq_len, float wh_ratio,
                                               float max_wh_ratio) {
  std::vector<BoundingBox> out;
  if (cols.empty() || seq_len == 0 || text.empty() || line.points.empty()) return out;
  const float eps = 1.1920929e-07f;  // f32::EPSILON
  const float effective = (float)seq_len * (wh_ratio / max_wh_ratio);
  if (effective <= eps) return out;
  float x_min = line.points[0].x, x_max = x_min, y_min = line.points[0].y, y_max = y_min;
  for (const Point& p : line.points) {
    x_min = p.x < x_min ? p.x : x_min, x_max = p.x > x_max ? p.x : x_max;
    y_min = p.y < y_min ? p.y : y_min, y_max = p.y > y_max ? p.y : y_max;
  }
  const float width = x_max - x_min;
  const float cell = width / (effective > eps ? effective : eps);
  const float avg = width / (float)(text.size() > 0 ? text.size() : 1);
  std::vector<float> centers(cols.size());
  for (size_t i = 0; i < cols.size(); ++i) centers[i] = x_min + ((float)cols[i] + 0.5f) * cell;
  auto box = [&](float a, float b) {
    return BoundingBox{{Point{a, y_min}, Point{b, y_min}, Point{b, y_max}, Point{a, y_max}}};  // from_coords
  };
  for (size_t i = 0; i < cols.size(); ++i) {
    const char32_t ch = i < text.size() ? text[i] : U'?';
    const float c = centers[i];
    float a, b;
    if (is_cjk(ch)) {
      const float half = avg / 2.0f;
      a = c - half, b = c + half;
    } else {
      a = i == 0 ? x_min : (centers[i - 1] + c) / 2.0f;
      b = i + 1 == cols.size() ? x_max : (c + centers[i + 1]) / 2.0f;
    }
    a = a > x_min ? a : x_min;
    b = b < x_max ? b : x_max;
    out.push_back(box(a, b));
  }
  return out;
}
struct OAROCRResult { size_t index = 0; std::vector<TextRegion> text_regions; };

class Context {
 public:
  explicit Context(int32_t device_id = 0) { check(oar_ctx_create(device_id, &ctx_)); }
  ~Context() { oar_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  oar_ctx* raw() const { return ctx_; }
 private:
  oar_ctx* ctx_ = nullptr;
};

class Model {  // stands where OrtInfer stands in the reference
 public:
  Model(Context& ctx, const void* blob, size_t len) { check(oar_model_load_blob(ctx.raw(), blob, len, &m_)); }
  ~Model() { oar_model_destroy(m_); }
  Model(const Model&) = delete;
  Model& operator=(const Model&) = delete;
  oar_model* raw() const { return m_; }
 private:
  oar_model* m_ = nullptr;
};

namespace detail {
inline void image_table(const std::vector<RgbImage>& images, const char* what, std::vector<const uint8_t*>& ptrs,
                        std::vector<int32_t>& hs, std::vector<int32_t>& ws) {
  if (images.empty())  // OCRError::validation_error(.., "non-empty slice", "empty slice"), ocr.rs:525-532
    throw OCRError("InvalidInput", std::string(what) + ": images: expected non-empty slice, got empty slice",
                   OAR_E_INVALID);
  for (const auto& im : images) {
    ptrs.push_back(im.data);
    hs.push_back(im.height);
    ws.push_back(im.width);
  }
}
}  // namespace detail

struct TextDetectionConfig {  // tasks/text_detection.rs:33-53; predictor defaults text_detection.rs:56-68
  float score_threshold = 0.3f, box_threshold = 0.6f, unclip_ratio = 1.5f;
  int32_t max_candidates = 1000, limit_side_len = 960, limit_type = 0, max_side_len = 4000;
  oar_det_config to_abi() const {
    oar_det_config c;
    oar_det_config_default(&c);
    c.thresh = score_threshold, c.box_thresh = box_threshold, c.unclip_ratio = unclip_ratio;
    c.max_candidates = max_candidates, c.limit_side_len = limit_side_len, c.limit_type = limit_type;
    c.max_side_limit = max_side_len;
    return c;
  }
};

class TextDetectionPredictor {
 public:
  TextDetectionPredictor(Model& model, TextDetectionConfig cfg = {}) : model_(model), cfg_(cfg) {}
  // unsorted, discovery order, as TextDetectionAdapter::execute returns them
  TextDetectionResult predict(const std::vector<RgbImage>& images) const {
    std::vector<const uint8_t*> ptrs;
    std::vector<int32_t> hs, ws;
    detail::image_table(images, "TextDetection", ptrs, hs, ws);
    const size_t n = images.size(), mc = (size_t)cfg_.max_candidates;
    oar_det_config c = cfg_.to_abi();
    std::vector<float> boxes(n * mc * 8), scores(n * mc);
    std::vector<int32_t> counts(n);
    check(oar_det_run(model_.raw(), ptrs.data(), hs.data(), ws.data(), (int32_t)n, &c, boxes.data(), scores.data(),
                      counts.data()));
    TextDetectionResult res;
    res.detections.resize(n);
    for (size_t i = 0; i < n; ++i)
      for (int32_t k = 0; k < counts[i]; ++k) {
        const float* p = &boxes[(i * mc + k) * 8];
        Detection d;
        for (int j = 0; j < 4; ++j) d.bbox.points.push_back(Point{p[2 * j], p[2 * j + 1]});
        d.score = scores[i * mc + k];
        if (!(d.score >= 0.0f && d.score <= 1.0f))  // validate_output, tasks/text_detection.rs:129-141
          throw OCRError("InvalidInput", "detection score outside [0,1]");
        res.detections[i].push_back(std::move(d));
      }
    return res;
  }
 private:
  Model& model_;
  TextDetectionConfig cfg_;
};

class TextRecognitionPredictor {
 public:
  TextRecognitionPredictor(Model& model, int32_t n_chars, float score_threshold = 0.0f)
      : model_(model), n_chars_(n_chars), thresh_(score_threshold) {}
  // the whole input is ONE batch (predictors/text_recognition.rs:38-45)
  TextRecognitionResult predict(const std::vector<RgbImage>& images) const {
    std::vector<const uint8_t*> ptrs;
    std::vector<int32_t> hs, ws;
    detail::image_table(images, "TextRecognition", ptrs, hs, ws);
    const size_t n = images.size();
    const int32_t t_cap = 3200 / 8 + 2;
    std::vector<int32_t> labels(n * t_cap), cols(n * t_cap), lens(n);
    std::vector<float> scores(n);
    int32_t t_out = 0;
    check(oar_rec_run(model_.raw(), ptrs.data(), hs.data(), ws.data(), (int32_t)n, n_chars_, labels.data(), cols.data(),
                      lens.data(), scores.data(), t_cap, &t_out));
    TextRecognitionResult r;
    for (size_t i = 0; i < n; ++i) {
      const bool keep = scores[i] >= thresh_;  // text_recognition_adapter.rs:88-102
      r.label_indices.emplace_back(labels.begin() + i * t_cap, labels.begin() + i * t_cap + (keep ? lens[i] : 0));
      r.char_col_indices.emplace_back(cols.begin() + i * t_cap, cols.begin() + i * t_cap + lens[i]);
      r.scores.push_back(scores[i]);
      r.sequence_lengths.push_back((size_t)t_out);
    }
    return r;
  }
 private:
  Model& model_;
  int32_t n_chars_;
  float thresh_;
};

// ---- layout detection, host half: LayoutDetectionAdapter::postprocess_pp_doclayout over oar_layout_postprocess ----
enum class MergeBboxMode { Large = OAR_MERGE_LARGE, Small = OAR_MERGE_SMALL, Union = OAR_MERGE_UNION };
struct LayoutDetectionElement { BoundingBox bbox; std::string element_type; float score; };
struct LayoutDetectionConfig {  // tasks/layout_detection.rs:45-100
  float score_threshold = 0.5f;
  size_t max_elements = 100;
  std::map<std::string, float> class_thresholds;           // empty = None
  std::map<std::string, MergeBboxMode> class_merge_modes;  // empty = None
  bool layout_nms = true;
  int unclip_mode = OAR_UNCLIP_NONE;                       // OAR_UNCLIP_RATIO: (unclip_w, unclip_h)
  float unclip_w = 1.0f, unclip_h = 1.0f;
  std::map<size_t, std::pair<float, float>> class_unclip;  // OAR_UNCLIP_PER_CLASS
};
// predictions: [batch][num_boxes][feature_dim] rows [class_id, score, x1, y1, x2, y2, (order keys)]; src_wh = (w, h) of
// every source page; class_labels[id] = label (LayoutModelConfig.class_labels).  Host only: no Context needed.
inline std::vector<std::vector<LayoutDetectionElement>> postprocess_pp_doclayout(
    const std::vector<float>& predictions, size_t batch, size_t num_boxes, size_t feature_dim,
    const std::vector<std::pair<float, float>>& src_wh, const LayoutDetectionConfig& config,
    const std::vector<std::string>& class_labels) {
  if (config.max_elements < 1) throw OCRError("ConfigError", "max_elements must be at least 1");
  if (src_wh.size() != batch || predictions.size() != batch * num_boxes * feature_dim)
    throw OCRError("InvalidInput", "predictions / image shapes do not match the batch", OAR_E_INVALID);
  const size_t nc = class_labels.size();
  auto id_of = [&](const std::string& label) {
    for (size_t i = 0; i < nc; ++i)
      if (class_labels[i] == label) return (int32_t)i;
    return (int32_t)-1;
  };
  std::vector<float> thr(nc, NAN), cu(2 * nc, NAN);
  std::vector<int32_t> modes(nc, OAR_MERGE_UNSET);
  for (const auto& kv : config.class_thresholds)
    if (id_of(kv.first) >= 0) thr[id_of(kv.first)] = kv.second;
  for (const auto& kv : config.class_merge_modes)
    if (id_of(kv.first) >= 0) modes[id_of(kv.first)] = (int32_t)kv.second;
  for (const auto& kv : config.class_unclip)
    if (kv.first < nc) cu[2 * kv.first] = kv.second.first, cu[2 * kv.first + 1] = kv.second.second;
  oar_layout_config c;
  oar_layout_config_default(&c);
  c.score_threshold = config.score_threshold, c.max_elements = (int32_t)config.max_elements;
  c.layout_nms = config.layout_nms ? 1 : 0, c.num_classes = (int32_t)nc;
  c.class_thresholds = config.class_thresholds.empty() ? nullptr : thr.data();
  c.class_merge_modes = config.class_merge_modes.empty() ? nullptr : modes.data();
  c.image_class_id = id_of("image"), c.formula_class_id = id_of("formula");
  c.unclip_mode = config.unclip_mode, c.unclip_w = config.unclip_w, c.unclip_h = config.unclip_h;
  c.class_unclip = cu.data();
  std::vector<float> sw(batch), sh(batch), boxes(batch * config.max_elements * 4), scores(batch * config.max_elements);
  std::vector<int32_t> classes(batch * config.max_elements), counts(batch);
  for (size_t b = 0; b < batch; ++b) sw[b] = src_wh[b].first, sh[b] = src_wh[b].second;
  check(oar_layout_postprocess(predictions.data(), (int32_t)batch, (int32_t)num_boxes, (int32_t)feature_dim, sw.data(),
                               sh.data(), &c, boxes.data(), classes.data(), scores.data(), counts.data()));
  std::vector<std::vector<LayoutDetectionElement>> out(batch);
  for (size_t b = 0; b < batch; ++b)
    for (int32_t k = 0; k < counts[b]; ++k) {
      const size_t r = b * config.max_elements + (size_t)k;
      const float* q = &boxes[r * 4];
      BoundingBox bb{{Point{q[0], q[1]}, Point{q[2], q[1]}, Point{q[2], q[3]}, Point{q[0], q[3]}}};  // from_coords
      const size_t cid = (size_t)classes[r];
      out[b].push_back(LayoutDetectionElement{std::move(bb), cid < nc ? class_labels[cid] : "unknown", scores[r]});
    }
  return out;
}

class TextLineOrientationPredictor {
 public:
  // builder defaults: score_threshold 0.5, topk 2, input_shape (192, 48) read as (height, width)
  // (predictors/text_line_orientation.rs:53-61); the pipeline's adapter uses (80, 160)
  TextLineOrientationPredictor(Model& model, size_t topk = 2, int32_t input_h = 192, int32_t input_w = 48)
      : model_(model), topk_(topk), h_(input_h), w_(input_w) {
    if (topk == 0) throw OCRError("ConfigError", "topk must be at least 1");
  }
  TextLineOrientationResult predict(const std::vector<RgbImage>& images) const {
    if (images.empty())  // tasks/text_line_orientation.rs:88-95
      throw OCRError("InvalidInput", "No images provided for text line orientation classification", OAR_E_INVALID);
    std::vector<const uint8_t*> ptrs;
    std::vector<int32_t> hs, ws;
    detail::image_table(images, "TextLineOrientation", ptrs, hs, ws);
    const size_t n = images.size(), cap = n * 1024;
    std::vector<int32_t> ids(n);
    std::vector<float> scores(n), probs(cap);
    int32_t nc = 0;
    check(oar_cls_run(model_.raw(), ptrs.data(), hs.data(), ws.data(), (int32_t)n, h_, w_, ids.data(), scores.data(),
                      probs.data(), cap, &nc));
    TextLineOrientationResult r;
    r.orientations.resize(n);
    for (size_t i = 0; i < n; ++i)
      for (size_t c : topk_indices(&probs[i * nc], (size_t)nc, topk_))  // labels "0" / "180" (adapter labels())
        r.orientations[i].push_back(Classification{c, c < 2 ? std::to_string(c * 180) : "class_" + std::to_string(c),
                                                   probs[i * nc + c]});
    return r;
  }
 private:
  Model& model_;
  size_t topk_;
  int32_t h_, w_;
};

class OAROCR {
 public:
  OAROCR(Model& det, Model& rec, oar_pipeline_config cfg, Model* cls = nullptr)
      : det_(det), rec_(rec), cls_(cls), cfg_(cfg) {}
  std::vector<OAROCRResult> predict(const std::vector<RgbImage>& images) const {
    std::vector<const uint8_t*> ptrs;
    std::vector<int32_t> hs, ws;
    detail::image_table(images, "OCR Pipeline", ptrs, hs, ws);
    const size_t n = images.size();
    const int32_t cap_r = (int32_t)n * cfg_.det.max_candidates, cap_l = cap_r * 64;
    std::vector<int32_t> region_off(n + 1), det_index(cap_r), label_off(cap_r + 1), labels(cap_l);
    std::vector<float> boxes((size_t)cap_r * 8), scores(cap_r), wh_ratio(cap_r), max_wh_ratio(cap_r);
    std::vector<int32_t> cols(cap_l), seq_len(cap_r);
    std::vector<float> line_angle(cap_r, -1.0f);
    oar_ocr_result out{};
    out.line_angle = line_angle.data();
    out.cols = cols.data(), out.seq_len = seq_len.data(), out.wh_ratio = wh_ratio.data();
    out.max_wh_ratio = max_wh_ratio.data();
    out.cap_regions = cap_r, out.cap_labels = cap_l;
    out.region_off = region_off.data(), out.boxes = boxes.data(), out.scores = scores.data();
    out.det_index = det_index.data(), out.label_off = label_off.data(), out.labels = labels.data();
    check(oar_pipeline_run_cls(det_.raw(), rec_.raw(), cls_ ? cls_->raw() : nullptr, ptrs.data(), hs.data(), ws.data(),
                               (int32_t)n, 0, &cfg_, &out));
    std::vector<OAROCRResult> res(n);
    for (size_t i = 0; i < n; ++i) {
      res[i].index = i;
      for (int32_t r = region_off[i]; r < region_off[i + 1]; ++r) {
        TextRegion t;
        for (int j = 0; j < 4; ++j) t.bounding_box.points.push_back(Point{boxes[r * 8 + 2 * j], boxes[r * 8 + 2 * j + 1]});
        t.label_indices.assign(labels.begin() + label_off[r], labels.begin() + label_off[r + 1]);
        t.confidence = scores[r];
        t.detection_index = det_index[r];
        t.char_col_indices.assign(cols.begin() + label_off[r], cols.begin() + label_off[r + 1]);
        t.sequence_length = seq_len[r], t.wh_ratio = wh_ratio[r], t.max_wh_ratio = max_wh_ratio[r];
        t.has_orientation_angle = line_angle[r] >= 0.0f, t.orientation_angle = t.has_orientation_angle ? line_angle[r] : 0.0f;
        res[i].text_regions.push_back(std::move(t));
      }
    }
    return res;
  }
 private:
  Model& det_;
  Model& rec_;
  Model* cls_;
  oar_pipeline_config cfg_;
};

class OAROCRBuilder {
 public:
  static constexpr size_t MAX_BATCH_SIZE = 4096;  // ocr.rs:93
  OAROCRBuilder(Model& det, Model& rec, int32_t n_chars) : det_(det), rec_(rec) {
    oar_pipeline_config_default(&cfg_);  // thresh .3 / box .6 / unclip 2.0 / 960 Max 4000 (ocr.rs:351-364), 8 / 64
    cfg_.n_chars = n_chars;
  }
  static void validate_batch_size(const char* name, size_t size) {  // ocr.rs:419-430
    if (size == 0 || size > MAX_BATCH_SIZE)
      throw OCRError("ConfigError", std::string(name) + " must be in 1..=" + std::to_string(MAX_BATCH_SIZE) + ", got " +
                                        std::to_string(size));
  }
  OAROCRBuilder& image_batch_size(size_t s) { image_bs_ = s, has_image_bs_ = true; return *this; }
  OAROCRBuilder& region_batch_size(size_t s) { region_bs_ = s, has_region_bs_ = true; return *this; }
  OAROCRBuilder& text_detection_config(const TextDetectionConfig& c) { cfg_.det = c.to_abi(); return *this; }
  OAROCRBuilder& rec_score_threshold(float t) { cfg_.rec_score_thresh = t; return *this; }
  // OAROCRBuilder::with_text_line_orientation_classification (ocr.rs:197-203)
  OAROCRBuilder& with_text_line_orientation_classification(Model& cls) { cls_ = &cls; return *this; }
  OAROCR build() {
    if (has_image_bs_) validate_batch_size("image_batch_size", image_bs_), cfg_.image_batch_size = (int32_t)image_bs_;
    if (has_region_bs_) validate_batch_size("region_batch_size", region_bs_), cfg_.region_batch_size = (int32_t)region_bs_;
    return OAROCR(det_, rec_, cfg_, cls_);
  }
 private:
  Model& det_;
  Model& rec_;
  Model* cls_ = nullptr;
  oar_pipeline_config cfg_;
  size_t image_bs_ = 0, region_bs_ = 0;
  bool has_image_bs_ = false, has_region_bs_ = false;
};

}  // namespace oar
