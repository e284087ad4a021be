// Exercises include/oar_ocr.hpp (the C++ mirror of the Rust API) against liboar_b200.so.
// Without a GPU: checks the error behaviour the reference pins (ocr.rs:1168-1195, no CPU fallback).
// With a GPU (argv[1] = det blob, argv[2] = rec blob, optional argv[3] = line-orientation classifier blob): runs
// predict() on a blank page and one synthetic stripe, with and without the classifier.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

#include "oar_ocr.hpp"

static std::vector<char> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  int failures = 0;
  auto expect = [&](bool ok, const char* what) {
    if (!ok) ++failures, std::printf("FAIL %s\n", what);
  };
  // batch-size validation (ocr.rs:1168-1195)
  try {
    oar::OAROCRBuilder::validate_batch_size("image_batch_size", 1);
    oar::OAROCRBuilder::validate_batch_size("region_batch_size", oar::OAROCRBuilder::MAX_BATCH_SIZE);
  } catch (...) {
    expect(false, "bounds accepted");
  }
  try {
    oar::OAROCRBuilder::validate_batch_size("image_batch_size", 0);
    expect(false, "zero rejected");
  } catch (const oar::OCRError& e) {
    expect(std::strstr(e.what(), "image_batch_size") && std::strstr(e.what(), "1..=4096"), "zero message");
  }
  oar_pipeline_config pc;
  oar_pipeline_config_default(&pc);
  expect(pc.image_batch_size == 8 && pc.region_batch_size == 64 && pc.det.unclip_ratio == 2.0f, "defaults");

  // character table from dictionary content: Rust lines() + first char of each non-empty line (ocr.rs:386,
  // decode.rs:118-121); U+2028 / U+0085 / \x0b are ordinary characters, "\r\n" is one terminator, a bare CR is not
  {
    const std::string dict = "a\n\nbc\r\n\xE2\x80\xA8\n\x0bq\n\xC2\x85\n\xE4\xB8\x80z\nz\r";
    auto t = oar::character_list(dict);
    const std::vector<char32_t> want{U'\0', U'a', U'b', (char32_t)0x2028, (char32_t)0x0B, (char32_t)0x85, (char32_t)0x4E00, U'z', U' '};
    expect(t == want, "character_list");
    expect(oar::character_list("").size() == 2 && oar::character_list("x\n\n").size() == 3, "character_list edge cases");
  }

  // ctc_word_boxes: the reference's own vector (ocr.rs:1197-1232) + a CJK case
  {
    oar::BoundingBox line{{{0.f, 0.f}, {100.f, 0.f}, {100.f, 20.f}, {0.f, 20.f}}};
    auto wb = oar::ctc_word_boxes(line, U"ABC", {1, 4, 7}, 10, 5.0f, 5.0f);
    auto near = [](float a, float b) { return a - b < 1e-5f && b - a < 1e-5f; };
    expect(wb.size() == 3 && near(wb[0].points[0].x, 0.f) && near(wb[0].points[1].x, 30.f) &&
               near(wb[1].points[0].x, 30.f) && near(wb[1].points[1].x, 60.f) && near(wb[2].points[0].x, 60.f) &&
               near(wb[2].points[1].x, 100.f) && wb[0].points[0].y == 0.f && wb[0].points[2].y == 20.f,
           "ctc_word_boxes non-CJK");
    auto cj = oar::ctc_word_boxes(line, U"\u4e00\u4e8c", {0, 9}, 10, 5.0f, 5.0f);
    expect(cj.size() == 2 && near(cj[0].points[0].x, 0.f) && near(cj[0].points[1].x, 30.f) &&
               near(cj[1].points[0].x, 70.f) && near(cj[1].points[1].x, 100.f),
           "ctc_word_boxes CJK");
    expect(oar::ctc_word_boxes(line, U"", {1}, 10, 5.0f, 5.0f).empty(), "ctc_word_boxes empty text");
  }

  // Topk (utils/topk.rs:294-344): descending, stable, k clipped to the class count
  {
    const float p[3] = {0.1f, 0.8f, 0.1f};
    auto t = oar::topk_indices(p, 3, 2);
    expect(t.size() == 2 && t[0] == 1 && t[1] == 0, "topk order");
    expect(oar::topk_indices(p, 2, 5).size() == 2, "topk k larger than classes");
    try {
      oar::topk_indices(p, 3, 0);
      expect(false, "topk k = 0 rejected");
    } catch (const oar::OCRError&) {
    }
  }

  // layout post-process through the mirror (host only): threshold, same-class NMS, page-sized "image" box, per-label
  // threshold and merge mode resolved to class ids (layout_detection_adapter.rs:644-663), labels on the elements
  {
    const std::vector<std::string> labels = {"paragraph_title", "image", "text", "number", "abstract", "content",
                                             "figure_title", "formula", "table"};
    const std::vector<float> rows = {
        2, 0.90f, 0.10f, 0.10f, 0.50f, 0.30f,   // text
        2, 0.80f, 0.10f, 0.10f, 0.50f, 0.31f,   // near duplicate: suppressed
        1, 0.95f, 0.00f, 0.00f, 1.00f, 1.00f,   // page-sized image: dropped
        0, 0.45f, 0.12f, 0.12f, 0.30f, 0.20f,   // title inside the text box, below 0.5
        8, 0.60f, 0.55f, 0.55f, 0.95f, 0.95f};  // table
    oar::LayoutDetectionConfig lc;
    auto els = oar::postprocess_pp_doclayout(rows, 1, 5, 6, {{1000.0f, 800.0f}}, lc, labels);
    expect(els.size() == 1 && els[0].size() == 2 && els[0][0].element_type == "text" && els[0][1].element_type == "table" &&
               els[0][0].bbox.points[0].x == 100.0f && els[0][0].bbox.points[2].x == 500.0f,
           "layout defaults");
    lc.class_thresholds["paragraph_title"] = 0.3f;
    els = oar::postprocess_pp_doclayout(rows, 1, 5, 6, {{1000.0f, 800.0f}}, lc, labels);
    // (NMS hands the survivors back in score order: text 0.90, table 0.60, title 0.45)
    expect(els[0].size() == 3 && els[0][2].element_type == "paragraph_title", "layout per-label threshold");
    lc.class_merge_modes["text"] = oar::MergeBboxMode::Large;
    els = oar::postprocess_pp_doclayout(rows, 1, 5, 6, {{1000.0f, 800.0f}}, lc, labels);
    expect(els[0].size() == 2 && els[0][1].element_type == "table", "layout Large merge removes the contained title");
    try {
      oar::postprocess_pp_doclayout(rows, 2, 5, 6, {{1000.0f, 800.0f}}, lc, labels);
      expect(false, "layout batch mismatch rejected");
    } catch (const oar::OCRError&) {
    }
  }

  bool have_gpu = true;
  try {
    oar::Context probe(0);
  } catch (const oar::OCRError& e) {
    have_gpu = false;
    expect(e.code == OAR_E_NO_DEVICE && std::strstr(e.what(), "no CPU fallback"), "no-device error");
  }
  if (have_gpu && argc >= 3) {
    oar::Context ctx(0);
    auto db = slurp(argv[1]), rb = slurp(argv[2]);
    oar::Model det(ctx, db.data(), db.size()), rec(ctx, rb.data(), rb.size());
    oar::OAROCR ocr = oar::OAROCRBuilder(det, rec, 18385).image_batch_size(2).region_batch_size(8).build();
    try {
      ocr.predict({});
      expect(false, "empty input rejected");
    } catch (const oar::OCRError& e) {
      expect(std::strstr(e.what(), "non-empty slice") != nullptr, "empty input message");
    }
    std::vector<uint8_t> page(320 * 320 * 3, 240);
    for (int y = 100; y < 130; ++y)
      for (int x = 40; x < 280; ++x)
        for (int c = 0; c < 3; ++c) page[(y * 320 + x) * 3 + c] = (uint8_t)(30 + 40 * ((x / 4) & 1));
    auto res = ocr.predict({oar::RgbImage{page.data(), 320, 320}});
    expect(res.size() == 1 && res[0].text_regions.size() == 1, "one region on the striped page");
    if (res.size() == 1 && res[0].text_regions.size() == 1) {
      // return_word_box inputs arrive with the region: one CTC column per emitted character, T, the two ratios
      const oar::TextRegion& t = res[0].text_regions[0];
      expect(t.char_col_indices.size() == t.label_indices.size() && t.sequence_length > 0 && t.wh_ratio > 0.f &&
                 t.max_wh_ratio >= t.wh_ratio && t.max_wh_ratio >= 320.0f / 48.0f,
             "word-box inputs");
      std::u32string text(t.label_indices.size(), U'\u4e00');
      auto wb = oar::ctc_word_boxes(t.bounding_box, text, t.char_col_indices, (size_t)t.sequence_length, t.wh_ratio,
                                    t.max_wh_ratio);
      expect(wb.size() == t.label_indices.size(), "one word box per character");
    }
    auto dres = oar::TextDetectionPredictor(det).predict({oar::RgbImage{page.data(), 320, 320}});
    expect(dres.detections.size() == 1 && dres.detections[0].size() == 1, "detector predictor");
    if (res.size() == 1 && res[0].text_regions.size() == 1)
      expect(!res[0].text_regions[0].has_orientation_angle, "no classifier: orientation_angle is None");
    if (argc >= 4) {  // with_text_line_orientation_classification (ocr.rs:197-203)
      auto cb = slurp(argv[3]);
      oar::Model cls(ctx, cb.data(), cb.size());
      oar::OAROCR ocr2 = oar::OAROCRBuilder(det, rec, 18385).with_text_line_orientation_classification(cls).build();
      auto res2 = ocr2.predict({oar::RgbImage{page.data(), 320, 320}});
      expect(res2.size() == 1 && res2[0].text_regions.size() == 1 && res2[0].text_regions[0].has_orientation_angle &&
                 (res2[0].text_regions[0].orientation_angle == 0.0f || res2[0].text_regions[0].orientation_angle == 180.0f),
             "classifier attached: angle 0 or 180");
      std::vector<uint8_t> line(48 * 200 * 3, 90);
      auto ori = oar::TextLineOrientationPredictor(cls).predict({oar::RgbImage{line.data(), 48, 200}});
      expect(ori.orientations.size() == 1 && ori.orientations[0].size() == 2 &&
                 ori.orientations[0][0].score >= ori.orientations[0][1].score &&
                 (ori.orientations[0][0].label == "0" || ori.orientations[0][0].label == "180"),
             "orientation predictor: top-2, best first");
    }
    std::printf("gpu path ok: %zu regions\n", res[0].text_regions.size());
  }
  std::printf(failures ? "FAILED %d\n" : "ok\n", failures);
  return failures ? 1 : 0;
}
