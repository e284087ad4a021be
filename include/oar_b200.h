/* oar_b200.h -- C ABI of the B200-native OCR hot path (liboar_b200.so).
 *
 * Drop-in boundary for oar-ocr's det+rec path.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repo,
 * GreatV/oar-ocr v0.9.3).  Plain pointers and sizes only; no exceptions cross
 * the boundary.  Every function returns 0 on success or a negative OAR_E_*
 * code; oar_last_error() returns a thread-local UTF-8 message that a Rust
 * shim wraps in OCRError::Inference{model_name, context, source}
 * (oar-ocr-core/src/core/errors/types.rs:110-214).
 *
 * Threading mirrors OrtInfer's Mutex<Session>
 * (oar-ocr-core/src/core/inference/ort_infer_execution.rs:142-155): calls on
 * the same context serialise internally; different contexts (one per GPU) run
 * concurrently.
 *
 * There is no CPU fallback: every compute entry point fails with
 * OAR_E_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef OAR_B200_H
#define OAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAR_OK 0
#define OAR_E_INVALID (-1)   /* OCRError::InvalidInput / validation_error */
#define OAR_E_NO_DEVICE (-2) /* no usable CUDA device */
#define OAR_E_CUDA (-3)      /* CUDA runtime failure -> OCRError::Inference */
#define OAR_E_MODEL (-4)     /* malformed model blob -> OCRError::ModelLoad */
#define OAR_E_CAPACITY (-5)  /* caller-provided output buffer too small */
#define OAR_E_UNSUPPORTED (-6)

typedef struct oar_ctx oar_ctx;     /* one device + stream + workspace arena */
typedef struct oar_model oar_model; /* one network resident on a context */

#define OAR_KIND_DET 0
#define OAR_KIND_REC 1
#define OAR_KIND_CLS 2 /* PP-LCNet classifier: text-line orientation */
#define OAR_KIND_FEAT 3 /* feature extractor (HGNetV2 backbone, layout-detector encoder): oar_infer_f32 only */

/* Detection post-process configuration.
 * = DBPostProcess{thresh, box_thresh, max_candidates, unclip_ratio, min_size}
 *   (oar-ocr-core/src/processors/db_postprocess.rs:69-88) + the limit-side
 *   resize parameters of DetResizeForTest
 *   (oar-ocr-core/src/processors/resize_detection.rs:243-319).
 * Fixed as both reference adapters fix them: use_dilation=false,
 * ScoreMode::Fast, BoxType::Quad
 * (oar-ocr-core/src/domain/adapters/text_detection_adapter.rs:165-173). */
typedef struct {
  float thresh;            /* 0.3 */
  float box_thresh;        /* 0.6 */
  float unclip_ratio;      /* 2.0 (OAROCRBuilder default, src/oarocr/ocr.rs:351-364); 1.5 in the predictor */
  int32_t max_candidates;  /* 1000 */
  float min_size;          /* 3.0 */
  int32_t limit_side_len;  /* 960 */
  int32_t limit_type;      /* 0 = Max, 1 = Min, 2 = ResizeLong */
  int32_t max_side_limit;  /* 4000 */
} oar_det_config;

/* Pipeline configuration = the knobs OAROCRBuilder exposes for this path
 * (src/oarocr/ocr.rs:249-417). */
typedef struct {
  oar_det_config det;
  int32_t image_batch_size;  /* det_batch_size, ocr.rs:550-557 (accelerator default 8) */
  int32_t region_batch_size; /* ocr.rs:818-825 (accelerator default 64) */
  float rec_score_thresh;    /* TextRecognitionConfig.score_threshold, text_recognition_adapter.rs:88-102 */
  int32_t n_chars;           /* len(CTCLabelDecode.character) = dict + blank + space, decode.rs:392-423 */
} oar_pipeline_config;

void oar_det_config_default(oar_det_config* cfg);           /* ocr.rs:351-364 defaults */
void oar_pipeline_config_default(oar_pipeline_config* cfg); /* accelerator defaults 8 / 64 */

const char* oar_last_error(void);
int32_t oar_version(void);
/* number of kernel launches issued by this library on the calling process so far */
int64_t oar_launch_count(void);
/* number of host-visible submissions among them: a kernel launched directly counts one, a replayed CUDA graph of a
 * network's layer list (one per detector / recogniser batch after its shape has been seen twice) counts one */
int64_t oar_submit_count(void);

/* ---- context / model lifetime -------------------------------------------
 * replaces OrtInfer::new / from_config + Session::builder().commit_from_memory
 * (oar-ocr-core/src/core/inference/ort_infer_builders.rs:9-70, session.rs:21-46) */
int32_t oar_ctx_create(int32_t device_id, oar_ctx** out);
void oar_ctx_destroy(oar_ctx* ctx);
int32_t oar_ctx_synchronize(oar_ctx* ctx);
/* Loads an OARG layer-list blob (oar_ocr_b200/models.py); weights are copied to HBM. */
int32_t oar_model_load_blob(oar_ctx* ctx, const void* bytes, size_t len, oar_model** out);
/* Structural validation of an OARG blob without a device (what oar_model_load_blob checks before touching the GPU):
 * sizes bounded by `len`, tensor ids and weight slices in range, parameters positive, every weight slice exactly as
 * long as its op's parameters imply.  OAR_OK or OAR_E_MODEL with the reason in oar_last_error(). */
int32_t oar_model_validate_blob(const void* bytes, size_t len);
/* ModelSource::Memory / ModelSource::Path contents as the reference hands them to ONNX Runtime
 * (oar-ocr-core/src/core/config/model_source.rs:20-28, core/inference/ort_infer_builders.rs:9-70,
 * Session::builder().commit_from_memory): ONNX ModelProto bytes (or an OARG blob, recognised by its magic).
 * `kind` = OAR_KIND_DET / OAR_KIND_REC / OAR_KIND_CLS states the caller's role as the reference's per-task builders do
 * (a mismatch with the graph is OAR_E_MODEL), -1 = take whatever the graph is.  Operators outside the supported subset
 * fail with OAR_E_MODEL naming the node -- the ONNX graph is converted to the layer list the CUDA engine executes, it
 * is never interpreted on the CPU. */
int32_t oar_model_load_onnx(oar_ctx* ctx, const void* bytes, size_t len, int32_t kind, oar_model** out);
/* The conversion alone, host only (no device needed): writes the OARG blob to `out` (capacity `cap`) and its size to
 * `out_len`; out == NULL queries the size.  kind_hint = OAR_KIND_CLS marks a classifier (its MatMul + Softmax tail is
 * otherwise read as a CTC head), -1 = infer. */
int32_t oar_onnx_to_oarg(const void* onnx, size_t len, int32_t kind_hint, void* out, size_t cap, size_t* out_len);
void oar_model_destroy(oar_model* m);
int32_t oar_model_kind(const oar_model* m);
/* 0 = fp32 SIMT reference engine, 1 = tcgen05 tensor-core engine (one kernel per layer), 2 = tcgen05 engine with the
 * depthwise->pointwise blocks fused into one persistent kernel (default) */
int32_t oar_model_set_engine(oar_model* m, int32_t engine);

/* ---- seam 1: OrtInfer::infer / infer_first_output_f32 --------------------
 * (oar-ocr-core/src/core/inference/ort_infer_execution.rs:121-306)
 * in: host f32 [B,3,H,W] row-major (the tensor the reference feeds as "x").
 * det out: f32 [B,1,H,W] probability map; rec out: f32 [B,T,V] softmax; cls out: f32 [B,classes] softmax.
 * out_shape receives 4 dims (rec: B,T,V,1; cls: B,1,classes,1).  out_cap in floats. */
int32_t oar_infer_f32(oar_model* m, const float* in, const int64_t in_shape[4], float* out, size_t out_cap,
                      int64_t out_shape[4]);

/* ---- row 2: NormalizeImage::normalize_batch_refs --------------------------
 * (oar-ocr-core/src/processors/normalization.rs:429-482, simd.rs:28-45)
 * u8 HWC RGB [B,H,W,3] -> f32 NCHW, out[c] = rgb[src[c]]*alpha[c]+beta[c]
 * (separate multiply and add, as the reference). Host pointers. */
int32_t oar_normalize_chw(oar_ctx* ctx, const uint8_t* rgb, int32_t batch, int32_t h, int32_t w,
                          const int32_t src_channels[3], const float alpha[3], const float beta[3], float* out);

/* ---- rows 4-9: DBPostProcess::apply ---------------------------------------
 * (oar-ocr-core/src/processors/db_postprocess.rs:100-179, db_bitmap.rs:84-150)
 * pred: host f32 [B,H,W]; src_h/src_w: ImageScaleInfo source dims per image.
 * boxes: [B][max_candidates][4][2] (discovery order), scores, counts[B]. */
int32_t oar_db_postprocess(oar_ctx* ctx, const float* pred, int32_t batch, int32_t h, int32_t w, const int32_t* src_h,
                           const int32_t* src_w, const oar_det_config* cfg, float* boxes, float* scores,
                           int32_t* counts);

/* ---- seam 2 (detection): TextDetectionAdapter::execute --------------------
 * (oar-ocr-core/src/domain/adapters/text_detection_adapter.rs:36-79 ->
 *  DBModel::forward, oar-ocr-core/src/models/detection/db.rs:281-335)
 * images: n host pointers to u8 HWC RGB; boxes [n][max_candidates][4][2] in
 * source-image coordinates, discovery order (unsorted, as the adapter returns). */
int32_t oar_det_run(oar_model* det, const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                    const oar_det_config* cfg, float* boxes, float* scores, int32_t* counts);

/* ---- row 10: sort_quad_boxes (oar-ocr-core/src/processors/sorting.rs:35-84) */
int32_t oar_sort_quad_boxes(float* boxes, int32_t n, int32_t* order);

/* ---- row 11: get_rotate_crop_image ----------------------------------------
 * (oar-ocr-core/src/utils/transform.rs:76-191).  One image, n quads.
 * Phase 1 (out == NULL): fills out_w/out_h/status per quad (status != 0 where
 * the reference returns Err).  Phase 2: out = concatenated u8 HWC crops. */
int32_t oar_rotate_crop(oar_ctx* ctx, const uint8_t* image, int32_t h, int32_t w, const float* quads, int32_t n,
                        int32_t* out_w, int32_t* out_h, int32_t* status, uint8_t* out, size_t out_cap);

/* ---- row 13: CRNNModel::preprocess_refs -----------------------------------
 * (oar-ocr-core/src/models/recognition/crnn.rs:71-125, simd.rs:248-308)
 * crops: n host u8 HWC images; out f32 [n,3,48,tensor_w]; *tensor_w returned. */
int32_t oar_crnn_preprocess(oar_ctx* ctx, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws,
                            int32_t n, float* out, size_t out_cap, int32_t* tensor_w);

/* ---- rows 15-16: CTCLabelDecode::argmax_predictions + decode_argmax --------
 * (oar-ocr-core/src/processors/decode.rs:452-614, simd.rs:190-229)
 * pred: host f32 [B,T,V].  idx/prob [B,T]; labels/cols [B,T] (first len[b] valid). */
int32_t oar_ctc_decode(oar_ctx* ctx, const float* pred, int32_t b, int32_t t, int32_t v, int32_t n_chars,
                       int32_t* idx, float* prob, int32_t* labels, int32_t* cols, int32_t* lens, float* scores);

/* ---- seam 2 (recognition): TextRecognitionAdapter::execute ----------------
 * (oar-ocr-core/src/domain/adapters/text_recognition_adapter.rs:35-111 ->
 *  CRNNModel::forward_refs, crnn.rs:247-293).  The whole input is ONE batch.
 * labels/cols: [n][t_cap]; *t_out = sequence length T of this batch. */
int32_t oar_rec_run(oar_model* rec, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                    int32_t n_chars, int32_t* labels, int32_t* cols, int32_t* lens, float* scores, int32_t t_cap,
                    int32_t* t_out);

/* ---- the hot path: OAROCR::predict (src/oarocr/ocr.rs:518-659) -------------
 * Results, per input image i: regions [region_off[i], region_off[i+1]) in
 * reading order (sort_quad_boxes), dropped where the crop failed
 * (processors.rs:104-106).  For region r: box[r][4][2], score[r], CTC-collapsed
 * class indices labels[label_off[r] .. label_off[r+1]) (text = chars[index];
 * empty when score < rec_score_thresh), det_index[r] = detection_index. */
typedef struct {
  int32_t cap_regions; /* in: capacity of per-region arrays */
  int32_t cap_labels;  /* in: capacity of labels */
  int32_t* region_off; /* [n_images+1] */
  float* boxes;        /* [cap_regions][8] */
  float* scores;       /* [cap_regions] recognition confidence */
  int32_t* det_index;  /* [cap_regions] */
  int32_t* label_off;  /* [cap_regions+1] */
  int32_t* labels;     /* [cap_labels] */
  /* device-time breakdown of the last call, ms (0 when not measured) */
  float ms_h2d, ms_det, ms_post, ms_crop, ms_rec, ms_total;
  int64_t h2d_bytes, d2h_bytes;
  /* optional (may be NULL): what OAROCR::ctc_word_boxes needs per region when `return_word_box` is on
   * (src/oarocr/ocr.rs:827-868, 949-1022).  cols[label_off[r] .. label_off[r+1]) = CTC timestep of each emitted
   * character (char_col_indices), seq_len[r] = T of the recognition batch the region was in, wh_ratio[r] = the crop's
   * w / max(h,1), max_wh_ratio[r] = max(320/48, max wh_ratio of that batch) (chunk_max_wh_ratio). */
  int32_t* cols;       /* [cap_labels] */
  int32_t* seq_len;    /* [cap_regions] */
  float* wh_ratio;     /* [cap_regions] */
  float* max_wh_ratio; /* [cap_regions] */
  /* optional (may be NULL): TextRegion.orientation_angle of the line-orientation stage (ocr.rs:755-792, 888):
   * 0 or 180 when oar_pipeline_run_cls ran with a classifier, -1 (None) otherwise */
  float* line_angle;   /* [cap_regions] */
  float ms_cls;        /* device time of the orientation stage (inside ms_rec's bracket: it runs after the crop) */
} oar_ocr_result;

/* images_on_device != 0: `images` are device pointers already resident in HBM */
int32_t oar_pipeline_run(oar_model* det, oar_model* rec, const uint8_t* const* images, const int32_t* hs,
                         const int32_t* ws, int32_t n, int32_t images_on_device, const oar_pipeline_config* cfg,
                         oar_ocr_result* out);

/* ---- text-line orientation ---------------------------
 * TextLineOrientationAdapter::execute -> PPLCNetModel::forward_refs
 * (oar-ocr-core/src/domain/adapters/text_line_orientation_adapter.rs:63-121,
 *  oar-ocr-core/src/models/classification/pp_lcnet.rs:139-196, 255-300): every crop is resized straight to
 * input_w x input_h (Triangle, resize_short = None), normalised with scale 1/255 and the ImageNet mean/std in RGB
 * order, classified, and reduced by Topk (oar-ocr-core/src/utils/topk.rs).  crops: n host u8 HWC images.
 * class_ids/scores [n]: the top-1 (first maximal class, as the stable sort yields); probs (may be NULL) receives the
 * full [n][*n_classes] rows so the caller can form any top-k; *n_classes is always returned.
 * Unlike the reference's filter_map (pp_lcnet.rs:182-186), which silently drops zero-sized images and so shifts every
 * later result, an empty crop is rejected with OAR_E_INVALID. */
int32_t oar_cls_run(oar_model* cls, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                    int32_t input_h, int32_t input_w, int32_t* class_ids, float* scores, float* probs,
                    size_t probs_cap, int32_t* n_classes);

/* image::imageops::rotate180 as classify_line_orientations applies it to a crop of class 1 (ocr.rs:785-788).
 * image/out: host u8 HWC [h][w][3]. */
int32_t oar_rotate180(oar_ctx* ctx, const uint8_t* image, int32_t h, int32_t w, uint8_t* out);

/* OAROCR::predict with with_text_line_orientation_classification (ocr.rs:197-203, 615): as oar_pipeline_run, and
 * between cropping and recognition every crop is classified (input 80 x 160, DEFAULT_INPUT_SHAPE) and the crops of
 * class 1 are rotated by 180 degrees in HBM; wh_ratio keeps its pre-rotation value.  cls == NULL: exactly
 * oar_pipeline_run. */
int32_t oar_pipeline_run_cls(oar_model* det, oar_model* rec, oar_model* cls, const uint8_t* const* images,
                             const int32_t* hs, const int32_t* ws, int32_t n, int32_t images_on_device,
                             const oar_pipeline_config* cfg, oar_ocr_result* out);

/* ---- layout detection, host half ----------------------
 * LayoutDetectionAdapter::postprocess_pp_doclayout
 * (oar-ocr-core/src/domain/adapters/layout_detection_adapter.rs:631-846) with its helpers convert_bbox_coords
 * (:848-878), paddlex_layout_nms (:884-951), filter_large_image_boxes (:953-992), apply_paddlex_merge_modes /
 * check_containment / is_contained (:994-1106), and unclip_boxes
 * (oar-ocr-core/src/processors/layout_postprocess.rs:636-681).  Host code in the reference and here: it works on the
 * few hundred rows the detector returns, needs no context and no device.
 * Configuration = LayoutDetectionConfig (oar-ocr-core/src/domain/tasks/layout_detection.rs:45-100) with the label-keyed
 * maps resolved to class ids by the caller (the adapter does the same through LayoutModelConfig.class_labels). */
#define OAR_MERGE_UNSET (-1)
#define OAR_MERGE_LARGE 0 /* MergeBboxMode::Large */
#define OAR_MERGE_SMALL 1
#define OAR_MERGE_UNION 2
#define OAR_UNCLIP_NONE 0      /* layout_unclip_ratio = None */
#define OAR_UNCLIP_RATIO 1     /* UnclipRatio::Uniform(r) -> (r, r); Separate(w, h) */
#define OAR_UNCLIP_PER_CLASS 2 /* UnclipRatio::PerClass */
typedef struct {
  float score_threshold;            /* 0.5 */
  int32_t max_elements;             /* 100 */
  int32_t layout_nms;               /* 1 */
  int32_t num_classes;              /* LayoutModelConfig.num_classes (23 for PP-DocLayout-L) */
  const float* class_thresholds;    /* [num_classes], NaN = not configured; NULL = None */
  const int32_t* class_merge_modes; /* [num_classes] OAR_MERGE_*; NULL = None */
  int32_t image_class_id;           /* id of the label "image", -1 if the model has none */
  int32_t formula_class_id;         /* id of the label "formula", -1 if none */
  int32_t unclip_mode;              /* OAR_UNCLIP_* */
  float unclip_w, unclip_h;         /* OAR_UNCLIP_RATIO */
  const float* class_unclip;        /* OAR_UNCLIP_PER_CLASS: [num_classes][2] (w, h), NaN = (1, 1) */
} oar_layout_config;
void oar_layout_config_default(oar_layout_config* cfg); /* LayoutDetectionConfig::default + PP-DocLayout-L ids */
/* pred: host f32 [batch][num_boxes][feature_dim] rows [class_id, score, x1, y1, x2, y2, (order keys)] exactly as the
 * adapter reads them (:689-690; 7 columns = one reading-order key, 8 = (column, row)); src_w/src_h: ImageScaleInfo
 * source dims per image.  Outputs per image: up to max_elements rows -- boxes [batch][max_elements][4] as
 * (x1, y1, x2, y2) = the arguments of BoundingBox::from_coords, class ids, scores -- and counts[batch]. */
int32_t oar_layout_postprocess(const float* pred, int32_t batch, int32_t num_boxes, int32_t feature_dim,
                               const float* src_w, const float* src_h, const oar_layout_config* cfg, float* boxes,
                               int32_t* classes, float* scores, int32_t* counts);

/* ---- seam 2 (recognition from pages + boxes): the second half of OAROCR::predict --------------------------------
 * (get_rotate_crop_image, oar-ocr-core/src/utils/transform.rs:76-502 -> recognize_global, src/oarocr/ocr.rs:802-897
 *  -> TextRecognitionAdapter::execute, text_recognition_adapter.rs:35-111).  SURVEY.md 8b: oar_crop_rec_run.
 * images: n pages (host, or device pointers when images_on_device != 0); boxes [n_boxes][4][2] in page coordinates,
 * box k cropped from page img_index[k].  The crops are pooled in the caller's box order (flush at 4096), stably sorted
 * by width / height and recognised in batches of region_batch_size -- exactly as predict() batches its own boxes.
 * Per box k: status[k] = 0 ok, 1 = the crop failed (the reference skips it, processors.rs:104-106);
 * labels/cols [k][t_cap] (first lens[k] valid; empty when scores[k] < rec_score_thresh), seq_len[k] (may be NULL) =
 * sequence length T of the batch box k was recognised in. */
int32_t oar_crop_rec_run(oar_model* rec, const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                         int32_t images_on_device, const float* boxes, const int32_t* img_index, int32_t n_boxes,
                         int32_t region_batch_size, int32_t n_chars, float rec_score_thresh, int32_t* status,
                         int32_t* labels, int32_t* cols, int32_t* lens, float* scores, int32_t* seq_len, int32_t t_cap);

/* oar_rec_run with crops that may already be resident in HBM (crops_on_device != 0: device pointers, u8 HWC).  Host
 * crops are packed into one pinned buffer and uploaded with a single copy. */
int32_t oar_rec_run_ex(oar_model* rec, const uint8_t* const* crops, const int32_t* hs, const int32_t* ws, int32_t n,
                       int32_t crops_on_device, int32_t n_chars, int32_t* labels, int32_t* cols, int32_t* lens,
                       float* scores, int32_t t_cap, int32_t* t_out);

/* ---- OAROCR::predict on several GPUs inside one process (one host thread + CUDA context per GPU) ----
 * dets[g] / recs[g]: the same two models loaded on n_ctx different contexts (normally one per device).  Equal to ONE
 * un-sharded oar_pipeline_run over all pages: pages are dealt to the contexts in contiguous blocks for upload,
 * detection and cropping; recognize_global is planned once over ALL crops and whole batches are dealt round-robin, a
 * context fetching the crops that live in another context's HBM over the peer link.  No collective on the path.
 * Host pages only.  Stage times in `out` are those of the slowest context. */
int32_t oar_pipeline_run_multi(oar_model* const* dets, oar_model* const* recs, int32_t n_ctx,
                               const uint8_t* const* images, const int32_t* hs, const int32_t* ws, int32_t n,
                               const oar_pipeline_config* cfg, oar_ocr_result* out);

/* ---- encoded pages -----------------------------------------------------------------------
 * In the reference a page enters through load_image (oar-ocr-core/src/core/utils/image.rs:88: image::open -> RGB8) on
 * the CPU.  oar_pipeline_run_encoded takes the JPEG streams themselves: nvJPEG decodes them into HBM (interleaved RGB)
 * and OAROCR::predict proceeds as oar_pipeline_run does on device-resident pages.  nvJPEG is looked up at first use;
 * without it the call fails with OAR_E_UNSUPPORTED (there is no CPU decode).  out_hs / out_ws (may be NULL): decoded
 * page sizes.  oar_decode_jpeg returns one decoded page to the host (rgb == NULL: size only). */
int32_t oar_pipeline_run_encoded(oar_model* det, oar_model* rec, const uint8_t* const* jpegs, const size_t* lens, int32_t n,
                                 const oar_pipeline_config* cfg, oar_ocr_result* out, int32_t* out_hs, int32_t* out_ws);
int32_t oar_decode_jpeg(oar_ctx* ctx, const uint8_t* jpeg, size_t len, uint8_t* rgb, size_t rgb_cap, int32_t* h, int32_t* w);

/* ---- layout detection on the device (BASELINE.json configs[4]) --------------------------------
 * LayoutDetectionAdapter::execute (oar-ocr-core/src/domain/adapters/layout_detection_adapter.rs:1128-1197) ->
 * ScaleAwareDetectorModel::preprocess / infer (models/detection/scale_aware_detector.rs:150-333) for the PP-DocLayout
 * family: resize_exact to (input_h, input_w) with FilterType::CatmullRom, scale 1/255, RGB; RT-DETR-L
 * (HGNetV2-L, hybrid encoder, 6-layer deformable decoder; oar-ocr-vl/src/models/pp_doclayout/) ; the exported model's
 * own tail (sigmoid, top-300 over (query, class), cxcywh -> xyxy scaled to the source image).
 * encoder / head: two OAR_KIND_FEAT models on one context (oar_ocr_b200.models.build_layout_encoder for this input
 * size, build_layout_head).  oar_layout_rows returns the model's output tensor, rows [n][300][6] =
 * [class_id, score, x1, y1, x2, y2] -- what the adapter reads (images_on_device != 0: `images` are device pointers);
 * oar_layout_run (host pages) feeds it to oar_layout_postprocess. */
int32_t oar_layout_rows(oar_model* encoder, oar_model* head, const uint8_t* const* images, const int32_t* hs,
                        const int32_t* ws, int32_t n, int32_t images_on_device, int32_t input_h, int32_t input_w,
                        float* rows, size_t rows_cap);
int32_t oar_layout_run(oar_model* encoder, oar_model* head, const uint8_t* const* images, const int32_t* hs,
                       const int32_t* ws, int32_t n, int32_t input_h, int32_t input_w, const oar_layout_config* cfg,
                       float* boxes, int32_t* classes, float* scores, int32_t* counts);

/* device memory helpers for callers that keep inputs resident (bench `value` leg) */
int32_t oar_device_alloc(oar_ctx* ctx, size_t bytes, void** out);
int32_t oar_device_free(oar_ctx* ctx, void* p);
int32_t oar_memcpy_h2d(oar_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- measurement helpers (bench.py) ----------------------------------------
 * CUDA-event stopwatch on the context's own launch stream: start records an
 * event, stop records another, synchronises and returns the elapsed device
 * time.  l2_flush overwrites a scratch buffer larger than the 126 MB L2. */
int32_t oar_timer_start(oar_ctx* ctx);
int32_t oar_timer_stop(oar_ctx* ctx, float* ms);
int32_t oar_l2_flush(oar_ctx* ctx);

/* per-kernel timing of the most recent oar_infer / pipeline call on a model's
 * context: enables cudaEvent brackets around every launch (bench roofline leg) */
int32_t oar_profile_enable(oar_ctx* ctx, int32_t on);
/* writes up to cap records; returns the count.  name: static string */
typedef struct {
  const char* name;
  float ms;
  double flops;
  double bytes;
} oar_kernel_record;
int32_t oar_profile_read(oar_ctx* ctx, oar_kernel_record* recs, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* OAR_B200_H */
