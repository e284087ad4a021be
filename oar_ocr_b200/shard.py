"""Replica data-parallelism for the hot path: images shard across ranks as contiguous blocks, every
rank runs the whole pipeline on its block, results are concatenated in input order.  No collective touches image
or activation data; the only communication is the optional gather of the (small) results.

An R-rank run is exactly R independent `predict()` calls, one per block: crop pooling and the wh-ratio sort of
recognize_global are per call (src/oarocr/ocr.rs:603-633, 811-827), so block membership decides the recognition
batches.  The CPU oracle must therefore be sharded the same way when it checks a multi-GPU run.
"""
from __future__ import annotations


def block_partition(n: int, world: int) -> list[tuple[int, int]]:
    """contiguous blocks, sizes differing by at most one, earlier ranks take the remainder"""
    if world <= 0:
        raise ValueError("world size must be positive")
    base, rem = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def predict_sharded(predict_fn, images, rank: int, world: int, gather: bool = True):
    """Runs `predict_fn(block)` on this rank's block.  With gather=True (needs an initialised torch.distributed
    process group) every rank returns the full result list in input order; otherwise only its own block's results.
    A rank whose block is empty makes no call (OAROCR::predict rejects an empty list, ocr.rs:525-532)."""
    start, end = block_partition(len(images), world)[rank]
    local = predict_fn(images[start:end]) if end > start else []
    if not gather or world == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    out = []
    for p in parts:
        out.extend(p)
    return out


# ---------------------------------------------------------------------------------------------------------------
# Global crop pooling across ranks: an R-rank run that equals ONE un-sharded predict()
# ---------------------------------------------------------------------------------------------------------------
MAX_POOLED_CROPS = 4096  # src/oarocr/ocr.rs:603


def plan_pooled_chunks(meta, region_batch_size: int, max_pooled: int = MAX_POOLED_CROPS):
    """The batch composition of OAROCR::recognize_global (ocr.rs:603-633, 802-827) for the WHOLE image list.

    `meta` lists every successfully cropped region in (image index, detection index) order as
    (image index, detection index, wh_ratio).  Regions are pooled in that order; a pool is flushed when it holds
    `max_pooled` crops (and once at the end); a flush stable-sorts its pool by wh_ratio ascending and cuts it into
    chunks of `region_batch_size`.  Returns the chunks, each a list of positions into `meta`, in processing order."""
    if region_batch_size <= 0:
        raise ValueError("region_batch_size must be positive")
    chunks, pool = [], []

    def flush():
        order = sorted(pool, key=lambda i: meta[i][2])  # Python's sort is stable, like slice::sort_by
        for s in range(0, len(order), region_batch_size):
            chunks.append(order[s:s + region_batch_size])

    for i in range(len(meta)):
        pool.append(i)
        if len(pool) >= max_pooled:
            flush()
            pool = []
    if pool:
        flush()
    return chunks


def predict_pooled(stages, images, rank: int, world: int, region_batch_size: int, rec_score_thresh: float = 0.0,
                   max_pooled: int = MAX_POOLED_CROPS):
    """Det + crop on this rank's contiguous block of images, then recognition batches built from the GLOBAL pool of
    crops exactly as one un-sharded OAROCR::predict would build them, dealt to ranks as whole chunks.

    Block sharding alone (predict_sharded) changes which crops share a recognition batch, hence `tensor_w` and the
    padded logits; here every chunk has the membership, order and width of the single-process run, so results are
    identical to it.  One exchange step: region sizes are all-gathered (a few bytes per region), each rank receives
    the u8 crops of the chunks it was dealt (gather_object per destination; NVLink all-to-all on a GPU box would
    carry the same payload), results are all-gathered.  world == 1 needs no process group.

    `stages` provides detect(images) -> per image boxes [n,4,2] in reading order; crop(image, boxes) -> list of
    HxWx3 u8 crops (None where cropping fails); recognize(crops) -> list of (labels, score) for ONE batch; and
    optionally orient(crops) -> class id per crop + rotate180(crop), the text-line orientation stage
    (classify_line_orientations, ocr.rs:615, 755-792): it runs per page on the rank that cropped it, before the
    exchange, so rotated crops are what travels; wh_ratio keeps its pre-rotation value as in the reference.
    Returns, on every rank, per image a list of dicts {box, det_index, labels, score, angle} in detection order
    (angle = None without an orientation stage)."""
    import numpy as np
    start, end = block_partition(len(images), world)[rank]
    block = images[start:end]
    boxes = stages.detect(block) if block else []
    local_crops, local_meta, local_boxes, local_angles = {}, [], {}, {}
    orient = getattr(stages, "orient", None)
    for k, (img, b) in enumerate(zip(block, boxes)):
        gi = start + k
        local_boxes[gi] = np.asarray(b, np.float32).reshape(-1, 4, 2)
        crops = stages.crop(img, local_boxes[gi]) if len(local_boxes[gi]) else []
        page = []
        for d, c in enumerate(crops):
            if c is None:
                continue
            ratio = float(np.float32(c.shape[1]) / np.float32(max(c.shape[0], 1)))  # ocr.rs:739, f32
            page.append((d, c))
            local_meta.append((gi, d, ratio))
        if orient is not None and page:
            for (d, c), cid in zip(page, orient([c for _, c in page])):
                local_angles[(gi, d)] = float(int(cid)) * 180.0
                local_crops[(gi, d)] = stages.rotate180(c) if int(cid) == 1 else c
        else:
            for d, c in page:
                local_crops[(gi, d)] = c
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, (local_meta, local_boxes, local_angles))
        meta = [m for p in parts for m in p[0]]  # blocks are contiguous: rank order is image order
        all_boxes, angles = {}, {}
        for p in parts:
            all_boxes.update(p[1])
            angles.update(p[2])
    else:
        meta, all_boxes, angles = local_meta, local_boxes, local_angles
    chunks = plan_pooled_chunks(meta, region_batch_size, max_pooled)
    owner = [j % world for j in range(len(chunks))]  # whole chunks, round robin
    # crops travel once, to the rank that recognises their chunk
    if world > 1:
        need = {}
        for dst in range(world):
            mine = {}
            for j, ch in enumerate(chunks):
                if owner[j] != dst:
                    continue
                for pos in ch:
                    key = (meta[pos][0], meta[pos][1])
                    if key in local_crops:
                        mine[key] = local_crops[key]
            got = [None] * world if rank == dst else None
            dist.gather_object(mine, got, dst=dst)
            if rank == dst:
                for g in got:
                    need.update(g)
    else:
        need = local_crops
    out = []
    for j, ch in enumerate(chunks):
        if owner[j] != rank:
            continue
        keys = [(meta[pos][0], meta[pos][1]) for pos in ch]
        rec = stages.recognize([need[k] for k in keys])
        for (gi, d), (labels, score) in zip(keys, rec):
            labels = np.asarray(labels)
            if not (float(score) >= rec_score_thresh):  # text_recognition_adapter.rs:88-102: text dropped, index kept
                labels = labels[:0]
            out.append((gi, d, labels, float(score)))
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, out)
        out = [r for p in parts for r in p]
    results = [[] for _ in images]
    for gi, d, labels, score in sorted(out, key=lambda r: (r[0], r[1])):
        results[gi].append(dict(box=all_boxes[gi][d], det_index=d, labels=labels, score=score,
                                angle=angles.get((gi, d))))
    return results


class GpuStages:
    """predict_pooled stages on the CUDA library (stage entry points of include/oar_b200.h)."""

    def __init__(self, ocr):
        self.ocr = ocr

    def detect(self, images):
        from . import ffi
        out = []
        # chunks of image_batch_size, as TextDetectionAdapter::execute batches them (a page's boxes do not depend on
        # its batch mates; the chunking bounds the activation arena by the configured batch, not by the rank's block)
        step = max(1, int(getattr(self.ocr, "image_batch_size", 0) or len(images) or 1))
        for s0 in range(0, len(images), step):
            for boxes, _scores in self.ocr.det.det_run(images[s0:s0 + step], self.ocr.det_cfg.to_ffi()):
                out.append(ffi.sort_quad_boxes(boxes)[0] if len(boxes) else boxes)
        return out

    def crop(self, image, boxes):
        return self.ocr.ctx.rotate_crop(image, boxes)

    def recognize(self, crops):
        r = self.ocr.rec.rec_run(crops, len(self.ocr.chars))
        return list(zip(r["labels"], [float(s) for s in r["scores"]]))

    def __getattr__(self, name):
        # the orientation stage exists only when the pipeline was built with a classifier
        # (OAROCRBuilder::with_text_line_orientation_classification)
        if name == "orient" and self.__dict__.get("ocr") is not None and self.ocr.cls is not None:
            return lambda crops: self.ocr.cls.cls_run(crops, (80, 160), want_probs=False)["class_ids"]
        raise AttributeError(name)

    def rotate180(self, crop):
        return self.ocr.ctx.rotate180(crop)
