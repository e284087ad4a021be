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
