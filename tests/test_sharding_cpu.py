"""Multi-rank host logic on CPU: gloo, world size 2.  The per-rank predictor here is the CPU oracle (there is no
GPU in this container); what is under test is the sharding contract of oar_ocr_b200/shard.py: contiguous blocks,
input-order gather, and that a 2-rank run equals two sharded predict() calls."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oar_ocr_b200.shard import block_partition, predict_sharded


def test_block_partition():
    assert block_partition(256, 8) == [(32 * r, 32 * r + 32) for r in range(8)]
    assert block_partition(5, 2) == [(0, 3), (3, 5)]
    assert block_partition(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert block_partition(0, 2) == [(0, 0), (0, 0)]
    with pytest.raises(ValueError):
        block_partition(4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _summarise(res):
    return [[(r["box"].tolist(), r["labels"].tolist(), round(float(r["score"]), 6)) for r in img] for img in res]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    torch.set_num_threads(2)
    from oar_ocr_b200 import models, synth
    from oracle import pipeline
    from oracle.net import OracleNet
    det, rec = OracleNet(models.get_blob("det")), OracleNet(models.get_blob("rec"))
    images = [synth.page(80 + i, 320) for i in range(3)]

    def predict(block):
        return _summarise(pipeline.predict(det, rec, block, 18385, image_batch_size=2, region_batch_size=4))

    full = predict_sharded(predict, images, rank, world)
    if rank == 0:
        # reference for the sharded run: one predict() per block, concatenated
        want = []
        for s, e in block_partition(len(images), world):
            want.extend(predict(images[s:e]))
        ok = full == want and len(full) == len(images)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([ok, sum(len(x) for x in full)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_predict_matches_per_block_oracle(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ok, n_regions = np.load(tmp_path / "ok.npy")
    assert ok == 1 and n_regions >= 3
