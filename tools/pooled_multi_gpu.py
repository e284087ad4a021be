"""Cross-rank global crop pooling on real GPUs: `torchrun --nproc-per-node R tools/pooled_multi_gpu.py`.
Every rank runs det + crop on its block of pages, recognition chunks are built from the global pool and dealt to
ranks (oar_ocr_b200/shard.py: predict_pooled); rank 0 then runs ONE un-sharded predict() on all pages and checks that
the R-rank result is identical (boxes, labels, confidences bit for bit)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oar_ocr_b200 import models, synth
    from oar_ocr_b200.ocr import OAROCRBuilder
    from oar_ocr_b200.shard import GpuStages, predict_pooled, predict_sharded
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")  # host-side exchange of sizes, u8 crops and results
    n_pages, rb = 16, 64
    ocr = (OAROCRBuilder(models.get_blob("det"), models.get_blob("rec"))
           .character_dict_content("\n".join(models.synthetic_dict())).device_id(local)
           .image_batch_size(8).region_batch_size(rb).build())
    pages = [synth.page(500 + i, 960) for i in range(n_pages)]
    got = predict_pooled(GpuStages(ocr), pages, rank, world, region_batch_size=rb)

    def summary(res):
        return [[(r["box"].tobytes(), np.asarray(r["labels"]).tobytes(), float(r["score"])) for r in img] for img in res]

    def of_predict(results):
        return [[(t.bounding_box.points.tobytes(), t.label_indices.tobytes(), float(t.confidence)) for t in r.text_regions]
                for r in results]
    if rank == 0:
        want = of_predict(ocr.predict(pages))
        n = sum(len(x) for x in want)
        print(f"world {world}: pooled == one un-sharded predict(): {summary(got) == want} ({n} regions)", flush=True)
    if world > 1:
        sharded = predict_sharded(lambda blk: of_predict(ocr.predict(blk)), pages, rank, world)
        if rank == 0:
            print(f"world {world}: block-sharded == un-sharded: {sharded == want} (expected False in general)", flush=True)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
