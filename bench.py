#!/usr/bin/env python
"""bench.py -- end-to-end OCR images/s (PP-OCRv5-mobile det+rec, 960x960) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on host cores

One step = one OAROCR::predict-equivalent pass (oar_pipeline_run) over one batch of `--batch` synthetic
960x960 pages per GPU (BASELINE.json configs[1]).  Pages shard across ranks with no data-path collective
(SURVEY.md 8e): scaling is weak, `value` = all ranks' images / max-over-ranks device time.

  --workload rec512 : BASELINE.json configs[2] instead -- one oar_rec_run of 512 synthetic 48x320 crops per step
                      (TextRecognitionPredictor semantics: the whole input is one batch); metric = crops/s

  value : inputs resident in HBM before the timed region; timed with CUDA events on the library's launch stream
  e2e   : the same call with HOST (pinned) page buffers: H2D of the pages and D2H of boxes/labels inside
  roofline : dominant kernel's achieved rate, from per-launch CUDA events in profiled steps after the timed region
  cpu_baseline : oracle/ (CPU port of the reference path) timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "end-to-end OCR images/sec (PP-OCRv5 det+rec, 960x960)"
UNIT = "images/s"
SIZE = 960


# The dense contractions run as 3 fp16 tcgen05 MMAs per k-step (hi*hi + hi*lo + lo*hi): the tensor-pipe
# time of a *_tc kernel is 3 x its algorithmic FLOPs / the measured dense 16-bit peak.
TC_PASSES = 3


def is_tc(name: str) -> bool:
    return name.endswith("_tc")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tc=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tc=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, name in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pages(rank: int, batch: int):
    from oar_ocr_b200 import synth
    return [synth.page(rank * batch + i, SIZE) for i in range(batch)]


def cpu_reference_pass(pages, image_bs, region_bs, nets=None, return_results=False):
    """One pass of the reference's CPU path (oracle port: Rust pre/post restated in C++, networks on torch-CPU
    fp32 standing in for ONNX Runtime CPU) over `pages`; returns (seconds, regions)."""
    import torch
    from oar_ocr_b200 import models
    from oracle import pipeline
    from oracle.net import OracleNet
    torch.set_num_threads(os.cpu_count() or 1)
    if nets is None:
        nets = (OracleNet(models.get_blob("det")), OracleNet(models.get_blob("rec")))
    t0 = time.perf_counter()
    res = pipeline.predict(nets[0], nets[1], pages, 18385, image_batch_size=image_bs, region_batch_size=region_bs)
    dt = time.perf_counter() - t0
    if return_results:
        return dt, sum(len(r) for r in res), nets, res
    return dt, sum(len(r) for r in res), nets


def workload_config(args, world):
    return {"workload": f"PP-OCRv5-mobile det+rec end-to-end, batch {args.batch} synthetic {SIZE}x{SIZE} pages per GPU "
                        "(BASELINE.json configs[1])",
            "images_per_step": args.batch * world, "image_batch_size": args.image_batch_size,
            "region_batch_size": args.region_batch_size, "weights": "synthetic planted-signal, seed 42, V=18385",
            "l2": "flushed between steps (256 MiB memset on the launch stream)", "sharding": f"replicas x{world}",
            **({"line_orientation": "on (optional stage; the CPU arms do not run it)"}
               if getattr(args, "line_orientation", False) else {})}


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = args.ref_sample
    pages = make_pages(0, sample)
    nets = None
    for _ in range(max(args.warmup, 0)):
        _, _, nets = cpu_reference_pass(pages[:1], args.image_batch_size, args.region_batch_size, nets)
    total = 0.0
    regions = 0
    for _ in range(args.steps):
        dt, regions, nets = cpu_reference_pass(pages, args.image_batch_size, args.region_batch_size, nets)
        total += dt
    value = sample * args.steps / total
    cores = os.cpu_count() or 1
    desc = (f"{sample} of the workload's {SIZE}x{SIZE} pages per step (seeds 0..{sample - 1}), "
            f"{regions} text regions; oracle port (C++ pre/post + torch-CPU fp32 nets)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def kernel_roofline(ctx, step_fn, extra=None, P=2, workload="pipeline"):
    """Profiles P more steps with CUDA events around every launch; returns (roofline dict of the dominant kernel,
    per-kernel table).  Algorithmic bytes / FLOPs per launch are the ones the library states at each launch site
    (DESIGN.md 5); *_tc kernels are charged TC_PASSES fp16 MMAs per algorithmic FLOP."""
    peaks = load_peaks()
    ctx.profile(True)
    agg = {}
    raw = []
    for _ in range(P):
        step_fn()
        recs = ctx.profile_read()
        raw = recs
        for r in recs:
            a = agg.setdefault(r["name"], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
            a["ms"] += r["ms"]
            a["flops"] += r["flops"]
            a["bytes"] += r["bytes"]
            a["n"] += 1
    ctx.profile(False)
    kernels = []
    tot = sum(a["ms"] for a in agg.values()) or 1.0
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        t = a["ms"] / 1000.0
        gbs = a["bytes"] / t / 1e9 if t > 0 else 0.0
        tfl = a["flops"] / t / 1e12 if t > 0 else 0.0
        t_hbm = a["bytes"] / (peaks["hbm"] * 1e9)
        t_tc = (TC_PASSES * a["flops"] / (peaks["tc"] * 1e12)) if is_tc(name) else 0.0
        kernels.append(dict(name=name, launches_per_step=a["n"] // P, ms_per_step=a["ms"] / P,
                            share=a["ms"] / tot, gbs=gbs, tflops=tfl, bound="hbm" if t_hbm >= t_tc else "tensor",
                            avg_launch_us=1000.0 * a["ms"] / a["n"],
                            roofline_frac=max(t_hbm, t_tc) / t if t > 0 else 0.0))
    top = kernels[0]
    if top["bound"] == "hbm":
        roof = {"bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": top["tflops"], "peak": peaks["tc"] / TC_PASSES, "unit": "TFLOP/s",
                "note": f"algorithmic FLOPs; every k-step is {TC_PASSES} fp16 MMAs (hi/lo operand split), so the "
                        f"peak is the measured dense 16-bit rate / {TC_PASSES}"}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        tk = tj.get("kernels", {}).get(top["name"])
        # a capture may be scoped to the workloads whose launches of that kernel it represents
        if tk and (not tk.get("workloads") or workload in tk["workloads"]):
            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
            # (profiles/traffic.json, keyed by kernel), scaled to the average launch of this run by the capture's
            # traffic / algorithmic ratio
            per_launch = (agg[top["name"]]["bytes"] / agg[top["name"]]["n"]) if agg[top["name"]]["n"] else 0.0
            traffic = {"bytes_per_launch": tk["traffic_over_algorithmic"] * per_launch,
                       "algorithmic_bytes_per_launch": per_launch,
                       "ratio": tk["traffic_over_algorithmic"], "capture": tk["launch"]}
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=top["name"],
                share_of_step=top["share"], avg_launch_us=top["avg_launch_us"], peak_source=peaks["source"],
                measured=f"per-launch CUDA events on the launch stream, {P} profiled steps after the timed region")
    # whole-step roofline: sum over kernels of max(bytes/BW, passes*flops/peak) / sum of kernel times
    roof["step_frac"] = sum(k["roofline_frac"] * k["ms_per_step"] for k in kernels) / \
        (sum(k["ms_per_step"] for k in kernels) or 1.0)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "bench_kernels.json"), "w") as f:
            json.dump(dict(kernels=kernels, **(extra or {})), f, indent=1)
        with open(os.path.join(out_dir, "bench_records.json"), "w") as f:
            json.dump([[r["name"], round(r["ms"], 5), r["flops"], r["bytes"]] for r in raw], f)
    return roof, kernels


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from oar_ocr_b200 import ffi, models
    from oar_ocr_b200.ocr import OAROCRBuilder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    builder = (OAROCRBuilder(models.get_blob("det"), models.get_blob("rec"))
               .character_dict_content("\n".join(models.synthetic_dict())).device_id(local_rank)
               .image_batch_size(args.image_batch_size).region_batch_size(args.region_batch_size))
    if args.line_orientation:  # optional stage, off in the headline configuration
        builder = builder.with_text_line_orientation_classification(models.get_blob("cls"))
    ocr = builder.build()
    if args.engine is not None:
        ocr.det.set_engine(args.engine)
        ocr.rec.set_engine(args.engine)
        if ocr.cls is not None:
            ocr.cls.set_engine(args.engine)
    ctx = ocr.ctx
    B = args.batch
    pages = make_pages(rank, B)
    hs = np.full(B, SIZE, np.int32)
    ws = np.full(B, SIZE, np.int32)
    page_bytes = SIZE * SIZE * 3

    # ---- device-resident inputs (value leg)
    d_base = ctx.device_alloc(B * page_bytes)
    for i, p in enumerate(pages):
        ctx.memcpy_h2d(d_base + i * page_bytes, p)
    dev_ptrs = (C.c_void_p * B)(*[d_base + i * page_bytes for i in range(B)])
    # ---- pinned host inputs (e2e leg)
    pinned = torch.empty((B, SIZE, SIZE, 3), dtype=torch.uint8).pin_memory()
    pinned.numpy()[...] = np.stack(pages)
    host_ptrs = (C.c_void_p * B)(*[pinned.data_ptr() + i * page_bytes for i in range(B)])

    def step(ptrs, on_device):
        ctx.l2_flush()
        return ocr.predict_raw(ptrs, hs, ws, on_device)

    def timed(ptrs, on_device, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            step(ptrs, on_device)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks and rank == 0 else None
        n0, s0 = ffi.launch_count(), ffi.submit_count()
        ctx.timer_start()
        w0 = time.perf_counter()
        regions = 0
        for _ in range(steps):
            b = step(ptrs, on_device)
            regions = int(b.region_off[B])
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - w0) * 1000.0
        launches = (ffi.launch_count() - n0, ffi.submit_count() - s0)
        barrier()
        clocks = sampler.stop() if sampler else None
        return max_over_ranks(ms), max_over_ranks(wall), launches, regions, clocks

    ms, wall, (launches, submits), regions, clocks = timed(dev_ptrs, True, args.steps, args.warmup, True)
    stage = dict(ocr.last_timing)
    value = world * B * args.steps / (ms / 1000.0)
    # (the host-page leg splits the detector batch in two: other graph keys than the device leg's, so it warms up in full)
    e_ms, e_wall, _, _, _ = timed(host_ptrs, False, args.steps, args.warmup)
    e2e_stage = dict(ocr.last_timing)
    e2e_value = world * B * args.steps / (e_ms / 1000.0)

    # ---- per-kernel roofline from profiled steps (after the timed region; events around every launch)
    roof, kernels = (kernel_roofline(ctx, lambda: step(dev_ptrs, True), dict(stage_ms=stage, e2e_stage_ms=e2e_stage))
                     if rank == 0 else (None, []))

    cpu_base = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.ref_sample
        dt0, _, nets = cpu_reference_pass(pages[:1], args.image_batch_size, args.region_batch_size)
        dt, n_reg, _, ref = cpu_reference_pass(pages[:sample], args.image_batch_size, args.region_batch_size, nets,
                                               return_results=True)
        cpu_base = {"value": sample / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"first {sample} of this run's {B} pages, one pass after one warm-up page "
                              f"({n_reg} text regions, {dt:.1f} s); oracle port: C++ restatement of the Rust "
                              "pre/post + torch-CPU fp32 networks (ONNX Runtime is not installable offline)"}
        # the same pages through the CUDA path at the same batch sizes, compared region by region with the oracle run
        # that was just timed: a parity datum at the bench configuration in the bench line itself
        got = ocr.predict(pages[:sample])
        boxes_equal = labels_equal = True
        max_dscore = 0.0
        n_cmp = label_diffs = 0
        for g, w in zip(got, ref):
            if len(g.text_regions) != len(w):
                boxes_equal = labels_equal = False
                continue
            for r, o in zip(g.text_regions, w):
                boxes_equal = boxes_equal and bool(np.array_equal(r.bounding_box.points, o["box"]))
                same = bool(np.array_equal(r.label_indices, o["labels"]))
                labels_equal = labels_equal and same
                label_diffs += 0 if same else 1
                max_dscore = max(max_dscore, abs(float(r.confidence) - float(o["score"])))
                n_cmp += 1
        parity = {"pages": sample, "regions": n_cmp, "boxes_equal": boxes_equal, "labels_equal": labels_equal,
                  "regions_with_label_diffs": label_diffs, "max_dscore": max_dscore,
                  "against": "oracle/ (CPU port), same pages, same image/region batch sizes"}

    if rank == 0:
        engine_dtype = "f32" if args.engine == 0 else ocr_dtype(ocr)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": engine_dtype, "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_stage["h2d_bytes"]),
                    "d2h_bytes_per_step": int(e2e_stage["d2h_bytes"]), "ms_per_step": e_ms / args.steps},
            "gpu_launches": int(launches),
            # host-visible submissions among them: a network batch whose shape has been seen twice is one CUDA graph
            "host_submissions_per_step": submits / args.steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
            "parity_check": parity,
            "regions_per_step_rank0": regions, "wall_ms_per_step": wall / args.steps,
            "stage_ms_last_step": {k: round(v, 3) for k, v in stage.items() if k.startswith("ms_")},
            "top_kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in kk.items()}
                            for kk in kernels[:6]],
        }))
    if world > 1:
        dist.destroy_process_group()


METRIC_REC = "recognizer-only text-line crops/sec (PP-OCRv5 mobile rec: SVTR neck + CTC, 48x320 crops)"
UNIT_REC = "crops/s"


def rec512_config(args, world):
    return {"workload": f"recognizer only: one batch of {args.batch} synthetic 48x320 text-line crops per GPU per step "
                        "(BASELINE.json configs[2]; TextRecognitionPredictor semantics: the whole input is ONE batch)",
            "crops_per_step": args.batch * world, "weights": "synthetic planted-signal, seed 42, V=18385",
            "l2": "flushed between steps (256 MiB memset on the launch stream)", "sharding": f"replicas x{world}"}


def run_rec512(args, rank, local_rank, world):
    """configs[2]: oar_rec_run(_ex) on 512 crops of 48x320.  value: crops resident in HBM; e2e: host crops, one pinned
    staging copy + H2D inside, labels/scores D2H inside."""
    import torch
    import torch.distributed as dist
    from oar_ocr_b200 import ffi, models, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = ffi.Context(local_rank)
    rec = ffi.Model(ctx, models.get_blob("rec"))
    if args.engine is not None:
        rec.set_engine(args.engine)
    B = args.batch
    crops = [synth.crop(rank * B + j, 48, 320) for j in range(B)]
    cb = 48 * 320 * 3
    d_base = ctx.device_alloc(B * cb)
    for i, c in enumerate(crops):
        ctx.memcpy_h2d(d_base + i * cb, np.ascontiguousarray(c))
    dev_ptrs = (C.c_void_p * B)(*[d_base + i * cb for i in range(B)])
    hs = np.full(B, 48, np.int32)
    ws = np.full(B, 320, np.int32)
    last = {}

    def step_dev():
        ctx.l2_flush()
        last["r"] = rec.rec_run_device(dev_ptrs, hs, ws, 18385, 48)

    # e2e leg: the crops lie in pinned host memory (one block, like the pages of the det+rec workload), as the contract
    # asks; oar_rec_run copies page-locked crops from where they lie
    pinned = torch.empty((B, 48, 320, 3), dtype=torch.uint8).pin_memory()
    pinned.numpy()[...] = np.stack(crops)
    host_crops = [pinned.numpy()[i] for i in range(B)]

    def step_host():
        ctx.l2_flush()
        last["r"] = rec.rec_run(host_crops, 18385)

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks and rank == 0 else None
        n0, s0 = ffi.launch_count(), ffi.submit_count()
        ctx.timer_start()
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - w0) * 1000.0
        launches = (ffi.launch_count() - n0, ffi.submit_count() - s0)
        barrier()
        return max_over_ranks(ms), max_over_ranks(wall), launches, (sampler.stop() if sampler else None)

    ms, wall, (launches, submits), clocks = timed(step_dev, args.steps, args.warmup, True)
    value = world * B * args.steps / (ms / 1000.0)
    e_ms, _, _, _ = timed(step_host, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (e_ms / 1000.0)
    T = last["r"]["T"]
    roof, kernels = kernel_roofline(ctx, step_dev, workload="rec512") if rank == 0 else (None, [])
    cpu_base = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        from oracle import pipeline
        from oracle.net import OracleNet
        _t.set_num_threads(os.cpu_count() or 1)
        net = OracleNet(models.get_blob("rec"))
        sample = min(B, args.ref_sample * 16)
        pipeline.rec_forward(net, crops[:4], 18385)
        t0 = time.perf_counter()
        want = pipeline.rec_forward(net, crops[:sample], 18385)
        dt = time.perf_counter() - t0
        cpu_base = {"value": sample / dt, "unit": UNIT_REC, "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"first {sample} of the step's {B} crops as one batch ({dt:.1f} s); oracle port: C++ CRNN "
                              "preprocess + torch-CPU fp32 network + CTC decode"}
        got = rec.rec_run(crops[:sample], 18385)
        parity = {"crops": sample,
                  "labels_equal": all(bool(np.array_equal(a, b)) for a, b in zip(got["labels"], want["labels"])),
                  "max_dscore": float(np.abs(got["scores"] - want["scores"]).max()),
                  "against": "oracle/ (CPU port) on the same crops as one batch"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC_REC, "value": value, "unit": UNIT_REC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": rec512_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT_REC, "h2d_bytes_per_step": B * cb,
                    "d2h_bytes_per_step": B * (T * 8 + 8), "ms_per_step": e_ms / args.steps},
            "gpu_launches": int(launches),
            # host-visible submissions among them: a network batch whose shape has been seen twice is one CUDA graph
            "host_submissions_per_step": submits / args.steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
            "parity_check": parity, "seq_len": T, "wall_ms_per_step": wall / args.steps,
            "top_kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in kk.items()}
                            for kk in kernels[:6]],
        }))
    if world > 1:
        dist.destroy_process_group()


def run_reference_rec512(args, rank, world):
    if rank != 0:
        return
    import torch
    from oar_ocr_b200 import models, synth
    from oracle import pipeline
    from oracle.net import OracleNet
    torch.set_num_threads(os.cpu_count() or 1)
    net = OracleNet(models.get_blob("rec"))
    sample = min(args.batch, args.ref_sample * 16)
    crops = [synth.crop(j, 48, 320) for j in range(sample)]
    for _ in range(max(args.warmup, 0)):
        pipeline.rec_forward(net, crops[:4], 18385)
    total = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        pipeline.rec_forward(net, crops, 18385)
        total += time.perf_counter() - t0
    value = sample * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": METRIC_REC, "value": value, "unit": UNIT_REC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": rec512_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT_REC, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": f"{sample} of the workload's crops per step as one batch; oracle port"},
        "e2e": {"value": value, "unit": UNIT_REC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


METRIC_LAYOUT = "layout-detection pages/sec (PP-DocLayout-L: RT-DETR-L, 1024x1024 pages -> 640x640)"
UNIT_LAYOUT = "pages/s"
LAYOUT_IN = (640, 640)
LAYOUT_PAGE = 1024


def layout_config(args, world):
    return {"workload": f"PP-DocLayout-L layout detection, batch {args.batch} synthetic {LAYOUT_PAGE}x{LAYOUT_PAGE} pages per "
                        f"GPU per step, resized to {LAYOUT_IN[0]}x{LAYOUT_IN[1]} (BASELINE.json configs[4]: batch 64 over 8 GPUs "
                        "= 8 pages per GPU)",
            "pages_per_step": args.batch * world, "weights": "synthetic He-normal, seed 42, 23 classes",
            "l2": "flushed between steps (256 MiB memset on the launch stream)", "sharding": f"replicas x{world}"}


def layout_pages(rank, batch):
    from oar_ocr_b200 import synth
    return [synth.page(5000 + rank * batch + i, LAYOUT_PAGE) for i in range(batch)]


def run_layout(args, rank, local_rank, world):
    """configs[4]: pages -> CatmullRom resize -> RT-DETR-L -> rows (oar_layout_rows).  value: pages resident in HBM;
    e2e: host pages (pinned staging + H2D inside), rows D2H inside both."""
    import torch
    import torch.distributed as dist
    from oar_ocr_b200 import ffi, models
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = ffi.Context(local_rank)
    w = models.layout_weights(42)
    shapes = [(LAYOUT_IN[0] // s, LAYOUT_IN[1] // s) for s in (8, 16, 32)]
    enc = ffi.Model(ctx, models.build_layout_encoder(w, seed=42, shapes_hw=shapes))
    head = ffi.Model(ctx, models.build_layout_head(w))
    if args.engine is not None:
        enc.set_engine(args.engine)
        head.set_engine(args.engine)
    B = args.batch
    pages = layout_pages(rank, B)
    pb = LAYOUT_PAGE * LAYOUT_PAGE * 3
    d_base = ctx.device_alloc(B * pb)
    for i, p in enumerate(pages):
        ctx.memcpy_h2d(d_base + i * pb, p)
    dev_ptrs = (C.c_void_p * B)(*[d_base + i * pb for i in range(B)])
    hs = np.full(B, LAYOUT_PAGE, np.int32)
    ws = np.full(B, LAYOUT_PAGE, np.int32)
    last = {}

    def step_dev():
        ctx.l2_flush()
        last["rows"] = ffi.layout_rows(enc, head, None, LAYOUT_IN, device_table=(dev_ptrs, hs, ws))

    # e2e leg: pages in pinned host memory, as the contract asks (page-locked pages are copied from where they lie)
    pinned = torch.empty((B, LAYOUT_PAGE, LAYOUT_PAGE, 3), dtype=torch.uint8).pin_memory()
    pinned.numpy()[...] = np.stack(pages)
    host_pages = [pinned.numpy()[i] for i in range(B)]

    def step_host():
        ctx.l2_flush()
        last["rows"] = ffi.layout_rows(enc, head, host_pages, LAYOUT_IN)

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks and rank == 0 else None
        n0, s0 = ffi.launch_count(), ffi.submit_count()
        ctx.timer_start()
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - w0) * 1000.0
        launches = (ffi.launch_count() - n0, ffi.submit_count() - s0)
        barrier()
        return max_over_ranks(ms), max_over_ranks(wall), launches, (sampler.stop() if sampler else None)

    ms, wall, (launches, submits), clocks = timed(step_dev, args.steps, args.warmup, True)
    value = world * B * args.steps / (ms / 1000.0)
    e_ms, _, _, _ = timed(step_host, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (e_ms / 1000.0)
    roof, kernels = kernel_roofline(ctx, step_dev, workload="layout") if rank == 0 else (None, [])
    cpu_base = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        from oracle import cpu
        from oracle.rtdetr import RTDetrL
        _t.set_num_threads(os.cpu_count() or 1)
        net = RTDetrL(42)
        sample = min(B, max(1, args.ref_sample // 2))
        sizes = [(float(LAYOUT_PAGE), float(LAYOUT_PAGE))] * sample
        x1, _ = cpu.layout_preprocess(pages[:1], LAYOUT_IN)
        net.rows(x1, sizes[:1])
        t0 = time.perf_counter()
        x, _ = cpu.layout_preprocess(pages[:sample], LAYOUT_IN)
        want = net.rows(x, sizes)
        dt = time.perf_counter() - t0
        cpu_base = {"value": sample / dt, "unit": UNIT_LAYOUT, "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": f"first {sample} of the step's {B} pages ({dt:.1f} s); oracle port: C++ CatmullRom resize + "
                              "torch-CPU fp32 RT-DETR-L (oracle/rtdetr.py)"}
        got = last["rows"][:sample]
        k = 50  # the 50 best rows of every page: same classes, scores and corners close
        parity = {"pages": sample,
                  "top_score_max_diff": float(np.abs(got[:, 0, 1] - want[:, 0, 1]).max()),
                  "top50_score_max_diff": float(np.abs(np.sort(got[:, :k, 1], 1) - np.sort(want[:, :k, 1], 1)).max()),
                  "against": "oracle/rtdetr.py (CPU port) on the same pages"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC_LAYOUT, "value": value, "unit": UNIT_LAYOUT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": layout_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT_LAYOUT, "h2d_bytes_per_step": B * pb,
                    "d2h_bytes_per_step": B * 300 * 6 * 4, "ms_per_step": e_ms / args.steps},
            "gpu_launches": int(launches),
            # host-visible submissions among them: a network batch whose shape has been seen twice is one CUDA graph
            "host_submissions_per_step": submits / args.steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
            "parity_check": parity, "wall_ms_per_step": wall / args.steps,
            "top_kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in kk.items()}
                            for kk in kernels[:8]],
        }))
    if world > 1:
        dist.destroy_process_group()


def run_reference_layout(args, rank, world):
    if rank != 0:
        return
    import torch
    from oracle import cpu
    from oracle.rtdetr import RTDetrL
    torch.set_num_threads(os.cpu_count() or 1)
    net = RTDetrL(42)
    sample = min(args.batch, max(1, args.ref_sample // 2))
    pages = layout_pages(0, sample)
    sizes = [(float(LAYOUT_PAGE), float(LAYOUT_PAGE))] * sample
    for _ in range(max(args.warmup, 0)):
        x, _ = cpu.layout_preprocess(pages[:1], LAYOUT_IN)
        net.rows(x, sizes[:1])
    total = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        x, _ = cpu.layout_preprocess(pages, LAYOUT_IN)
        net.rows(x, sizes)
        total += time.perf_counter() - t0
    value = sample * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": METRIC_LAYOUT, "value": value, "unit": UNIT_LAYOUT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": layout_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT_LAYOUT, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": f"{sample} of the workload's pages per step; oracle port"},
        "e2e": {"value": value, "unit": UNIT_LAYOUT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ocr_dtype(ocr) -> str:
    # arithmetic type of the path: fp32 tensors end to end; dense contractions run as 3 fp16 tcgen05 MMAs per k-step
    # (hi/lo operand split) into fp32 TMEM accumulators, which reproduces fp32 GEMM results to ~1e-6
    return "f32"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "rec512", "layout"],
                    help="pipeline = BASELINE.json configs[1] (the headline metric); rec512 = configs[2], recognizer only; "
                         "layout = configs[4], PP-DocLayout-L")
    ap.add_argument("--batch", type=int, default=None, help="pages (pipeline: 32) or crops (rec512: 512) per GPU per step")
    ap.add_argument("--image-batch-size", type=int, default=32)
    ap.add_argument("--region-batch-size", type=int, default=256)
    ap.add_argument("--engine", type=int, default=None, help="0 = fp32 SIMT engine, 1 = tcgen05 one kernel per layer, 2 = tcgen05 fused persistent blocks (default)")
    ap.add_argument("--ref-sample", type=int, default=4, help="pages per CPU-reference pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--line-orientation", action="store_true",
                    help="also run the optional text-line orientation classifier (not the headline configuration)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries ONE JSON line.  Libraries print there too (NCCL's version banner comes out of a plain printf at the
    # first communicator): file descriptor 1 is pointed at stderr for the whole run, and Python's sys.stdout -- which
    # only the JSON line goes through -- keeps the real one.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.batch is None:
        args.batch = {"rec512": 512, "layout": 8}.get(args.workload, 32)
    if args.workload == "layout":
        if args.impl == "reference":
            run_reference_layout(args, rank, world)
        else:
            run_layout(args, rank, local_rank, world)
    elif args.workload == "rec512":
        if args.impl == "reference":
            run_reference_rec512(args, rank, world)
        else:
            run_rec512(args, rank, local_rank, world)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
