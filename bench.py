#!/usr/bin/env python
"""bench.py -- images/sec through proposal -> NMS -> RoIAlign -> RL-refine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "C4"): COCO scale 800x1200, Res-101 C4 (stride 16 -> 50x75x1024),
12 anchors (45 000 per image), TEST proposal config 6000 -> NMS 0.7 -> 300 rois, RoIAlignAvg
7x7, 16 box-delta actions rewarded against 20 gt boxes (bbox_overlaps IoU), every box takes its
best positive action, RoIAlignAvg re-pool of the refined boxes.  24 images per GPU per step
(the reference's 8-GPU batch), weak scaling: every rank runs its own 24 images and the step ends
with ONE NCCL all-gather of the packed detections || rewards.

One JSON line on rank 0 (see the task contract): value = device-resident throughput, e2e =
through the public API from pinned HOST buffers (H2D of all inputs + D2H of detections and
rewards inside the timed region), roofline = the dominant kernel (k_align8_fwd_walk) timed
live with CUDA events on its own stream, cpu_baseline = the CPU port (oracle/) on a bounded
sample.  --impl reference times that CPU port as the reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "images/sec through proposal->NMS->RoIAlign->RL-refine"
UNIT = "images/s"
# C4 workload
IM_H, IM_W, STRIDE = 800, 1200, 16
FH, FW, C = 50, 75, 1024
SCALES, RATIOS = (4, 8, 16, 32), (0.5, 1, 2)
A = len(SCALES) * len(RATIOS)
PRE, POST, NMS_T = 6000, 300, 0.7
G, ACT_DELTA, POOL = 20, (0.5, 0.25), 7
IMAGES_PER_GPU = 24


def make_inputs(seed, batch):
    from rlobjectdetection_b200 import synthetic as syn
    scores, deltas, im_info = syn.rpn_outputs(seed, batch, A, FH, FW, IM_H, IM_W, IM_H / 600.0)
    g = torch.Generator().manual_seed(seed + 1000)
    feat = torch.randn(batch, C, FH, FW, generator=g)
    gt, _ = syn.gt_boxes(seed + 2000, batch, G, IM_H, IM_W)
    return scores, deltas, im_info, feat, gt


# ------------------------------------------------------------------------------------------
# CPU port of the step (oracle/): cpu_baseline and the --impl reference arm
# ------------------------------------------------------------------------------------------
def cpu_step(orc, inputs, anchors, act):
    scores, deltas, im_info, feat, gt = (t.numpy() for t in inputs)
    rois = orc.proposal_layer(scores, deltas, im_info, anchors, STRIDE, PRE, POST, NMS_T)
    B, N, _ = rois.shape
    pooled = orc.roi_align(feat, rois.reshape(-1, 5), POOL, POOL, 1.0 / STRIDE, pool_mode=orc.POOL_AVG)
    reward, label, _ = orc.action_reward(rois[:, :, 1:5], gt, act, mode=orc.MODE_RCNN)
    # refine: best action per box if its label is +1 (x1y1x2y2, +1 convention)
    best = reward.argmax(axis=2)
    bi, ni = np.meshgrid(np.arange(B), np.arange(N), indexing="ij")
    take = label[bi, ni, best] == 1
    d = act[best]
    b = rois[:, :, 1:5]
    w = b[..., 2] - b[..., 0] + np.float32(1)
    h = b[..., 3] - b[..., 1] + np.float32(1)
    nx, ny = b[..., 0] + d[..., 0] * w, b[..., 1] + d[..., 1] * h
    nw, nh = w + d[..., 2] * w, h + d[..., 3] * h
    moved = np.stack([nx, ny, nx + nw - np.float32(1), ny + nh - np.float32(1)], -1).astype(np.float32)
    refined = rois.copy()
    refined[:, :, 1:5] = np.where(take[..., None], moved, b)
    pooled2 = orc.roi_align(feat, refined.reshape(-1, 5), POOL, POOL, 1.0 / STRIDE, pool_mode=orc.POOL_AVG)
    return rois, reward, refined, pooled, pooled2


def cpu_measure(sample_images, repeats, warmup=1):
    from oracle import oracle as orc
    orc.lib()
    inputs = make_inputs(7, sample_images)
    anchors = orc.generate_anchors(16, RATIOS, SCALES).astype(np.float32)
    act = orc.action_table(list(ACT_DELTA))
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        cpu_step(orc, inputs, anchors, act)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, orc.max_threads()


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path.  Nothing of the path
    compiles for the CPU from /root/reference (RoIAlign is CUDA-only there, its NMS scan is
    host code inside a .cu, roi_pooling.c needs TH) so this is the CPU port in oracle/
    (kind "port"), OpenMP over all host cores, on a bounded sample per step."""
    if rank != 0:
        return
    sample = 4
    times, cores = cpu_measure(sample, args.steps, warmup=max(1, min(args.warmup, 2)))
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(sample, world, "reference-cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images/step of the C4 workload, {len(times)} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(images_per_gpu, world, arm):
    return {
        "workload": "C4: COCO 800x1200 Res-101 C4 full path proposal(45000 anchors, 6000->300, nms 0.7)"
                    " -> RoIAlignAvg 7x7 (1024 ch) -> 16-action IoU reward vs 20 gt -> refine -> re-pool",
        "images_per_gpu": images_per_gpu, "global_batch": images_per_gpu * world,
        "feature": [C, FH, FW], "anchors_per_image": A * FH * FW, "rois_per_image": POST,
        "actions": 4 * len(ACT_DELTA) * 2, "gt_per_image": G, "parallelism": f"image-sharded x{world}",
        "collective": "one all_gather of rois||rewards per step" if world > 1 else "none",
        "l2": "inputs (393 MB/step) and outputs (2.9 GB/step) exceed the 126 MB L2; no flush needed",
        "arm": arm,
    }


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread
    (every 2 ms -- a timed region of tens of milliseconds is too short for `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        import threading
        self.idx, self.samples, self.mask, self.max_mhz = gpu_index, [], 0, None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.idx]) if visible and visible.split(",")[self.idx].isdigit() else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001 -- clocks are reported, never fatal
            self.err = repr(e)

    def start(self):
        self._thread.start()

    def stop(self):
        self._stop.set()
        self._thread.join(timeout=5)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None,
               "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    from rlobjectdetection_b200.model import _backend as be
    from rlobjectdetection_b200.shard import gather_results, pack_results

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = be.lib()
    nb = IMAGES_PER_GPU
    global_batch = nb * world
    first_image = rank * nb

    host = [t.pin_memory() for t in make_inputs(100 + rank, nb)]
    step = DetectRefineStep(STRIDE, SCALES, RATIOS, "TEST", POOL, ACT_DELTA, backward=False)
    from rlobjectdetection_b200.model.utils.config import cfg
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = PRE, POST, NMS_T

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(dev_in):
        out = step(*dev_in)
        packed = pack_results(out["refined"], out["reward"], first_image)
        return out, gather_results(packed, global_batch)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (inputs already in HBM) -----------------------------
    dev_in = [t.to(dev) for t in host]
    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        device_step(dev_in)
    barrier()
    launches0 = lib.rlod_launch_count()
    lib.rlod_profile_enable(1)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: device_step(dev_in), args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    lib.rlod_profile_enable(0)
    launches = lib.rlod_launch_count() - launches0
    prof = be.profile_collect()
    value = global_batch * args.steps / (ms * 1e-3)

    # ---- end to end from pinned host buffers ---------------------------------------------
    result_host = torch.empty(global_batch, POST, 5 + 4 * len(ACT_DELTA) * 2).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = result_host.numel() * result_host.element_size()

    def e2e_step():
        d_in = [t.to(dev, non_blocking=True) for t in host]
        _, gathered = device_step(d_in)
        result_host.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the detections

    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e = timed(e2e_step, e2e_steps, 2)
    e2e_value = global_batch * e2e_steps / (ms_e2e * 1e-3)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ---------------------------------------------------
    R = nb * POST
    alg_bytes = 4 * (nb * C * FH * FW + 5 * R + R * C * POOL * POOL)  # feat once + rois + out once
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("align_fwd_bytes")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": traffic, "kernel": "k_align8_fwd_walk<AVG>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes}
    if "align_fwd" in prof:
        kms, kn = prof["align_fwd"]
        ach = alg_bytes / (kms / kn * 1e-3) / 1e9
        roofline.update(achieved=ach, frac=ach / peak, launches_timed=kn, avg_launch_us=1e3 * kms / kn)
    kernel_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}

    # ---- CPU baseline (port), bounded sample -----------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sample = 4
        times, cores = cpu_measure(sample, 2, warmup=1)
        cpu = {"value": sample / min(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} images of the C4 workload, best of 2 after 1 warm-up (oracle/ CPU port, OpenMP)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(nb, world, "sm_100a kernels"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "kernel_ms_per_step": kernel_ms,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
