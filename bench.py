#!/usr/bin/env python
"""bench.py -- images/sec through proposal -> NMS -> RoIAlign -> RL-refine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "C4"): COCO scale 800x1200, Res-101 C4 (stride 16 -> 50x75x1024),
12 anchors (45 000 per image), TEST proposal config 6000 -> NMS 0.7 -> 300 rois, RoIAlignAvg
7x7, 16 box-delta actions rewarded against 20 gt boxes (bbox_overlaps IoU), every box takes its
best positive action, RoIAlignAvg re-pool of the refined boxes.  The step is the reference's
8-GPU batch: 24 images (README.md:42 "3 per GPU, totally 24").  --gpus N SHARDS that batch
(shard_bounds: 12 / 6 / 3 images per GPU at N = 2 / 4 / 8, RCNN_bases/trainval_net.py:292-293):
strong scaling, every rank runs its images and the step ends with ONE NCCL all-gather of the
packed detections || rewards.  N = 1 launches eagerly (two streams, the light kernels one step
ahead); N > 1 replays the same launches from a CUDA graph (a 3-image shard needs ~0.2 ms of GPU
time, less than the host needs to enqueue ~25 launches).  The 24-images-per-GPU (weak) number of
round 1 is kept as the secondary field `weak_scaling`.

One JSON line on rank 0 (see the task contract): value = device-resident throughput, e2e =
through the public API from pinned HOST buffers (H2D of all inputs + D2H of detections and
rewards inside the timed region), e2e_features_resident = the same with the feature map already
on the device (where the reference's backbone leaves it, faster_rcnn.py:47), roofline = the
dominant kernel (k_align8_fwd_walk2) timed live with CUDA events on its own stream, cpu_baseline /
--impl reference = the reference's own Python for the stages it has on the CPU + the CPU port
(oracle/) for its CUDA-only stages (baseline/ref_arm.py), ops = every other BASELINE config next
to the reference's legacy CUDA kernels, verified = the timed step's outputs against the oracle.
"""
import argparse
import gc
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
IMAGES = 24  # BASELINE config 4's global batch (the reference's 8-GPU batch, README.md:42)


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
    refined, _ = orc.refine_best_action(rois, reward, label, act)
    pooled2 = orc.roi_align(feat, refined.reshape(-1, 5), POOL, POOL, 1.0 / STRIDE, pool_mode=orc.POOL_AVG)
    return rois, reward, refined, pooled, pooled2


def cpu_measure(sample_images, repeats, warmup=1, budget_s=None):
    """Times the CPU arm: the reference's own Python where the reference has a CPU implementation
    (baseline/ref_arm.py over baseline/_ref/lib, vendored at build time) + the OpenMP port for its
    CUDA-only stages -> kind "reference+port"; the port alone ("port") when the vendored files are
    missing.  Returns (times, cores, kind)."""
    from oracle import oracle as orc
    from baseline import ref_arm
    orc.lib()
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every core this process may run on
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    orc.set_threads(ncores)
    torch.set_num_threads(ncores)
    inputs = make_inputs(100, sample_images)
    anchors = orc.generate_anchors(16, RATIOS, SCALES).astype(np.float32)
    act = orc.action_table(list(ACT_DELTA))
    use_ref = ref_arm.available()

    def one(inp):
        if use_ref:
            ref_arm.step(orc, inp, STRIDE, SCALES, RATIOS, PRE, POST, NMS_T, POOL, ACT_DELTA)
        else:
            cpu_step(orc, inp, anchors, act)
    # warm-up (imports, thread pools, page faults) on a 2-image slice: a whole 24-image step of the reference's
    # Python costs ~10 s (its per-image loop runs a full 6000-box NMS per image on one core)
    small = [t[:2].contiguous() for t in inputs]
    for _ in range(warmup):
        one(small)
    times, t_start = [], time.perf_counter()
    for i in range(repeats):
        t0 = time.perf_counter()
        one(inputs)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    return times, orc.max_threads(), ("reference+port" if use_ref else "port")


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the host's cores, the SAME
    24-image step as our arm.  The reference's own Python runs unmodified for the stages it has on
    the CPU (_ProposalLayer with bbox_transform_inv / clip_boxes / torch.sort, bbox_overlaps for the
    rewards, the Action table); RoIAlign is CUDA-only in the reference and its NMS is a cffi
    extension that cannot be built here, so those two stages are the OpenMP port in oracle/
    (kind "reference+port")."""
    if rank != 0:
        return
    sample = IMAGES
    warm = max(1, min(args.warmup, 2))
    times, cores, kind = cpu_measure(sample, args.steps, warmup=warm, budget_s=300.0)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(sample, 1, "reference-cpu", sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} images/step of the C4 workload (the whole step), {len(times)} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(images_per_gpu, world, arm, global_batch=None):
    return {
        "workload": "C4: COCO 800x1200 Res-101 C4 full path proposal(45000 anchors, 6000->300, nms 0.7)"
                    " -> RoIAlignAvg 7x7 (1024 ch) -> 16-action IoU reward vs 20 gt -> refine -> re-pool",
        "images_per_gpu": images_per_gpu, "global_batch": global_batch if global_batch is not None else images_per_gpu * world,
        "feature": [C, FH, FW], "anchors_per_image": A * FH * FW, "rois_per_image": POST,
        "actions": 4 * len(ACT_DELTA) * 2, "gt_per_image": G, "parallelism": f"image-sharded x{world}",
        "collective": ("one all_gather of rois||rewards per step, asynchronous: it overlaps the next step's kernels "
                       "(shard.PipelinedGather); all of them complete inside the timed region") if world > 1 else "none",
        "l2": "inputs (393 MB/step) and outputs (2.9 GB/step) exceed the 126 MB L2; no flush needed",
        "streams": "2 per GPU: per-image kernels (proposal select/sort/NMS, reward, refine) on a light stream that runs "
                   "one step ahead under the RoIAlign kernels of the caller's stream; every step does all of its work",
        "arm": arm,
    }


REPOOL_NOTE = {
    "merged": "the proposals and the refined boxes (which come from the IoU rewards, not from the pooled features) are "
              "pooled by ONE RoIAlignAvg call on their concatenation: every feature plane is staged once per step and "
              "the rois are planned once; outputs identical to two calls (other_repool_form times those)",
    "separate": "two RoIAlignAvg calls per step (proposals, refined boxes)",
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
        self._armed = threading.Event()   # samples count only while the timed region runs
        self._ready = threading.Event()   # NVML is initialised (nvmlInit can take > 50 ms)
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
            # the first query of each kind is the slow one (tens of ms): spend it before the timed region
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self._ready.set()
            while not self._stop.is_set():
                # NVML is queried only while armed (the timed region).  (While a cudaMalloc still fell into
                # the first timed step, concurrent NVML queries stretched that stall from milliseconds to tens
                # of milliseconds -- the queries contend for driver locks; without the cudaMalloc 11 runs in a
                # row were within 0.1 %.)
                if self._armed.is_set():
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    time.sleep(0.0005)
                else:
                    self._stop.wait(0.002)
        except Exception as e:  # noqa: BLE001 -- clocks are reported, never fatal
            self.err = repr(e)
            self._ready.set()

    def start(self):
        self._thread.start()
        self._ready.wait(timeout=10)

    def arm(self, on=True):
        (self._armed.set if on else self._armed.clear)()

    def stop(self):
        late = False
        if not self.samples and self._thread.is_alive() and not self.err:
            # the timed region ended before the first NVML query returned: one sample right after it
            late = True
            self._armed.set()
            t0 = time.time()
            while not self.samples and time.time() - t0 < 0.5:
                time.sleep(0.002)
            self._armed.clear()
        self._late = late
        self._stop.set()
        self._thread.join(timeout=5)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None,
               "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}
        if self.err:
            out["error"] = self.err
        if getattr(self, "_late", False):
            out["note"] = "sampled right after the timed region (it ended before the first NVML query returned)"
        return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa(gpu_index):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE the pinned host
    buffers are allocated, so that first touch puts them on the GPU's NUMA node: eight ranks
    uploading 394 MB each share the host's memory and PCIe root complexes (round 1: 120 GB/s
    aggregate at N = 4).  Returns (cpus now allowed, cpus allowed before)."""
    before = sorted(os.sched_getaffinity(0))
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= set(before)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001 -- binding is an optimisation, never fatal
        pass
    return sorted(os.sched_getaffinity(0)), before


def verify_step(dev_in, host, timed_packed, nv=2):
    """The oracle check of what was just timed: the first `nv` images of the bench inputs through a
    fresh step (all outputs) against the CPU oracle -- sort order / NMS / rewards / refined boxes
    bit-exact on the GPU's decoded boxes, decoded boxes within 3e-6, pooled features within 1e-5 --
    and the timed step's own packed rows of those images equal to it bit for bit (the path is
    independent per image)."""
    from oracle import oracle as orc
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    from rlobjectdetection_b200.model import _backend as be
    nv = min(nv, dev_in[0].size(0))
    sub = [t[:nv].contiguous() for t in dev_in]
    step = DetectRefineStep(STRIDE, SCALES, RATIOS, "TEST", POOL, ACT_DELTA, backward=False,
                            outputs=("rois", "reward", "label", "refined", "packed"), first_image=0)
    out = step(*sub)
    torch.cuda.synchronize()
    scores, deltas, im_info, feat, gt = (t[:nv].numpy() for t in host)
    anchors = step.proposal._anchors.cpu().numpy()
    _, order, props, nkeep = be.proposal_forward(sub[0], sub[1], sub[2], step.proposal._anchors, STRIDE, PRE, POST, NMS_T,
                                                 return_taps=True)
    o_rois, o_order, o_props, _, _ = orc.proposal_layer(scores, deltas, im_info, anchors, STRIDE, PRE, POST, NMS_T,
                                                        return_taps=True)
    res = {"images": nv, "sort_order_bit_exact": bool(np.array_equal(order.cpu().numpy(), o_order))}
    props = props.cpu().numpy()
    res["decoded_boxes_max_abs_err"] = float(np.abs(props - o_props).max())
    g_rois = orc.proposal_layer(scores, deltas, im_info, anchors, STRIDE, PRE, POST, NMS_T, boxes_override=props)
    rois = out["rois"].cpu().numpy()
    res["rois_bit_exact_on_gpu_boxes"] = bool(np.array_equal(rois, g_rois))
    res["rois_max_abs_err_vs_cpu_chain"] = float(np.abs(rois - o_rois).max())
    act = step.action.actDeltas
    rr, rl, _ = orc.action_reward(rois[:, :, 1:5], gt, act, mode=orc.MODE_RCNN)
    refined, _ = orc.refine_best_action(rois, rr, rl, act)
    res["reward_bit_exact"] = bool(np.array_equal(out["reward"].cpu().numpy(), rr))
    res["refined_bit_exact"] = bool(np.array_equal(out["refined"].cpu().numpy(), refined))
    worst = 0.0
    for name, boxes in (("pooled", rois), ("pooled_refined", refined)):
        ref = orc.roi_align(feat, boxes.reshape(-1, 5), POOL, POOL, 1.0 / STRIDE, pool_mode=orc.POOL_AVG)
        got = out[name].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max() / max(float(np.abs(ref).max()), 1e-30)))
    res["pooled_max_err_rel_to_max"] = worst
    res["timed_output_matches"] = bool(torch.equal(timed_packed[:nv, :, 1:], out["packed"][:, :, 1:]))
    res["ok"] = bool(res["sort_order_bit_exact"] and res["rois_bit_exact_on_gpu_boxes"] and res["reward_bit_exact"]
                     and res["refined_bit_exact"] and worst <= 1e-5 and res["decoded_boxes_max_abs_err"] <= 2e-3
                     and res["timed_output_matches"])
    return res


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    from rlobjectdetection_b200.model import _backend as be
    from rlobjectdetection_b200.model.utils.config import cfg
    from rlobjectdetection_b200.shard import gather_results, shard_bounds

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpus_bound, cpus_all = bind_to_gpu_numa(local_rank)
    lib = be.lib()
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = PRE, POST, NMS_T
    # strong scaling: BASELINE config 4's batch of 24 images, contiguous shards (12 / 6 / 3 per GPU)
    lo, hi = shard_bounds(IMAGES, rank, world)
    nb = hi - lo
    graphed = world > 1 and not args.eager
    full = make_inputs(100, IMAGES)  # every rank generates the same global batch and keeps its shard
    host = [t[lo:hi].contiguous().pin_memory() for t in full]
    n_act = 4 * len(ACT_DELTA) * 2

    def new_step(first_image, repool=None):
        return DetectRefineStep(STRIDE, SCALES, RATIOS, "TEST", POOL, ACT_DELTA, backward=False,
                                outputs=("packed",), first_image=first_image,  # the rows of the gathered result
                                repool=repool or args.repool)
    step = new_step(lo)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eager_step(st, dev_in, global_batch, ready=True, nxt=None, nxt_ready=None):
        # ready=True: the inputs are resident and stable; an Event: ready once it fired.  nxt: the next
        # step's inputs, whose per-image kernels are enqueued between this step's two RoIAlign launches
        s_, d_, i_, f_, g_ = dev_in
        out = st(s_, d_, i_, f_, g_, inputs_ready=ready,
                 next_inputs=None if nxt is None else (nxt[0], nxt[1], nxt[2], nxt[4]), next_ready=nxt_ready)
        return out, gather_results(out["packed"], global_batch)

    step_marks, host_ms = [], []

    def timed(fn, steps, warmup, marks=None, before_timed=None):
        for _ in range(warmup):
            fn()
        gc.collect()
        gc.disable()  # like timeit: a cyclic collection in the launching thread is a multi-millisecond stall
        barrier()
        if before_timed is not None:
            before_timed()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host_t = [time.perf_counter()]
        for k_ in range(steps):
            fn()
            if marks is not None:  # one event per step on the caller's stream: the spread of the step time
                m = torch.cuda.Event(enable_timing=True)
                m.record()
                marks.append(m)
                host_t.append(time.perf_counter())
        e1.record()
        barrier()
        gc.enable()
        ms = e0.elapsed_time(e1)
        if marks:
            ts = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
            marks[:] = ts
            host_ms[:] = [1e3 * (b - a) for a, b in zip(host_t, host_t[1:])]
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput (inputs already in HBM) -----------------------------
    dev_in = [t.to(dev) for t in host]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # spare small-pool segments for the caller's stream too (see DetectRefineStep._light_stream): no cudaMalloc
    # may happen inside the timed region
    spare = [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _ in range(32)]
    del spare
    n_warm = max(args.warmup, 3)
    count0, mstat, last = [0], {}, {}

    def before_timed():
        count0[0] = lib.rlod_launch_count()
        mstat["alloc0"] = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        sampler.arm(True)

    if not graphed:
        # Inside the timed region only the roofline kernel's launches are bracketed by events (two event records
        # per launch cost 1-2 us of stream time each: 3 % of the step when all ~20 launches are bracketed); the
        # other kernels' times come from a few extra, untimed steps afterwards.
        lib.rlod_profile_only(be.KERNELS.index("align_fwd"))
        lib.rlod_profile_enable(1)  # before the warm-up: the first event pairs are created outside the timed region

        def fn():
            last["out"], last["gathered"] = eager_step(step, dev_in, IMAGES, True, dev_in, True)
        for _ in range(n_warm - 1):
            fn()
        torch.cuda.synchronize()
        be.profile_collect()        # drop the warm-up launches (host-side work: milliseconds)
        # the last warm-up step runs AFTER the host-side housekeeping, so that the GPU (and the link to it) has
        # been idle for one synchronise only when the timed region starts
        ms = timed(fn, args.steps, 1, step_marks, before_timed=before_timed)
        launches_per_step = None
    else:
        # the whole step as ONE CUDA graph per rank, pipelined over its own inputs (hotpath.GraphedStep): the
        # graph pools the rois its previous replay prepared while its light branch prepares the next step's
        gs = step.capture(*dev_in, next_inputs=(dev_in[0], dev_in[1], dev_in[2], dev_in[4]))
        gs.prime()
        c0 = lib.rlod_launch_count()
        eager_step(new_step(lo), dev_in, IMAGES)   # what one step launches when it is not replayed from the graph
        launches_per_step = lib.rlod_launch_count() - c0

        # the all-gather of step i runs while step i+1 computes (shard.PipelinedGather: staging copy + async
        # collective); the timed region ends with a barrier, so every collective of it completes inside it
        from rlobjectdetection_b200.shard import PipelinedGather
        pg = PipelinedGather(gs.replay()["packed"], IMAGES)
        torch.cuda.synchronize()

        def fn():
            last["out"] = gs.replay()
            pg.submit(last["out"]["packed"])
        ms = timed(fn, args.steps, n_warm, step_marks, before_timed=before_timed)
        last["gathered"] = pg.result()
        pg.drain()
    mallocs_in_region = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mstat["alloc0"]
    sampler.arm(False)
    clocks = sampler.stop() if rank == 0 else None
    lib.rlod_profile_enable(0)
    launches = (lib.rlod_launch_count() - count0[0]) if not graphed else launches_per_step * args.steps
    prof = be.profile_collect()
    prof_steps = args.steps + 1  # the per-kernel event times also cover that last warm-up step
    value = IMAGES * args.steps / (ms * 1e-3)
    timed_gathered = last["gathered"].clone()
    timed_packed = last["out"]["packed"].clone()
    # per-kernel times: a few untimed EAGER steps with every launch bracketed (graphed: incl. the roofline kernel)
    lib.rlod_profile_only(-1)
    lib.rlod_profile_enable(1)
    extra_steps = 5
    for _ in range(extra_steps):
        eager_step(step, dev_in, IMAGES, True, dev_in, True)
    torch.cuda.synchronize()
    lib.rlod_profile_enable(0)
    prof_all = be.profile_collect()
    if graphed:
        prof, prof_steps = {"align_fwd": prof_all["align_fwd"]} if "align_fwd" in prof_all else {}, extra_steps

    # ---- end to end from pinned host buffers ---------------------------------------------
    result_host = torch.empty(IMAGES, POST, 5 + n_act).pin_memory()
    d2h = result_host.numel() * result_host.element_size()
    # Every step copies ITS inputs host -> device and ITS result device -> host inside the timed
    # region.  The copies run on their own stream into double-buffered device tensors, so step
    # i+1's upload overlaps step i's kernels (what a serving loop does); the caller-visible
    # result of step i is complete (synchronised) before step i+1's kernels are enqueued.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    def e2e_leg(upload_idx):
        """upload_idx: which of (scores, deltas, im_info, feat, gt) cross PCIe every step; the others stay
        resident (the feature map is produced on the GPU in the reference, faster_rcnn.py:47)."""
        bufs = [[torch.empty_like(t, device=dev) if i in upload_idx else dev_in[i] for i, t in enumerate(host)]
                for _ in range(2)]
        up_done = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0, "primed": False}
        e2e_stepper = new_step(lo)

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[slot])  # the step that last read this buffer is done
                for i in upload_idx:
                    bufs[slot][i].copy_(host[i], non_blocking=True)
                up_done[slot].record(copy_stream)

        def e2e_step():
            i = state["i"]
            slot = i & 1
            if not state["primed"]:
                free[0].record(main_stream), free[1].record(main_stream)
                upload(slot)
                state["primed"] = True
            upload(slot ^ 1)                     # next step's inputs travel while this step computes
            main_stream.wait_event(up_done[slot])
            _, gathered = eager_step(e2e_stepper, bufs[slot], IMAGES, ready=up_done[slot], nxt=bufs[slot ^ 1],
                                     nxt_ready=up_done[slot ^ 1])
            free[slot].record(main_stream)
            result_host.copy_(gathered, non_blocking=True)
            main_stream.synchronize()            # the caller reads the detections
            state["i"] = i + 1

        n = max(2, min(args.steps, 20))
        ms_ = timed(e2e_step, n, 2)
        h2d_rank = sum(host[i].numel() * host[i].element_size() for i in upload_idx)
        h2d_all = sum(full[i].numel() * full[i].element_size() for i in upload_idx)
        return {"value": IMAGES * n / (ms_ * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_ / n, "steps": n, "h2d_bytes_per_step_per_gpu": h2d_rank,
                "aggregate_h2d_gbs": round(h2d_all / (ms_ / n * 1e-3) / 1e9, 2)}

    e2e = e2e_leg((0, 1, 2, 3, 4))
    e2e_resident = e2e_leg((0, 1, 2, 4))

    # ---- the round-1 weak-scaling number (24 images on EVERY GPU), secondary ----------------
    weak = None
    if world > 1:
        w_host = make_inputs(100 + rank, IMAGES)
        w_in = [t.to(dev) for t in w_host]
        w_step = new_step(rank * IMAGES)
        w_steps = max(2, min(args.steps, 10))
        ms_w = timed(lambda: eager_step(w_step, w_in, IMAGES * world, True, w_in, True), w_steps, 3)
        weak = {"images_per_gpu": IMAGES, "global_batch": IMAGES * world, "value": IMAGES * world * w_steps / (ms_w * 1e-3),
                "unit": UNIT, "ms_per_step": ms_w / w_steps, "steps": w_steps, "launch": "eager"}
        del w_in, w_step
        torch.cuda.empty_cache()

    # ---- the other form of the step (two RoIAlign launches instead of one), short leg, N = 1 only ----
    other_form = None
    if world == 1 and not graphed:
        o_name = "separate" if args.repool == "merged" else "merged"
        o_step = new_step(lo, o_name)

        def o_fn():
            eager_step(o_step, dev_in, IMAGES, True, dev_in, True)
        o_steps = max(10, args.steps // 2)
        ms_o = timed(o_fn, o_steps, 3)
        other_form = {"repool": o_name, "value": IMAGES * o_steps / (ms_o * 1e-3), "unit": UNIT,
                      "ms_per_step": ms_o / o_steps, "steps": o_steps}
        del o_step
    # ---- N > 1: the gathered result of the sharded batch == one rank running the whole batch ----
    multi_ok = None
    if world > 1 and rank == 0:
        whole = [t.to(dev) for t in full]
        ref = new_step(0)(*whole)
        torch.cuda.synchronize()
        multi_ok = bool(torch.equal(timed_gathered, ref["packed"]))
        del whole, ref
    if rank != 0:
        return
    # ---- roofline of the dominant kernel ---------------------------------------------------
    R = nb * POST * (2 if args.repool == "merged" else 1)  # rois per launch: proposals + refined boxes when merged
    alg_bytes = 4 * (nb * C * FH * FW + 5 * R + R * C * POOL * POOL)  # feat once + rois + out once
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    if world == 1:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(
                "align_fwd_merged_bytes" if args.repool == "merged" else "align_fwd_bytes")
        except (OSError, ValueError):
            pass
    roofline = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": traffic, "kernel": "k_align8_fwd_walk2<AVG>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "timed": "CUDA events around every launch of the kernel inside the timed region" if not graphed else
                         f"CUDA events around the kernel's launches in {extra_steps} untimed eager steps after the "
                         "graph-replayed timed region (events inside a graph cannot be read back)"}
    if "align_fwd" in prof:
        kms, kn = prof["align_fwd"]
        ach = alg_bytes / (kms / kn * 1e-3) / 1e9
        roofline.update(achieved=ach, frac=ach / peak, launches_timed=kn, avg_launch_us=1e3 * kms / kn)
    kernel_ms = {k: round(v[0] / extra_steps, 4) for k, v in prof_all.items()}
    if not graphed:
        kernel_ms["align_fwd"] = round(prof["align_fwd"][0] / prof_steps, 4) if "align_fwd" in prof else None

    # ---- what was timed, against the oracle -------------------------------------------------
    verified = None
    if not args.no_verify:
        verified = verify_step(dev_in, host, timed_packed)
        if multi_ok is not None:
            verified["gathered_equals_single_rank"] = multi_ok
            verified["ok"] = bool(verified["ok"] and multi_ok)

    # ---- every other BASELINE config, next to the reference's legacy CUDA kernels ------------
    ops = None
    if world == 1 and not args.no_ops:
        dev_in = None
        torch.cuda.empty_cache()
        ops = ops_summary(measure_ops(iters=20, legacy_iters=4))

    # ---- CPU baseline, bounded sample ---------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, cpus_all)  # the CPU arm uses every core, not only the GPU's NUMA node
        times, cores, kind = cpu_measure(IMAGES, 2, warmup=1, budget_s=12.0)
        cpu = {"value": IMAGES / min(times), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{IMAGES} images (one whole step of the C4 workload), best of {len(times)} after 1 warm-up; "
                         "reference Python for proposal / decode / clip / sort / rewards, OpenMP port for NMS and RoIAlign"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(nb, world, "sm_100a kernels", IMAGES),
        "clocks": clocks,
        "e2e": e2e,
        "e2e_features_resident": e2e_resident,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "verified": verified,
        "weak_scaling": weak,
        "other_repool_form": other_form,
        "ops": ops,
        "kernel_ms_per_step": kernel_ms,
        # one event per step on the caller's stream.  The first step after the synchronise has nothing queued
        # ahead of it, so every host-side hiccup of its launches is GPU idle time (2-70 ms seen on busy hosts);
        # from the second step on the host runs ahead and the step time is the GPU's
        "step_ms_spread": {"min": round(min(step_marks), 4), "median": round(statistics.median(step_marks), 4),
                           "max": round(max(step_marks), 4), "first5": [round(v, 3) for v in step_marks[:5]],
                           "host_enqueue_ms_median": round(statistics.median(host_ms), 4),
                           "cudaMalloc_calls_in_timed_region": int(mallocs_in_region)} if step_marks else None,
        "host": {"cpus_bound_to_gpu_numa": len(cpus_bound), "cpus_total": len(cpus_all)},
        "kernel_ms_note": "CUDA events around every launch on its own stream; eager: align_fwd from the timed region, the "
                          "others from 5 untimed steps after it with every launch bracketed; graphed: all from those 5 eager "
                          "steps.  For the light stream's kernels this is launch-to-finish time, queueing behind the "
                          "RoIAlign launches for a free SM included",
    }
    line["config"]["repool"] = REPOOL_NOTE[args.repool]
    line["config"]["launch"] = ("one CUDA graph per step and rank (hotpath.GraphedStep, pipelined) + one NCCL all_gather"
                                if graphed else "eager launches on two streams")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# --ops: every op of the path next to the reference's own legacy CUDA kernel (oracle/_ref/
# libref_legacy.so = the reference's .cu files compiled unchanged for sm_100a) on the same GPU.
# This is the baseline leg north_star asks for ("for ops that are CUDA-only in the reference,
# the reference's legacy CUDA kernel on the same GPU is reported alongside"); like
# cpu_baseline it is the only place outside tests/ that executes anything under oracle/.
# ------------------------------------------------------------------------------------------
def measure_ops(iters=20, legacy_iters=8):
    import ctypes
    from oracle import oracle as orc
    from rlobjectdetection_b200 import synthetic as syn
    from rlobjectdetection_b200.model import _backend as be
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    P = ctypes.c_void_p
    leg = orc.ref_legacy()
    if leg is not None:
        leg.nms_cuda_compute.restype = None
        leg.nms_cuda_compute.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        leg.ROIAlignForwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 6 + [P, P, P]
        leg.ROIAlignBackwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 7 + [P, P, P]
        leg.ROIPoolForwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 6 + [P, P, P, P]
        leg.ROIPoolBackwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 7 + [P, P, P, P]
    dp = lambda t: P(t.data_ptr())  # noqa: E731
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def time_us(fn, iters=iters, warm=3):
        # warm-up by TIME as well as by count: this leg follows seconds of host-only work (the CPU baseline), the SM
        # clock ramps up over the first milliseconds of load, and the issue-bound ops (RoIAlign backward, RoIPool
        # forward) read 15-25 % slow in a median of 10 samples taken right after three calls
        t_end = time.perf_counter() + 0.03
        n = 0
        while n < warm or time.perf_counter() < t_end:
            fn()
            n += 1
            if n % 4 == 0:
                torch.cuda.synchronize()
            if n >= 200:
                break
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()  # cold L2 for every sample
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts)

    peak = 6553.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    rows = []

    def add(name, alg_bytes, ours, legacy, note=""):
        r = {"op": name, "ours_us": round(ours, 1), "legacy_cuda_us": None if legacy is None else round(legacy, 1),
             "speedup": None if legacy is None else round(legacy / ours, 1), "alg_bytes": alg_bytes,
             "gbs": round(alg_bytes / ours / 1e3, 1), "hbm_frac": round(alg_bytes / ours / 1e3 / peak, 3), "note": note}
        rows.append(r)
        print(json.dumps(r), file=sys.stderr, flush=True)

    def align_case(tag, B, C, H, W, n_per, seed, fwd_only=False):
        g = torch.Generator().manual_seed(seed)
        feat = torch.randn(B, C, H, W, generator=g).to(dev)
        rois = syn.rois_for_batch(seed + 1, B, n_per, H * 16.0, W * 16.0).to(dev)
        R = rois.size(0)
        gout = torch.randn(R, C, 7, 7, generator=g).to(dev)
        fb = 4 * (B * C * H * W + 5 * R + R * C * 49)
        # forward
        ours = time_us(lambda: be.roi_align_forward(feat, rois, 7, 7, 1 / 16.0, be.POOL_AVG))
        lg = None
        if leg is not None:
            top = torch.empty(R, C, 8, 8, device=dev)

            def legacy_fwd():
                top.zero_()  # functions/roi_align.py:22
                leg.ROIAlignForwardLaucher(dp(feat), 1 / 16.0, R, H, W, C, 8, 8, dp(rois), dp(top),
                                           P(torch.cuda.current_stream().cuda_stream))
                return torch.nn.functional.avg_pool2d(top, kernel_size=2, stride=1)  # modules/roi_align.py:29
            lg = time_us(legacy_fwd, iters=legacy_iters, warm=2)
        add(f"RoIAlignAvg fwd {tag}", fb, ours, lg, f"B={B} C={C} {H}x{W} R={R}")
        if fwd_only:
            return None
        # backward
        ours = time_us(lambda: be.roi_align_backward(gout, rois, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG))
        lg = None
        if leg is not None:
            bottom = torch.empty(B, C, H, W, device=dev)

            def legacy_bwd():
                # autograd of avg_pool2d(2,1), then the legacy atomics into a zeroed buffer
                ggrid = torch.ops.aten.avg_pool2d_backward(gout, top, [2, 2], [1, 1], [0, 0], False, True, None)
                bottom.zero_()  # functions/roi_align.py:38-39
                leg.ROIAlignBackwardLaucher(dp(ggrid), 1 / 16.0, B, R, H, W, C, 8, 8, dp(rois), dp(bottom),
                                            P(torch.cuda.current_stream().cuda_stream))
            lg = time_us(legacy_bwd, iters=legacy_iters, warm=2)
        add(f"RoIAlignAvg bwd {tag}", fb, ours, lg, f"B={B} C={C} {H}x{W} R={R}")
        return feat, rois, gout

    feat, rois, gout = align_case("C2", 4, 1024, 38, 63, 256, 1)
    # RoIPool at C2
    B, C, H, W = feat.shape
    R = rois.size(0)
    pb = 4 * (B * C * H * W + 5 * R + 2 * R * C * 49)
    ours = time_us(lambda: be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0))
    out, am = be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0)
    lg = lgb = None
    if leg is not None:
        top = torch.empty(R, C, 7, 7, device=dev)
        arg = torch.empty(R, C, 7, 7, dtype=torch.int32, device=dev)
        st = lambda: P(torch.cuda.current_stream().cuda_stream)  # noqa: E731

        def legacy_pool():
            top.zero_(), arg.zero_()  # functions/roi_pool.py:17-18
            leg.ROIPoolForwardLaucher(dp(feat), 1 / 16.0, R, H, W, C, 7, 7, dp(rois), dp(top), dp(arg), st())
        lg = time_us(legacy_pool, iters=legacy_iters, warm=2)
        bottom = torch.empty(B, C, H, W, device=dev)

        def legacy_pool_bwd():
            bottom.zero_()
            leg.ROIPoolBackwardLaucher(dp(gout), 1 / 16.0, B, R, H, W, C, 7, 7, dp(rois), dp(bottom), dp(arg), st())
        lgb = time_us(legacy_pool_bwd, iters=3, warm=1)
    add("RoIPool fwd C2", pb, ours, lg)
    ours_inf = time_us(lambda: be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0, want_argmax=False))
    add("RoIPool fwd C2, inference (no argmax buffer)", 4 * (B * C * H * W + 5 * R + R * C * 49), ours_inf, lg,
        "legacy = the same ROIPoolForwardLaucher (it always writes argmax)")
    ours = time_us(lambda: be.roi_pool_backward(gout, am, rois, (B, C, H, W), 7, 7, 1 / 16.0))
    add("RoIPool bwd C2", pb, ours, lgb, "legacy = O(B*C*H*W*R) gather")
    # RoICrop (POOLING_MODE 'crop'): 14x14 sampling grid + 2x2 max pool, C2 shape
    gxy = be.affine_grid(rois, (H, W), 14, True)
    gyx = torch.stack([gxy[..., 1], gxy[..., 0]], 3).contiguous()
    cb = 4 * (B * C * H * W + R * 14 * 14 * 2 + R * C * 14 * 14)
    ours = time_us(lambda: be.roi_crop_forward(feat, gyx))
    gout14 = torch.randn(R, C, 14, 14, device=dev)
    ours_b = time_us(lambda: be.roi_crop_backward(gout14, gyx, (B, C, H, W)))
    lgf = lgb2 = None
    if leg is not None and hasattr(leg, "BilinearSamplerBHWD_updateOutput_cuda_kernel"):
        I = ctypes.c_int
        lf = leg.BilinearSamplerBHWD_updateOutput_cuda_kernel
        lf.restype, lf.argtypes = I, [I] * 8 + [P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P]
        lb = leg.BilinearSamplerBHWD_updateGradInput_cuda_kernel
        lb.restype, lb.argtypes = I, [I] * 8 + [P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P]
        o14 = torch.empty(R, C, 14, 14, device=dev)
        gs_ = (gyx.stride(0), gyx.stride(3), gyx.stride(1), gyx.stride(2))

        def legacy_crop():
            o14.zero_()  # functions/roi_crop.py:11
            lf(C, 14, 14, R, C, H, W, B, dp(feat), *feat.stride(), dp(gyx), *gs_, dp(o14), *o14.stride(), st())
        lgf = time_us(legacy_crop, iters=legacy_iters, warm=2)
        gfe, ggr = torch.empty(B, C, H, W, device=dev), torch.zeros_like(gyx)

        def legacy_crop_bwd():
            gfe.zero_()
            lb(C, 14, 14, R, C, H, W, B, dp(feat), *feat.stride(), dp(gyx), *gs_, dp(gfe), *gfe.stride(), dp(ggr), *gs_,
               dp(gout14), *gout14.stride(), st())
        lgb2 = time_us(legacy_crop_bwd, iters=4, warm=1)
    add("RoICrop 14x14 fwd C2", cb, ours, lgf)
    add("RoICrop 14x14 bwd C2", cb, ours_b, lgb2)
    del feat, rois, gout, out, am, gout14, gxy, gyx
    align_case("C4", 24, 1024, 50, 75, 300, 3)
    torch.cuda.empty_cache()
    # a map beyond one CTA's shared memory (1600 x 2400 image at stride 16): the plane kernel over overlapping tiles
    align_case("large map 100x150, 2000 rois/image (tiled planes)", 2, 1024, 100, 150, 2000, 7, fwd_only=True)
    torch.cuda.empty_cache()

    # NMS, C1 size: 12000 sorted boxes, thr 0.7
    g = torch.Generator().manual_seed(0)
    n = 12000
    bx = syn.random_boxes(g, n, 600, 1000, 8.0, 300.0)
    sc = syn.distinct_scores(g, (n,)).sort(descending=True).values
    dets = torch.cat([bx, sc[:, None]], 1).contiguous()
    d = dets.to(dev)
    ours = time_us(lambda: be.nms_padded(d, 0.7))
    lg = None
    if leg is not None:
        host = np.ascontiguousarray(dets.numpy())
        keep = torch.zeros(n, dtype=torch.int32, device=dev)
        num = torch.zeros(1, dtype=torch.int32, device=dev)

        def legacy_nms():  # as the reference runs it: per-call malloc, mask D2H, host scan (nms_cuda_kernel.cu:87-161)
            leg.nms_cuda_compute(dp(keep), dp(num), P(host.ctypes.data), n, 5, 0.7)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            legacy_nms()
        torch.cuda.synchronize()
        lg = (time.perf_counter() - t0) / 5 * 1e6
    add("nms n=12000 thr 0.7 (C1)", 20 * n + 4 * n, ours, lg, "legacy timed on the host clock: it synchronises internally")
    # C5: 64 images x 80 classes x 300 boxes, thr 0.3, one launch
    dets5, offs = syn.clustered_dets(4, 64, 80, 300, 600, 1000)
    d5, o5 = dets5.to(dev), offs.to(dev)
    ours = time_us(lambda: be.nms_batched(d5, o5, 0.3, max_seg=300))
    lg = None
    if leg is not None:
        sample = 256
        host = np.ascontiguousarray(dets5.numpy())
        keep = torch.zeros(300, dtype=torch.int32, device=dev)
        num = torch.zeros(1, dtype=torch.int32, device=dev)
        offs_h = offs.numpy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for sgi in range(sample):
            seg = host[offs_h[sgi]:offs_h[sgi + 1]]
            leg.nms_cuda_compute(dp(keep), dp(num), P(seg.ctypes.data), seg.shape[0], 5, 0.3)
        torch.cuda.synchronize()
        lg = (time.perf_counter() - t0) / sample * (offs_h.size - 1) * 1e6
    add("per-class nms 64x80x300 thr 0.3 (C5)", 20 * d5.size(0), ours, lg,
        "legacy = 5120 calls, extrapolated from 256")
    # proposal layer, C1 (TRAIN cfg) and C4 (TEST cfg, 24 images)
    for tag, Bp, A, Hh, Ww, imh, imw, pre, post, scales in (("C1 12000->2000", 1, 9, 37, 62, 600, 1000, 12000, 2000, (8, 16, 32)),
                                                              ("C4 6000->300 x24", 24, 12, 50, 75, 800, 1200, 6000, 300, SCALES)):
        scores, deltas, im_info = syn.rpn_outputs(5, Bp, A, Hh, Ww, imh, imw, imh / 600.0)
        anchors = torch.from_numpy(orc.generate_anchors(16, RATIOS, scales).astype(np.float32)).to(dev)
        sd, dd, ii = scores.to(dev), deltas.to(dev), im_info.to(dev)
        ours = time_us(lambda: be.proposal_forward(sd, dd, ii, anchors, 16, pre, post, 0.7))
        add(f"_ProposalLayer {tag}", Bp * (4 * 5 * A * Hh * Ww + 12 + 20 * post), ours, None,
            "reference = torch ops + per-image nms_gpu; no standalone legacy kernel")
    # C3: the RL refinement step alone -- 8 images x 300 rois x 16 actions vs 20 gt: rewards + best action +
    # pack (one launch) and the RoIAlignAvg re-pool of the refined boxes over 50x75x1024 features
    B3, N3 = 8, 300
    g = torch.Generator().manual_seed(2)
    feat3 = torch.randn(B3, C, FH, FW, generator=g).to(dev)
    rois3 = syn.rois_for_batch(3, B3, N3, IM_H, IM_W, edge_cases=False).view(B3, N3, 5).to(dev)
    gt3 = syn.gt_boxes(21, B3, G, IM_H, IM_W)[0].to(dev)
    act3 = torch.from_numpy(orc.action_table(list(ACT_DELTA))).to(dev)

    def c3_step():
        t = be.reward_refine(rois3, gt3, act3, want=("reward", "label", "weight", "refined"))
        return be.roi_align_forward(feat3, t["refined"].view(-1, 5), POOL, POOL, 1.0 / STRIDE, be.POOL_AVG)
    R3 = B3 * N3
    c3_bytes = 4 * (4 * R3 + 4 * B3 * G + 4 * 16 + R3 * 16) + 4 * (B3 * C * FH * FW + 5 * R3 + R3 * C * POOL * POOL)
    add("RL refine step C3 (reward+refine+re-pool, 8 img)", c3_bytes, time_us(c3_step), None,
        "reference = per-(box, action) Python loop over bbIou in DataLoader workers + legacy RoIAlign")
    add("  of which reward+refine+pack kernel", 4 * (4 * R3 + 4 * B3 * G + 4 * 16 + R3 * 16),
        time_us(lambda: be.reward_refine(rois3, gt3, act3, want=("reward", "label", "weight", "refined"))), None,
        "latency-bound: 195 KB of traffic")
    return {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak,
            "timing": "CUDA events, median, L2 flushed before every sample", "ops": rows}


def ops_summary(table):
    """Compact form for the bench line: {op: [ours_us, legacy_cuda_us, speedup, hbm_frac]} + the C2 train step."""
    rows = {r["op"].strip(): [r["ours_us"], r["legacy_cuda_us"], r["speedup"], r["hbm_frac"]] for r in table["ops"]}
    out = {"columns": ["ours_us", "legacy_cuda_us", "speedup_vs_legacy_cuda", "hbm_frac"], "timing": table["timing"],
           "rows": rows}
    f, b = rows.get("RoIAlignAvg fwd C2"), rows.get("RoIAlignAvg bwd C2")
    if f and b:
        alg = 2 * 244764672  # C2 forward + backward
        t = f[0] + b[0]
        out["train_step_c2_fwd_bwd"] = {"us": round(t, 1), "target_us_60pct": 124.5, "alg_bytes": alg,
                                        "hbm_frac": round(alg / t / 1e3 / table["hbm_peak_gbs"], 3),
                                        "legacy_cuda_us": None if f[1] is None or b[1] is None else round(f[1] + b[1], 1)}
    return out


def run_ops(args):
    out = measure_ops()
    with open(os.path.join(ROOT, "profiles", "ops_latest.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ops", action="store_true", help="skip the per-config `ops` object of the bench line")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the timed step")
    ap.add_argument("--eager", action="store_true", help="N > 1: eager launches instead of the CUDA graph")
    ap.add_argument("--repool", default="merged", choices=["merged", "separate"],
                    help="merged: the proposals and the refined boxes are pooled by one RoIAlign call (default); "
                         "separate: two calls")
    ap.add_argument("--ops", action="store_true", help="per-op table vs the reference's legacy CUDA kernels")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.ops:
        if rank == 0:
            run_ops(args)
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
