"""The whole hot path as one device-resident step: proposal -> NMS -> RoIAlignAvg ->
RL-refine (action rewards, apply the best positive action per box) -> RoIAlignAvg re-pool
[-> RoIAlign backward].  This is what bench.py times; it strings together the same modules a
user of the reference would call (_ProposalLayer, RoIAlignAvg, Action) -- nothing here computes
on the host and nothing synchronises until the caller reads a result.

Two forms:
  * DetectRefineStep(...)(scores, deltas, im_info, feat, gt): eager launches on two streams;
  * DetectRefineStep.capture(...) -> GraphedStep: the same launches recorded once into a CUDA
    graph over fixed input buffers and replayed with one host call per step -- for shards so
    small (3 images per GPU when BASELINE config 4's batch of 24 is split over 8 GPUs) that
    the ~25 launches of a step cost more host time than the GPU needs to run them."""
import collections
import os

import torch

from .model import _backend as be
from .model.Reinforcement.action import Action, exp_abs, wtrans_code
from .model.roi_align.modules.roi_align import RoIAlignAvg
from .model.rpn.proposal_layer import _ProposalLayer

LIGHT_NAMES = ("rois", "reward", "label", "weight", "refined", "moved", "packed")


class DetectRefineStep:
    """Two streams per device.  Everything that is per-image and latency-bound -- the proposal
    layer's select / sort / NMS and the fused reward + refine + pack kernel -- runs on a private
    "light" stream; the two RoIAlign calls (a full-GPU kernel each) run on the caller's stream
    and wait, by event, for the rois they pool.  Within a call that only takes reward + refine
    off the critical path.  Across calls the caller can hand over the NEXT step's inputs
    (`next_inputs`): their light work is enqueued between this step's two RoIAlign launches, so
    it is dispatched in the tail of the first one and runs under the second, and the next call
    finds its rois ready.  (Enqueue order matters: nothing with more than ~0.9 KB of shared
    memory can co-reside with two RoIAlign CTAs, so a light kernel is only dispatched once every
    CTA of the RoIAlign launch enqueued before it has been dispatched.)
    Results are ordered on the caller's stream as usual: what the caller gets back are copies
    made on ITS stream; the light stream's own tensors are kept alive by the step until the
    caller's stream has read them, and at most `max_ahead` steps of them exist
    (Tensor.record_stream would do the same bookkeeping inside the caching allocator, but with
    a run-ahead producer it costs ~0.5 ms of host time per step: the light pool cannot recycle
    blocks and keeps growing).

    outputs: which of the light stream's small tensors a call hands back -- "rois" (B,N,5),
    "reward" / "label" / "weight" (B,N,A), "refined" (B,N,5), "moved" (1,) and "packed"
    (B,N,5+A) = [image index + first_image, refined box, rewards], the row layout of the
    multi-GPU gather (shard.py)."""

    def __init__(self, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), cfg_key="TEST",
                 pool=7, act_delta=(0.5, 0.25), backward=True,
                 outputs=("rois", "reward", "label", "weight", "refined", "moved"), first_image=0, repool="merged"):
        self.proposal = _ProposalLayer(feat_stride, list(scales), list(ratios))
        self.align = RoIAlignAvg(pool, pool, 1.0 / feat_stride)
        self.action = Action(list(act_delta), wtrans=exp_abs)  # Config.act_wtrans (config.py:48-51)
        self.cfg_key = cfg_key
        self.pool = pool
        self.scale = 1.0 / feat_stride
        self.backward = backward
        unknown = set(outputs) - set(LIGHT_NAMES)
        if unknown:
            raise ValueError(f"unknown outputs {sorted(unknown)}")
        self.outputs = tuple(outputs)
        self.first_image = int(first_image)  # global index of this shard's image 0 (packed rows)
        # repool = "merged": the proposals and the refined boxes are pooled by ONE RoIAlign call on their
        # concatenation -- in this step the refined boxes come from the IoU rewards, not from the pooled
        # features, so both roi sets exist before the first pooling; one call fills every feature plane once
        # instead of twice and plans once (C4: 858 -> 776 us for the two poolings, bit-identical outputs).
        # "separate": two calls, for a caller whose refinement depends on the first pooling (a policy network
        # between the two, as in the reference's RL loop).
        if repool not in ("merged", "separate"):
            raise ValueError("repool must be 'merged' or 'separate'")
        self.repool = repool
        # RLOD_PLAN_AHEAD=1 plans the rois of the pooling call on the light stream (RoIAlignAvg.plan / forward_planned).
        # Off by default: measured at C4, N=1 it lengthens the step (0.892 -> 0.941 ms) -- the plan kernels then
        # compete with the running pooling for SM slots and stretch it by more than their own time
        self.plan_ahead = os.environ.get("RLOD_PLAN_AHEAD", "0") == "1"
        self.max_ahead = 2
        self._light = {}
        self._inflight = collections.deque()  # (event on the caller's stream, the light stream's tensors)
        self._prefetched = None               # (key, light-work record) of next_inputs

    def _light_stream(self, device):
        key = (device.type, device.index)
        if key not in self._light:
            self._light[key] = s = torch.cuda.Stream(device=device)
            # Pre-size the caching allocator's pool of this stream (pools are per stream): its small tensors
            # and workspaces then never need a cudaMalloc in steady state.  A cudaMalloc blocks the launching
            # thread for 0.5-70 ms while the GPU is busy -- measured as a stalled first step after every
            # synchronise as long as the pool was one 2 MB segment short of its steady-state size.
            with torch.cuda.stream(s):
                spare = [torch.empty(512 << 10, dtype=torch.uint8, device=device) for _ in range(32)]
                spare += [torch.empty(8 << 20, dtype=torch.uint8, device=device) for _ in range(8)]
            del spare
        return self._light[key]

    @staticmethod
    def _key(scores, deltas, im_info, gt):
        return tuple((t.data_ptr(), tuple(t.shape), t._version) for t in (scores, deltas, im_info, gt))

    def _light_kernels(self, scores, deltas, im_info, gt, after_rois=None, feat_size=None):
        """proposal -> NMS -> fused reward / refine / pack on the CURRENT stream: three launches.
        Returns {name: tensor} with rois, refined and the requested outputs."""
        rois = self.proposal((scores, deltas, im_info, self.cfg_key))          # (B, post, 5)
        if after_rois is not None:
            after_rois()
        want = tuple(n for n in LIGHT_NAMES[1:] if n in self.outputs or n == "refined")
        t = be.reward_refine(rois, gt, self.action.table(rois.device), iou_thres=float(self.action.iou_thres),
                             first_image=self.first_image, wtrans=_kernel_wtrans(self.action), want=want)
        t["rois"] = rois
        if self.repool == "merged":
            t["both"] = torch.cat([rois.view(-1, 5), t["refined"].view(-1, 5)])   # (2 B N, 5): proposals, then refined
            if feat_size is not None and self.plan_ahead:
                # the plan of the one pooling call, here on the light stream: ~45 us of small launches that leave
                # the GPU mostly idle when they sit between two poolings on the caller's stream
                t.plan = self.align.plan(t["both"], feat_size)
        return t

    def _pool(self, feat, lt, between=None):
        """pooled, pooled_refined for the light stream's roi sets: one RoIAlign call (merged) or two; `between`
        runs after the first launch has been enqueued (the next step's light work goes there)."""
        if self.repool == "merged":
            plan = getattr(lt, "plan", None)
            if plan is not None and plan.geometry[:4] == tuple(feat.shape) and not torch.is_grad_enabled():
                both = self.align.forward_planned(feat, plan)
            else:
                both = self.align(feat, lt["both"])
            if between is not None:
                between()
            n = both.size(0) // 2
            return both[:n], both[n:]
        pooled = self.align(feat, lt["rois"].view(-1, 5))
        if between is not None:
            between()
        return pooled, self.align(feat, lt["refined"].view(-1, 5))

    def _light_work(self, cur, light, scores, deltas, im_info, gt, ready, feat_size=None):
        """The light kernels on the light stream; returns the events the caller's stream has to
        wait for and the (light-stream-owned) tensors."""
        if ready is None:
            light.wait_stream(cur)
        elif ready is not True:
            light.wait_event(ready)
        # bounded run-ahead; dropping a step's tensors after this wait lets the light stream's
        # allocator reuse them safely (the reuse is ordered after the caller's reads)
        while len(self._inflight) > self.max_ahead:
            consumed, _ = self._inflight.popleft()
            light.wait_event(consumed)
        with torch.cuda.stream(light):
            have_rois = torch.cuda.Event()
            t = self._light_kernels(scores, deltas, im_info, gt, after_rois=lambda: have_rois.record(light),
                                    feat_size=feat_size)
            have_refined = torch.cuda.Event()
            have_refined.record(light)
        return have_rois, have_refined, t

    @torch.no_grad()
    def __call__(self, scores, deltas, im_info, feat, gt, grad_pooled=None, inputs_ready=None,
                 next_inputs=None, next_ready=None):
        """scores (B,2A,H,W), deltas (B,4A,H,W), im_info (B,3), feat (B,C,H,W), gt (B,G,4)
        x1y1x2y2; grad_pooled (B*post,C,pool,pool) is the upstream gradient of the re-pooled
        features (what layer4 would send back) when backward is on.

        inputs_ready: None (default) = the inputs are ordered on the current stream, the light
        stream waits for it; a torch.cuda.Event = they are ready once it has fired; True = they
        are resident and nobody is writing them.
        next_inputs: (scores, deltas, im_info, gt) of the NEXT call, with next_ready like
        inputs_ready (None is not allowed: the caller's stream is busy with this step).  The
        next call recognises them by pointer, shape and version; anything else is recomputed."""
        cur = torch.cuda.current_stream(feat.device)
        light = self._light_stream(feat.device)
        rec = None
        if self._prefetched is not None:
            key, cand = self._prefetched
            self._prefetched = None
            if key == self._key(scores, deltas, im_info, gt):
                rec = cand
            else:  # not what was announced: keep its tensors alive until the light stream is past them
                consumed = torch.cuda.Event()
                consumed.record(light)
                self._inflight.append((consumed, cand[2]))
        if rec is None:
            rec = self._light_work(cur, light, scores, deltas, im_info, gt, inputs_ready, tuple(feat.shape))
        have_rois, have_refined, lt = rec
        if next_inputs is not None and next_ready is None:
            raise ValueError("next_inputs need next_ready = True or an event")

        def enqueue_next():
            if next_inputs is not None:
                ns, nd, ni, ng = next_inputs
                self._prefetched = (self._key(ns, nd, ni, ng),
                                    self._light_work(cur, light, ns, nd, ni, ng, next_ready, tuple(feat.shape)))

        if self.repool == "merged":
            cur.wait_event(have_refined)
            pooled, pooled_refined = self._pool(feat, lt, enqueue_next)         # (B*N, C, p, p) each, one launch
        else:
            cur.wait_event(have_rois)
            pooled = self.align(feat, lt["rois"].view(-1, 5))
            enqueue_next()
            cur.wait_event(have_refined)
            pooled_refined = self.align(feat, lt["refined"].view(-1, 5))
        out = _copy_light(lt, [n for n in LIGHT_NAMES if n in lt and (n in self.outputs or (n == "refined" and self.backward))])
        consumed = torch.cuda.Event()
        consumed.record(cur)
        self._inflight.append((consumed, lt))
        out.update(pooled=pooled, pooled_refined=pooled_refined)
        if self.backward:
            if grad_pooled is None:
                raise ValueError("backward=True needs grad_pooled")
            out["grad_feat"] = be.roi_align_backward(grad_pooled, out["refined"].view(-1, 5), None,
                                                     tuple(feat.shape), self.pool, self.pool,
                                                     self.scale, be.POOL_AVG)
        return out

    def capture(self, scores, deltas, im_info, feat, gt, next_inputs=None):
        """Record the step over these (fixed) input buffers into a CUDA graph -> GraphedStep."""
        return GraphedStep(self, (scores, deltas, im_info, feat, gt), next_inputs)


def _copy_light(lt, names):
    """The caller's copies of the light stream's tensors, made on the CURRENT stream: everything the fused
    reward / refine kernel wrote lives in one allocation (lt.flat), so it is one device copy for all of
    them plus one for the rois (two launches on the critical stream instead of one per tensor)."""
    out = {}
    flat = getattr(lt, "flat", None)
    fused = [n for n in names if n != "rois"] if flat is not None else []
    if len(fused) > 1:
        mine = flat.clone()
        base = flat.data_ptr()
        for n in fused:
            t = lt[n]
            off = (t.data_ptr() - base) // 4
            v = mine[off:off + t.numel()]
            out[n] = (v.view(torch.int32) if t.dtype == torch.int32 else v).view(t.shape)
    for n in names:
        if n not in out:
            out[n] = lt[n].clone()
    return out


def _kernel_wtrans(action):
    code = wtrans_code(action)
    if code == be.WTRANS_RAW:
        raise ValueError("DetectRefineStep: Action.wtrans must be the identity or exp(|x|) (the fused refine kernel "
                         "applies it on the device); use model.Reinforcement.reward.action_rewards for other callables")
    return code


class GraphedStep:
    """One CUDA graph = one whole step over fixed input buffers; `replay()` is a single host
    call (cudaGraphLaunch) instead of ~25 launches, and returns the step's outputs in STATIC
    buffers (valid until the next replay, like any captured output).

    next_inputs = None: the light kernels of these inputs, then the two RoIAlign launches
    (the fused reward / refine kernel runs on a forked branch under the first RoIAlign).
    next_inputs = (scores, deltas, im_info, gt) of the step that follows (the same buffers in a
    steady loop): the graph pools the rois that the PREVIOUS replay's light branch left in the
    step's hand-over buffers while its own light branch prepares the next step's -- the
    software pipeline of DetectRefineStep's `next_inputs`, frozen into the graph.  `prime()`
    fills the hand-over buffers for the first replay (and after the inputs changed by more
    than one step)."""

    def __init__(self, step, inputs, next_inputs=None):
        if step.backward:
            raise ValueError("GraphedStep covers the forward step (backward=False)")
        self.step = step
        self.inputs = inputs
        self.next_inputs = next_inputs
        scores, deltas, im_info, feat, gt = inputs
        dev = feat.device
        self.device = dev
        self._cap = torch.cuda.Stream(device=dev)
        self._side = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.out = None
        self._hand = None  # hand-over buffers (pipelined form): rois / refined / outputs of the step to pool next
        profiling = be.lib().rlod_profile_enable(0)  # event records cannot be timed from inside a graph
        try:
            with torch.no_grad():
                self._warm_and_capture()
        finally:
            be.lib().rlod_profile_enable(profiling)

    def _heavy(self, feat, lt):
        return self.step._pool(feat, lt)

    def _warm_and_capture(self):
        step, cap, side = self.step, self._cap, self._side
        scores, deltas, im_info, feat, gt = self.inputs
        cap.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cap):
            # eager warm-up: function attributes, cached tables, allocator pools
            lt = step._light_kernels(scores, deltas, im_info, gt, feat_size=tuple(feat.shape))
            self._heavy(feat, lt)
            if self.next_inputs is not None:
                # what the next replay needs of this step's light results: the roi set(s) it pools and the outputs
                keep = set(step.outputs) | ({"both"} if step.repool == "merged" else {"rois", "refined"})
                self._hand = be.FusedOutputs((k, v.clone()) for k, v in lt.items() if k in keep)
                plan = getattr(lt, "plan", None)
                if plan is not None:  # the hand-over carries the planned rois of the step to pool next
                    self._hand.plan = be.RoiAlignPlan(plan.ws.clone(), plan.geometry, plan.n_rois, None)
            del lt
        cap.synchronize()
        with torch.cuda.graph(self.graph, stream=cap):
            if self.next_inputs is None:
                rois = step.proposal((scores, deltas, im_info, step.cfg_key))
                side.wait_stream(cap)                                    # fork: refine under the first RoIAlign
                with torch.cuda.stream(side):
                    want = tuple(n for n in LIGHT_NAMES[1:] if n in step.outputs or n == "refined")
                    lt = be.reward_refine(rois, gt, step.action.table(self.device),
                                          iou_thres=float(step.action.iou_thres), first_image=step.first_image,
                                          wtrans=_kernel_wtrans(step.action), want=want)
                    lt["rois"] = rois
                    if step.repool == "merged":
                        lt["both"] = torch.cat([rois.view(-1, 5), lt["refined"].view(-1, 5)])
                if step.repool == "merged":
                    cap.wait_stream(side)                                # join: one launch pools both roi sets
                    pooled, pooled_refined = step._pool(feat, lt)
                else:
                    pooled = step.align(feat, rois.view(-1, 5))
                    cap.wait_stream(side)                                # join
                    pooled_refined = step.align(feat, lt["refined"].view(-1, 5))
                out = {n: lt[n] for n in LIGHT_NAMES if n in lt and n in step.outputs}
            else:
                ns, nd, ni, ng = self.next_inputs
                hand = self._hand
                side.wait_stream(cap)                                    # fork: the NEXT step's light kernels
                with torch.cuda.stream(side):
                    nxt = step._light_kernels(ns, nd, ni, ng, feat_size=tuple(feat.shape))
                pooled, pooled_refined = self._heavy(feat, hand)
                out = {n: hand[n].clone() for n in LIGHT_NAMES if n in hand and n in step.outputs}
                cap.wait_stream(side)                                    # join, then hand over
                for k in hand:
                    hand[k].copy_(nxt[k])
                if getattr(hand, "plan", None) is not None:
                    hand.plan.ws.copy_(nxt.plan.ws)
            out.update(pooled=pooled, pooled_refined=pooled_refined)
            self.out = out
        cap.synchronize()

    @torch.no_grad()
    def prime(self):
        """Pipelined form: run the light kernels of `inputs` now so that the next replay pools them."""
        if self._hand is None:
            return
        scores, deltas, im_info, feat, gt = self.inputs
        cur = torch.cuda.current_stream(self.device)
        lt = self.step._light_kernels(scores, deltas, im_info, gt, feat_size=tuple(feat.shape))
        for k in self._hand:
            self._hand[k].copy_(lt[k])
        if getattr(self._hand, "plan", None) is not None:
            self._hand.plan.ws.copy_(lt.plan.ws)
        cur.synchronize()

    def replay(self):
        """Enqueue the whole step on the current stream; returns the static output dict."""
        self.graph.replay()
        return self.out
