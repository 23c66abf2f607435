"""The whole hot path as one device-resident step: proposal -> NMS -> RoIAlignAvg ->
RL-refine (action rewards, apply the best positive action per box) -> RoIAlignAvg re-pool
[-> RoIAlign backward].  This is what bench.py times; it strings together the same modules a
user of the reference would call (_ProposalLayer, RoIAlignAvg, Action) -- nothing here computes
on the host and nothing synchronises until the caller reads a result."""
import collections

import torch

from .model import _backend as be
from .model.Reinforcement.action import Action, exp_abs
from .model.Reinforcement.reward import IOU_RCNN, action_rewards
from .model.roi_align.modules.roi_align import RoIAlignAvg
from .model.rpn.proposal_layer import _ProposalLayer


class DetectRefineStep:
    """Two streams per device.  Everything that is per-image and latency-bound -- the proposal
    layer's select / sort / NMS (one CTA per image), the action rewards and the refine -- runs
    on a private "light" stream; the two RoIAlign calls (a full-GPU kernel each) run on the
    caller's stream and wait, by event, for the rois they pool.  Within a call that only takes
    reward + refine off the critical path.  Across calls the caller can hand over the NEXT
    step's inputs (`next_inputs`): their light work is enqueued between this step's two
    RoIAlign launches, so it is dispatched in the tail of the first one and runs under the
    second -- the per-image kernels use 24 of 148 SMs, so they cost next to nothing there --
    and the next call finds its rois ready.  (Enqueue order matters: a proposal CTA needs
    180 KB of shared memory and two resident RoIAlign CTAs leave an SM none, so it is only
    dispatched once every CTA of the RoIAlign launch enqueued before it has been dispatched.)
    Results are ordered on the caller's stream as usual: what the caller gets back are
    copies made on ITS stream; the light stream's own tensors are kept alive by the step until
    the caller's stream has read them, and at most `max_ahead` steps of them exist
    (Tensor.record_stream would do the same bookkeeping inside the caching allocator, but with
    a run-ahead producer it costs ~0.5 ms of host time per step: the light pool cannot recycle
    blocks and keeps growing)."""

    def __init__(self, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), cfg_key="TEST",
                 pool=7, act_delta=(0.5, 0.25), backward=True,
                 outputs=("rois", "reward", "label", "weight", "refined", "moved")):
        self.proposal = _ProposalLayer(feat_stride, list(scales), list(ratios))
        self.align = RoIAlignAvg(pool, pool, 1.0 / feat_stride)
        self.action = Action(list(act_delta), wtrans=exp_abs)  # Config.act_wtrans (config.py:48-51)
        self.cfg_key = cfg_key
        self.pool = pool
        self.scale = 1.0 / feat_stride
        self.backward = backward
        self.outputs = tuple(outputs)  # which of the light stream's small tensors a call hands back (as copies)
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

    def _light_work(self, cur, light, scores, deltas, im_info, gt, ready):
        """proposal -> NMS -> rewards -> refine on the light stream; returns the events the
        caller's stream has to wait for and the (light-stream-owned) tensors."""
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
            rois = self.proposal((scores, deltas, im_info, self.cfg_key))     # (B, post, 5)
            have_rois = torch.cuda.Event()
            have_rois.record(light)
            N = rois.size(1)
            reward, label, weight = action_rewards(self.action, rois[:, :, 1:5], gt, mode=IOU_RCNN)
            refined = rois.clone()
            # every box takes its best action if that action's label is +1 (move_from_act with
            # maxk = N and the rewards as predictions)
            moved = be.move_from_act(refined, reward, label, self.action.table(rois.device), N, corners=True)
            have_refined = torch.cuda.Event()
            have_refined.record(light)
        return have_rois, have_refined, (rois, reward, label, weight, refined, moved)

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
            rec = self._light_work(cur, light, scores, deltas, im_info, gt, inputs_ready)
        have_rois, have_refined, l_tensors = rec
        l_rois, l_refined = l_tensors[0], l_tensors[4]
        cur.wait_event(have_rois)
        pooled = self.align(feat, l_rois.view(-1, 5))                         # (B*N, C, p, p)
        if next_inputs is not None:
            if next_ready is None:
                raise ValueError("next_inputs need next_ready = True or an event")
            ns, nd, ni, ng = next_inputs
            self._prefetched = (self._key(ns, nd, ni, ng), self._light_work(cur, light, ns, nd, ni, ng, next_ready))
        cur.wait_event(have_refined)
        pooled_refined = self.align(feat, l_refined.view(-1, 5))
        names = ("rois", "reward", "label", "weight", "refined", "moved")
        out = {n: t.clone() for n, t in zip(names, l_tensors) if n in self.outputs or (n == "refined" and self.backward)}
        consumed = torch.cuda.Event()
        consumed.record(cur)
        self._inflight.append((consumed, l_tensors))
        out.update(pooled=pooled, pooled_refined=pooled_refined)
        if self.backward:
            if grad_pooled is None:
                raise ValueError("backward=True needs grad_pooled")
            out["grad_feat"] = be.roi_align_backward(grad_pooled, out["refined"].view(-1, 5), None,
                                                     tuple(feat.shape), self.pool, self.pool,
                                                     self.scale, be.POOL_AVG)
        return out
