"""The whole hot path as one device-resident step: proposal -> NMS -> RoIAlignAvg ->
RL-refine (action rewards, apply the best positive action per box) -> RoIAlignAvg re-pool
[-> RoIAlign backward].  This is what bench.py times; it strings together the same modules a
user of the reference would call (_ProposalLayer, RoIAlignAvg, Action) -- nothing here computes
on the host and nothing synchronises until the caller reads a result."""
import collections

import torch

from .model import _backend as be
from .model.Reinforcement.action import Action
from .model.Reinforcement.reward import IOU_RCNN, action_rewards
from .model.roi_align.modules.roi_align import RoIAlignAvg
from .model.rpn.proposal_layer import _ProposalLayer


class DetectRefineStep:
    """Two streams per device.  Everything that is per-image and latency-bound -- the proposal
    layer's select / sort / NMS (one CTA per image), the action rewards and the refine -- runs
    on a private "light" stream; the two RoIAlign calls (a full-GPU kernel each) run on the
    caller's stream and wait, by event, for the rois they pool.  Within a call that only takes
    reward + refine off the critical path; across calls, when the caller says its inputs are
    ready (`inputs_ready`), the light stream runs ahead and step i+1's proposal work fills
    the SMs' idle slots under step i's RoIAlign kernels -- the per-image kernels use 24 of 148
    SMs, so they cost next to nothing there.  Results are ordered on the caller's stream as
    usual: what the caller gets back are copies made on ITS stream; the light stream's own
    tensors are kept alive by the step until the caller's stream has read them, and the light
    stream may be at most `max_ahead` steps ahead (Tensor.record_stream would do the same
    bookkeeping inside the caching allocator, but with a run-ahead producer it costs ~0.5 ms
    of host time per step: the light pool cannot recycle blocks and keeps growing)."""

    def __init__(self, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), cfg_key="TEST",
                 pool=7, act_delta=(0.5, 0.25), backward=True):
        self.proposal = _ProposalLayer(feat_stride, list(scales), list(ratios))
        self.align = RoIAlignAvg(pool, pool, 1.0 / feat_stride)
        self.action = Action(list(act_delta))
        self.cfg_key = cfg_key
        self.pool = pool
        self.scale = 1.0 / feat_stride
        self.backward = backward
        self.max_ahead = 2
        self._light = {}
        self._inflight = collections.deque()  # (event on the caller's stream, the light stream's tensors)

    def _light_stream(self, device):
        key = (device.type, device.index)
        if key not in self._light:
            self._light[key] = torch.cuda.Stream(device=device)
        return self._light[key]

    @torch.no_grad()
    def __call__(self, scores, deltas, im_info, feat, gt, grad_pooled=None, inputs_ready=None):
        """scores (B,2A,H,W), deltas (B,4A,H,W), im_info (B,3), feat (B,C,H,W), gt (B,G,4)
        x1y1x2y2; grad_pooled (B*post,C,pool,pool) is the upstream gradient of the re-pooled
        features (what layer4 would send back) when backward is on.

        inputs_ready: None (default) = the inputs are ordered on the current stream, the light
        stream waits for it (no overlap across calls); a torch.cuda.Event = they are ready once
        it has fired; True = they are resident and nobody is writing them."""
        cur = torch.cuda.current_stream(feat.device)
        light = self._light_stream(feat.device)
        if inputs_ready is None:
            light.wait_stream(cur)
        elif inputs_ready is not True:
            light.wait_event(inputs_ready)
        # bounded run-ahead; dropping a step's tensors after this wait lets the light stream's
        # allocator reuse them safely (the reuse is ordered after the caller's reads)
        while len(self._inflight) > self.max_ahead:
            consumed, _ = self._inflight.popleft()
            light.wait_event(consumed)
        with torch.cuda.stream(light):
            l_rois = self.proposal((scores, deltas, im_info, self.cfg_key))   # (B, post, 5)
            have_rois = torch.cuda.Event()
            have_rois.record(light)
            B, N, _ = l_rois.shape
            l_reward, l_label, l_weight = action_rewards(self.action, l_rois[:, :, 1:5], gt, mode=IOU_RCNN)
            l_refined = l_rois.clone()
            # every box takes its best action if that action's label is +1 (move_from_act with
            # maxk = N and the rewards as predictions)
            l_moved = be.move_from_act(l_refined, l_reward, l_label, self.action.table(l_rois.device), N,
                                       corners=True)
            have_refined = torch.cuda.Event()
            have_refined.record(light)
        cur.wait_event(have_rois)
        pooled = self.align(feat, l_rois.view(-1, 5))                         # (B*N, C, p, p)
        cur.wait_event(have_refined)
        pooled_refined = self.align(feat, l_refined.view(-1, 5))
        rois, reward, label, weight, refined, moved = (t.clone() for t in (l_rois, l_reward, l_label, l_weight,
                                                                            l_refined, l_moved))
        consumed = torch.cuda.Event()
        consumed.record(cur)
        self._inflight.append((consumed, (l_rois, l_reward, l_label, l_weight, l_refined, l_moved)))
        out = dict(rois=rois, pooled=pooled, reward=reward, label=label, weight=weight,
                   refined=refined, moved=moved, pooled_refined=pooled_refined)
        if self.backward:
            if grad_pooled is None:
                raise ValueError("backward=True needs grad_pooled")
            out["grad_feat"] = be.roi_align_backward(grad_pooled, refined.view(-1, 5), None,
                                                     tuple(feat.shape), self.pool, self.pool,
                                                     self.scale, be.POOL_AVG)
        return out
