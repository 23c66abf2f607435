"""The whole hot path as one device-resident step: proposal -> NMS -> RoIAlignAvg ->
RL-refine (action rewards, apply the best positive action per box) -> RoIAlignAvg re-pool
[-> RoIAlign backward].  This is what bench.py times; it strings together the same modules a
user of the reference would call (_ProposalLayer, RoIAlignAvg, Action) -- nothing here computes
on the host and nothing synchronises until the caller reads a result."""
import torch

from .model import _backend as be
from .model.Reinforcement.action import Action
from .model.Reinforcement.reward import IOU_RCNN, action_rewards
from .model.roi_align.modules.roi_align import RoIAlignAvg
from .model.rpn.proposal_layer import _ProposalLayer


class DetectRefineStep:
    def __init__(self, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), cfg_key="TEST",
                 pool=7, act_delta=(0.5, 0.25), backward=True):
        self.proposal = _ProposalLayer(feat_stride, list(scales), list(ratios))
        self.align = RoIAlignAvg(pool, pool, 1.0 / feat_stride)
        self.action = Action(list(act_delta))
        self.cfg_key = cfg_key
        self.pool = pool
        self.scale = 1.0 / feat_stride
        self.backward = backward

    @torch.no_grad()
    def __call__(self, scores, deltas, im_info, feat, gt, grad_pooled=None):
        """scores (B,2A,H,W), deltas (B,4A,H,W), im_info (B,3), feat (B,C,H,W), gt (B,G,4)
        x1y1x2y2; grad_pooled (B*post,C,pool,pool) is the upstream gradient of the re-pooled
        features (what layer4 would send back) when backward is on."""
        rois = self.proposal((scores, deltas, im_info, self.cfg_key))         # (B, post, 5)
        B, N, _ = rois.shape
        pooled = self.align(feat, rois.view(-1, 5))                           # (B*N, C, p, p)
        reward, label, weight = action_rewards(self.action, rois[:, :, 1:5], gt, mode=IOU_RCNN)
        refined = rois.clone()
        # every box takes its best action if that action's label is +1 (move_from_act with
        # maxk = N and the rewards as predictions)
        moved = be.move_from_act(refined, reward, label, self.action.table(rois.device), N,
                                 corners=True)
        pooled_refined = self.align(feat, refined.view(-1, 5))
        out = dict(rois=rois, pooled=pooled, reward=reward, label=label, weight=weight,
                   refined=refined, moved=moved, pooled_refined=pooled_refined)
        if self.backward:
            if grad_pooled is None:
                raise ValueError("backward=True needs grad_pooled")
            out["grad_feat"] = be.roi_align_backward(grad_pooled, refined.view(-1, 5), None,
                                                     tuple(feat.shape), self.pool, self.pool,
                                                     self.scale, be.POOL_AVG)
        return out
