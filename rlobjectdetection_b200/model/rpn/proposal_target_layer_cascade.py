"""_ProposalTargetLayer (lib/model/rpn/proposal_target_layer_cascade.py:19-213) over
rlod_proposal_target.  forward(all_rois, gt_boxes, num_boxes) -> rois, labels, bbox_targets,
bbox_inside_weights, bbox_outside_weights, exactly the reference's tuple.

The reference samples with np.random inside the layer (after device->host reads of the fg / bg
counts).  Here the random numbers are drawn on the device (torch.rand with `generator`, or pass
fg_keys / bg_u explicitly) and the selection is a pure function of them: the fg rois with the
smallest keys, bg_inds[floor(u * n)] for the rest -- the reference's rules with the RNG factored
out; no host synchronisation."""
import numpy as np
import torch
import torch.nn as nn

from .. import _backend as be
from ..utils.config import cfg


class _ProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super(_ProposalTargetLayer, self).__init__()
        self._num_classes = nclasses
        self.generator = None   # torch.Generator on the device, for reproducible sampling
        self.last_status = None

    def forward(self, all_rois, gt_boxes, num_boxes, fg_keys=None, bg_u=None):
        B, N, _ = all_rois.shape
        G = gt_boxes.size(1)
        rois_per_image = int(cfg.TRAIN.BATCH_SIZE / 1)                       # :46-47
        fg_rois_per_image = int(np.round(cfg.TRAIN.FG_FRACTION * rois_per_image))
        fg_rois_per_image = 1 if fg_rois_per_image == 0 else fg_rois_per_image
        dev = all_rois.device
        if fg_keys is None:
            fg_keys = torch.rand(B, N + G, device=dev, generator=self.generator)
        if bg_u is None:
            bg_u = torch.rand(B, rois_per_image, device=dev, generator=self.generator)
        norm = cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED
        rois, labels, targets, inside, outside, status = be.proposal_target(
            all_rois, gt_boxes, fg_keys, bg_u, rois_per_image, fg_rois_per_image, cfg.TRAIN.FG_THRESH,
            cfg.TRAIN.BG_THRESH_HI, cfg.TRAIN.BG_THRESH_LO,
            means=cfg.TRAIN.BBOX_NORMALIZE_MEANS if norm else None,
            stds=cfg.TRAIN.BBOX_NORMALIZE_STDS if norm else None, inside_weights=cfg.TRAIN.BBOX_INSIDE_WEIGHTS)
        self.last_status = status  # device tensor; 1 = the reference would have raised ValueError (:196-197)
        return rois, labels, targets, inside, outside
