"""_AnchorTargetLayer (lib/model/rpn/anchor_target_layer.py:31-192) over rlod_anchor_target.
forward((rpn_cls_score, gt_boxes, im_info, num_boxes)) -> [labels (B,1,A*H,W), bbox_targets,
bbox_inside_weights, bbox_outside_weights (B,4A,H,W)], the reference's list.

Subsampling: the reference disables fg_inds[np.random.permutation(n)[:n - num_fg]] (host RNG after
a device->host read); here the anchors with the smallest random keys are disabled, keys drawn on
the device (`generator`) or passed in -- same rule, RNG factored out, no synchronisation."""
import numpy as np
import torch
import torch.nn as nn

from .. import _backend as be
from ..utils.config import cfg
from .generate_anchors import generate_anchors


class _AnchorTargetLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super(_AnchorTargetLayer, self).__init__()
        self._feat_stride = feat_stride
        self._scales = scales
        self._anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self._num_anchors = self._anchors.size(0)
        self._allowed_border = 0
        self.generator = None

    def forward(self, input, keys=None):
        rpn_cls_score, gt_boxes, im_info = input[0], input[1], input[2]
        H, W = rpn_cls_score.size(2), rpn_cls_score.size(3)
        B, A = gt_boxes.size(0), self._num_anchors
        dev = gt_boxes.device
        if self._anchors.device != dev:
            self._anchors = self._anchors.to(dev)
        if keys is None:
            keys = torch.rand(B, H * W * A, device=dev, generator=self.generator)
        labels, targets, inside, outside = be.anchor_target(
            gt_boxes, im_info, self._anchors, keys, A, H, W, self._feat_stride, cfg.TRAIN.RPN_POSITIVE_OVERLAP,
            cfg.TRAIN.RPN_NEGATIVE_OVERLAP, cfg.TRAIN.RPN_CLOBBER_POSITIVES, cfg.TRAIN.RPN_FG_FRACTION,
            cfg.TRAIN.RPN_BATCHSIZE, cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0], cfg.TRAIN.RPN_POSITIVE_WEIGHT)
        return [labels, targets, inside, outside]
