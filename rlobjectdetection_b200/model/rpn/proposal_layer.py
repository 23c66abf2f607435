"""RPN proposal layer (reference: lib/model/rpn/proposal_layer.py:26-161).

Same constructor and call convention -- `_ProposalLayer(feat_stride, scales, ratios)` called
with `(rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key)` -- but the whole forward (anchor grid,
decode, clip, descending sort, pre-NMS top-k, NMS, post-NMS top-k, zero padding) is three
sm_100a launches for the whole batch with no host synchronisation, instead of a Python loop
over images around ~20 eager ops and a host-scanned NMS."""
import numpy as np
import torch
import torch.nn as nn

from .. import _backend as be
from ..utils.config import cfg
from .generate_anchors import generate_anchors


class _ProposalLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super().__init__()
        self._feat_stride = int(feat_stride)
        table = generate_anchors(scales=np.array(scales), ratios=np.array(ratios))
        self._anchors = torch.from_numpy(table).float()
        self._num_anchors = self._anchors.size(0)

    def forward(self, input):
        rpn_cls_prob, rpn_bbox_pred, im_info, cfg_key = input[0], input[1], input[2], input[3]
        params = cfg[cfg_key]
        if self._anchors.device != rpn_cls_prob.device:
            self._anchors = self._anchors.to(rpn_cls_prob.device)
        # RPN_MIN_SIZE is read but never applied by the reference (:75, :113); same here
        return be.proposal_forward(rpn_cls_prob, rpn_bbox_pred, im_info, self._anchors,
                                   self._feat_stride, params.RPN_PRE_NMS_TOP_N,
                                   params.RPN_POST_NMS_TOP_N, params.RPN_NMS_THRESH)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients (reference :163-165)."""

    def reshape(self, bottom, top):
        """Shapes are decided in forward (reference :167-169)."""
