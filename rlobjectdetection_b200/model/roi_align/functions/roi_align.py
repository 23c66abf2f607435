"""RoIAlign autograd op (reference: lib/model/roi_align/functions/roi_align.py:7-47).

The reference is a legacy instance-style Function (removed from PyTorch); this is a static
`torch.autograd.Function`, and `RoIAlignFunction(ah, aw, scale)(features, rois)` keeps the
reference's call form.  pool_mode fuses the 2x2 stride-1 avg / max pooling that
RoIAlignAvg / RoIAlignMax apply afterwards (modules/roi_align.py:26-29, 39-42)."""
import torch
from torch.autograd import Function

from ... import _backend as be


class _RoIAlignOp(Function):
    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale, pool_mode):
        out = be.roi_align_forward(features, rois, ah, aw, scale, pool_mode)
        ctx.cfg = (int(ah), int(aw), float(scale), int(pool_mode), tuple(features.shape))
        if pool_mode == be.POOL_MAX:
            ctx.save_for_backward(rois, features)
        else:
            ctx.save_for_backward(rois)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        ah, aw, scale, pool_mode, fsize = ctx.cfg
        saved = ctx.saved_tensors
        rois = saved[0]
        feats = saved[1] if pool_mode == be.POOL_MAX else None
        grad_in = be.roi_align_backward(grad_output, rois, feats, fsize, ah, aw, scale, pool_mode)
        return grad_in, None, None, None, None, None


class RoIAlignFunction:
    """Callable with the reference's constructor: RoIAlignFunction(ah, aw, scale)(features, rois)
    -> (R, C, ah, aw) bilinear sample grid."""

    def __init__(self, aligned_height, aligned_width, spatial_scale, pool_mode=be.POOL_NONE):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.pool_mode = int(pool_mode)

    def __call__(self, features, rois):
        if not features.is_cuda:
            raise NotImplementedError  # reference :28-29
        return _RoIAlignOp.apply(features, rois, self.aligned_height, self.aligned_width,
                                 self.spatial_scale, self.pool_mode)
