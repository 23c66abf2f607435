"""RoIPool autograd op (reference: lib/model/roi_pooling/functions/roi_pool.py:6-38)."""
from torch.autograd import Function

from ... import _backend as be


class _RoIPoolOp(Function):
    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale):
        out, argmax = be.roi_pool_forward(features, rois, ph, pw, scale)
        ctx.cfg = (int(ph), int(pw), float(scale), tuple(features.shape))
        ctx.save_for_backward(argmax, rois)
        ctx.mark_non_differentiable(argmax)
        return out, argmax

    @staticmethod
    def backward(ctx, grad_output, _grad_argmax):
        ph, pw, scale, fsize = ctx.cfg
        argmax, rois = ctx.saved_tensors
        return be.roi_pool_backward(grad_output, argmax, rois, fsize, ph, pw, scale), None, None, None, None


class RoIPoolFunction:
    """RoIPoolFunction(ph, pw, scale)(features, rois) -> (R, C, ph, pw); `.argmax` holds the
    flat NCHW argmax of the last call like the reference's ctx.argmax."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.argmax = None

    def __call__(self, features, rois):
        if not features.is_cuda:
            # the reference's CPU branch reads NCHW memory with NHWC indexing (a latent bug,
            # functions/roi_pool.py:20-23); there is deliberately no CPU path here
            raise NotImplementedError
        import torch
        if not (torch.is_grad_enabled() and features.requires_grad):
            # inference (test_net.py runs under volatile / no_grad): no backward will ask for the argmax, so the
            # kernel tracks maxima only (2x faster); `.argmax` is None then
            out, self.argmax = be.roi_pool_forward(features, rois, self.pooled_height, self.pooled_width,
                                                   self.spatial_scale, want_argmax=False)
            return out
        out, self.argmax = _RoIPoolOp.apply(features, rois, self.pooled_height, self.pooled_width,
                                            self.spatial_scale)
        return out
