"""RoICropFunction (lib/model/roi_crop/functions/roi_crop.py:7-21) as a new-style autograd
Function over rlod_roi_crop_forward / _backward.  input1 = features (B,C,H,W), input2 = grid
(R,gh,gw,2) holding (y, x) in [-1,1]; roi r samples image r // (R // B), as the reference's
kernel does.  Like the reference, the grid receives a zero gradient."""
import torch
from torch.autograd import Function

from ... import _backend as be


class RoICropFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        ctx.save_for_backward(input2)
        ctx.feature_size = tuple(input1.shape)
        return be.roi_crop_forward(input1, input2)

    @staticmethod
    def backward(ctx, grad_output):
        (grid,) = ctx.saved_tensors
        grad_input1 = be.roi_crop_backward(grad_output.contiguous(), grid, ctx.feature_size)
        return grad_input1, torch.zeros_like(grid)
