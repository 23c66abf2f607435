"""RoIAlign / RoIAlignAvg / RoIAlignMax modules (reference:
lib/model/roi_align/modules/roi_align.py:6-42), same constructors.  Avg / Max sample an
(ah+1) x (aw+1) grid and pool it 2x2 stride 1 -- here inside the same kernel, so the
(R, C, ah+1, aw+1) intermediate (268 MB at 1024 rois x 1024 channels) never exists."""
from torch.nn.modules.module import Module

from ... import _backend as be
from ..functions.roi_align import RoIAlignFunction


class _RoIAlignBase(Module):
    pool_mode = be.POOL_NONE

    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super().__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale,
                                self.pool_mode)(features, rois)

    # Inference-only extension of the reference's interface: the forward in two calls, so that the rois can be
    # planned early (on another stream) and only the pooling kernel sits on the caller's critical stream.
    def plan(self, rois, feature_size):
        """-> a plan of `rois` for a (B, C, H, W) = feature_size map (rlod_roi_align_plan, current stream)."""
        return be.roi_align_plan(rois, feature_size, self.aligned_height, self.aligned_width, self.spatial_scale,
                                 self.pool_mode)

    def forward_planned(self, features, plan):
        """forward(features, the planned rois) without autograd; order it after the plan (same stream or an event)."""
        return be.roi_align_forward_planned(features, plan)


class RoIAlign(_RoIAlignBase):
    pool_mode = be.POOL_NONE


class RoIAlignAvg(_RoIAlignBase):
    pool_mode = be.POOL_AVG


class RoIAlignMax(_RoIAlignBase):
    pool_mode = be.POOL_MAX
