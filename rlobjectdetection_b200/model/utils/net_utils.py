"""The two helpers of lib/model/utils/net_utils.py that sit on the pooling path.

_affine_grid_gen (:143-165) builds the sampling grid of POOLING_MODE 'crop' from the rois; the
reference calls torch.nn.functional.affine_grid, whose base grid was linspace(-1, 1, g) in the
PyTorch 0.x it targets (today's align_corners=True) -- that is the default here; pass
align_corners=False for what an unmodified reference would compute under torch >= 1.3."""
import torch

from .. import _backend as be


def _affine_grid_gen(rois, input_size, grid_size, align_corners=True):
    """rois (R,5) [img,x1,y1,x2,y2] -> grid (R, grid_size, grid_size, 2) = (x, y) in [-1,1]."""
    return be.affine_grid(rois.detach(), input_size, grid_size, align_corners)


def _affine_theta(rois, input_size):
    """:167-199 (the (y, x)-ordered theta; used by the reference's STN variant only)."""
    rois = rois.detach().float()
    x1, y1, x2, y2 = (rois[:, i:i + 1] / 16.0 for i in (1, 2, 3, 4))
    height, width = input_size[0], input_size[1]
    zero = torch.zeros_like(x1)
    return torch.cat([(y2 - y1) / (height - 1), zero, (y1 + y2 - height + 1) / (height - 1),
                      zero, (x2 - x1) / (width - 1), (x1 + x2 - width + 1) / (width - 1)], 1).view(-1, 2, 3)


def crop_pool(base_feat, rois, pooling_size=7, max_pool=True, align_corners=True):
    """The 'crop' branch of _fasterRCNN.forward (lib/model/faster_rcnn/faster_rcnn.py:72-79):
    grid of size 2*POOLING_SIZE, (x, y) -> (y, x), _RoICrop, 2x2 max pool."""
    from ..roi_crop.modules.roi_crop import _RoICrop
    g = pooling_size * 2 if max_pool else pooling_size
    grid_xy = _affine_grid_gen(rois.view(-1, 5), base_feat.shape[2:], g, align_corners)
    grid_yx = torch.stack([grid_xy[:, :, :, 1], grid_xy[:, :, :, 0]], 3).contiguous()
    pooled = _RoICrop()(base_feat, grid_yx.detach())
    return torch.nn.functional.max_pool2d(pooled, 2, 2) if max_pool else pooled
