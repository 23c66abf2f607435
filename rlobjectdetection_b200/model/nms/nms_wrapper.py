"""nms(dets, thresh, force_cpu=False) (reference: lib/model/nms/nms_wrapper.py:11-18)."""
from .. import _backend as be
from .nms_gpu import nms_gpu

nms_padded = be.nms_padded    # (keep padded with -1, count) without a host sync
nms_batched = be.nms_batched  # many segments (e.g. 80 classes x images) in one launch


def nms(dets, thresh, force_cpu=False):
    """dets (n,5) = [x1,y1,x2,y2,score] sorted by score descending.  `force_cpu` is accepted
    and ignored, exactly like the reference; an empty input returns [] (:13-14)."""
    if dets.shape[0] == 0:
        return []
    return nms_gpu(dets, thresh)
