"""Test-time post-processing of the detector head on the device: the per-class threshold /
box-regression / clip / sort / NMS loop and the max_per_image cap of the reference's
RCNN_bases/test_net.py:244-307 (also demo.py:305-334) for a whole batch in two launches.

    dets, counts = postprocess_detections(rois, cls_prob, bbox_pred, im_info, thresh=0.0)
    all_boxes = to_all_boxes(dets, counts)        # all_boxes[j][i] = (k, 5) numpy, as the reference

cfg defaults follow the reference: TEST.NMS = 0.3 (lib/model/utils/config.py:175),
max_per_image = 100 (test_net.py:59), BBOX_NORMALIZE_STDS / MEANS (config.py:90-95).
"""
import numpy as np

from .model import _backend as be
from .model.utils.config import cfg


def postprocess_detections(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=None, max_per_image=100,
                           class_agnostic=False, bbox_reg=None, normalize=None):
    """rois (B,N,5), cls_prob (B,N,K), bbox_pred (B,N,4K) [(B,N,4) class-agnostic], im_info (B,3),
    all CUDA.  Returns dets (B,K,N,5) and counts (B,K) on the device, no host synchronisation."""
    nms_thresh = cfg.TEST.NMS if nms_thresh is None else nms_thresh
    bbox_reg = cfg.TEST.BBOX_REG if bbox_reg is None else bbox_reg
    normalize = cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED if normalize is None else normalize
    stds = cfg.TRAIN.BBOX_NORMALIZE_STDS if (bbox_reg and normalize) else None
    means = cfg.TRAIN.BBOX_NORMALIZE_MEANS if (bbox_reg and normalize) else None
    return be.detect_postprocess(rois, cls_prob, bbox_pred if bbox_reg else None, im_info, thresh, nms_thresh,
                                 max_per_image, stds, means, class_agnostic)


def to_all_boxes(dets, counts):
    """(B,K,N,5) + (B,K) -> all_boxes[j][i] numpy arrays, the reference's layout (test_net.py:217-219).
    One device->host copy for the whole batch (the reference does one per class and image)."""
    d, c = dets.cpu().numpy(), counts.cpu().numpy()
    B, K = c.shape
    return [[np.ascontiguousarray(d[i, j, :c[i, j]]) for i in range(B)] for j in range(K)]
