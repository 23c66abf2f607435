"""Reward / label / weight generation of the RL refinement step on the device.

Replaces the per-box x per-action loop inside COCODataset.__getitem__
(lib/datasets/RL_coco_dataset.py:119-137: one pycocotools bbIou call per (box, action) in a
DataLoader worker) with one launch over the whole batch (rlod_action_reward)."""
from .. import _backend as be

IOU_COCO, IOU_RCNN = be.IOU_COCO, be.IOU_RCNN


def action_rewards(action, boxes, gt_boxes, iscrowd=None, num_gt=None, mode=IOU_COCO,
                   pos_wratio=1.0, neg_wratio=1.0):
    """boxes (B,N,4), gt_boxes (B,G,4) -> reward, label, weight, each (B,N,num_acts).
    mode IOU_COCO: xywh + pycocotools semantics (crowd, fp64); IOU_RCNN: x1y1x2y2 + bbox_overlaps.
    weight = exp(|reward|) * (pos|neg)_wratio, the reference's Config.act_wtrans
    (config.py:48-51)."""
    return be.action_reward(boxes, gt_boxes, action.table(boxes.device), crowd=iscrowd, ngt=num_gt,
                            mode=mode, iou_thres=float(action.iou_thres), pos_wratio=pos_wratio,
                            neg_wratio=neg_wratio)
