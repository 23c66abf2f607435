"""Reward / label / weight generation of the RL refinement step on the device.

Replaces the per-box x per-action loop inside COCODataset.__getitem__
(lib/datasets/RL_coco_dataset.py:119-137: one pycocotools bbIou call per (box, action) in a
DataLoader worker) with one launch over the whole batch (rlod_action_reward)."""
from .. import _backend as be
from .action import apply_custom_wtrans, wtrans_code

IOU_COCO, IOU_RCNN = be.IOU_COCO, be.IOU_RCNN


def action_rewards(action, boxes, gt_boxes, iscrowd=None, num_gt=None, mode=IOU_COCO,
                   pos_wratio=1.0, neg_wratio=1.0):
    """boxes (B,N,4), gt_boxes (B,G,4) -> reward, label, weight, each (B,N,num_acts).
    mode IOU_COCO: xywh + pycocotools semantics (crowd, fp64); IOU_RCNN: x1y1x2y2 + bbox_overlaps.
    weight = action.wtrans(reward) * (pos|neg)_wratio (RL_coco_dataset.py:128-135): Identify
    (Action's default) and exp(|x|) (Config.act_wtrans, config.py:48-51) run inside the kernel,
    any other callable is applied to the raw rewards.  float64 boxes stay fp64 in COCO mode."""
    code = wtrans_code(action)
    reward, label, weight = be.action_reward(boxes, gt_boxes, action.table(boxes.device), crowd=iscrowd,
                                             ngt=num_gt, mode=mode, iou_thres=float(action.iou_thres),
                                             pos_wratio=pos_wratio, neg_wratio=neg_wratio, wtrans=code)
    if code == be.WTRANS_RAW:
        weight = apply_custom_wtrans(action, weight, label, pos_wratio, neg_wratio)
    return reward, label, weight
