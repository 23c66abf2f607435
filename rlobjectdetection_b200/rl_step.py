"""The RL refinement step either side of the reward kernel, on the device.

* generate_labels: what COCODataset.__getitem__ + COCODataLoader._collate_fn compute on the host
  for every batch (lib/datasets/RL_coco_dataset.py:107-145, RL_coco_loader.py:19-76): the padded
  detection tensor [b, N, 8] = (bid, x1, y1, x2, y2, score, cat_id, img_id) and the label tensor
  [b, N, A, 3] = (act_id, label, weight) -- here from device tensors in one launch, without the
  reference's in-place corruption of the cached detections (:142-143, not replicated).
* eval_step: the box update of the evaluation loop (trainval_net.py:202-237): x1y1x2y2 -> xywh,
  Action.move_from_act(maxk), back to image scale.
"""
import torch

from .model import _backend as be
from .model.Reinforcement.action import apply_custom_wtrans, wtrans_code


def generate_labels(action, dets_xywh, det_score, det_cat, det_img, ndet, gt_xywh, gt_cat, iscrowd=None, ngt=None,
                    pos_wratio=1.0, neg_wratio=1.0):
    """dets_xywh (B,N,4) COCO xywh boxes, det_score / det_cat / det_img (B,N), ndet (B) valid rows per
    image; gt_xywh (B,G,4), gt_cat (B,G), iscrowd (B,G), ngt (B).  Returns (bboxes [B,N,8],
    labels [B,N,A,3]) with the collate's zero padding."""
    B, N, _ = dets_xywh.shape
    code = wtrans_code(action)
    labels = be.rl_labels(dets_xywh, gt_xywh, action.table(dets_xywh.device), det_cat=det_cat, ndet=ndet,
                          gt_cat=gt_cat, crowd=iscrowd, ngt=ngt, iou_thres=float(action.iou_thres),
                          pos_wratio=pos_wratio, neg_wratio=neg_wratio, wtrans=code)
    if code == be.WTRANS_RAW:  # a callable the kernel does not know: applied to the raw delta_iou
        w = apply_custom_wtrans(action, labels[..., 2], labels[..., 1], pos_wratio, neg_wratio)
        labels[..., 2] = torch.where(labels[..., 1] != 0, w, torch.zeros_like(w))  # padded rows stay zero
    d = dets_xywh.float()
    x2, y2 = d[..., 0] + d[..., 2], d[..., 1] + d[..., 3]  # bbox[2] += bbox[0]; bbox[3] += bbox[1] (:142-143)
    rows = torch.stack([d[..., 0], d[..., 1], x2, y2, det_score.float(), det_cat.float(), det_img.float()], 2)
    valid = (torch.arange(N, device=d.device)[None, :] < ndet.to(d.device)[:, None])[..., None]
    rows = torch.where(valid, rows, torch.zeros_like(rows))
    bid = torch.arange(B, device=d.device, dtype=torch.float32)[:, None, None].expand(B, N, 1)
    return torch.cat([bid, rows], 2).contiguous(), labels


def eval_step(action, bboxes8, preds, targets, im_scale, maxk=1):
    """trainval_net.py:202-215: bboxes8 [b,N,8] (bid,x1,y1,x2,y2,score,cat,img) on the device ->
    refined boxes in xywh at the ORIGINAL image scale (bbox[1:5] /= scale, :221) and the number of
    boxes moved.  preds / targets (b,N,A); im_scale (b,) resize scales."""
    out = bboxes8.clone().float()
    out[:, :, 3] = out[:, :, 3] - out[:, :, 1]   # :204
    out[:, :, 4] = out[:, :, 4] - out[:, :, 2]   # :205
    b, N, _ = out.shape
    moved = _move_xywh(action, out, preds, targets, maxk)
    out[:, :, 1:5] = out[:, :, 1:5] / im_scale.to(out.device).float()[:, None, None]
    return out, moved


def _move_xywh(action, out, preds, targets, maxk):
    # rlod_move_from_act works on rows of `box_stride` floats: columns 1..4 of the 8-wide rows
    b, N, _ = out.shape
    boxes = out[:, :, 1:5].contiguous()
    moved = be.move_from_act(boxes, preds, targets, action.table(out.device), maxk, corners=False)
    out[:, :, 1:5] = boxes
    return moved
