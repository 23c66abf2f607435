"""Generate tests/golden/reference_rl.npz: one collated RL batch produced by the reference's
OWN COCODataLoader._collate_fn (lib/datasets/RL_coco_loader.py, imported unmodified) over
per-image (bboxes, labels) built by the label loop of COCODataset.__getitem__
(lib/datasets/RL_coco_dataset.py:107-145, transcribed with its line numbers: the class itself
needs COCO files) with IoU = the reference's vendored maskApi.c bbIou (oracle/_ref/libmaskapi.so)
and Action from lib/model/Reinforcement/action.py.

    python tests/golden/make_golden_rl.py      # authoring container only
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_stubs, orc  # noqa: E402


def getitem(dt_boxes, gt_boxes, cat_ids, action, pos_wratio, neg_wratio, img_id):
    """RL_coco_dataset.py:107-145 for one image; dt_boxes / gt_boxes: {cat_id: [dict, ...]}."""
    IoU = orc.ref_bbiou
    generate_bboxes, generate_labels = [], []
    for cat_id in cat_ids:                                               # :107
        for dt_box in dt_boxes.get(cat_id, []):                          # :108
            bbox = list(dt_box["bbox"])                                  # :110 (copy: the in-place bug :142 is not replicated)
            w, h = bbox[2], bbox[3]                                      # :111
            gtboxes = [g["bbox"] for g in gt_boxes.get(cat_id, [])]      # :113
            iscrowd = [int(g["iscrowd"]) for g in gt_boxes.get(cat_id, [])]
            if len(gtboxes) == 0:                                        # :115
                gtboxes, iscrowd = [[0, 0, 0, 0]], [0]
            origin_ious = IoU([bbox], gtboxes, iscrowd)                  # :119
            generate_label = []
            for act_id, act_delta in enumerate(action.actDeltas):       # :123
                new_bbox = bbox + act_delta * np.array([w, h, w, h])     # :124
                new_ious = IoU([new_bbox], gtboxes, iscrowd)             # :125
                delta_iou = new_ious.max() - origin_ious.max()           # :126
                if delta_iou > action.iou_thres:                         # :128
                    label, weight = 1, action.wtrans(delta_iou) * pos_wratio
                else:
                    label, weight = -1, action.wtrans(delta_iou) * neg_wratio
                generate_label.append([act_id, label, weight])           # :137
            score = dt_box["score"]
            bbox[2] += bbox[0]                                           # :142
            bbox[3] += bbox[1]                                           # :143
            generate_bboxes.append(bbox + [score] + [cat_id] + [img_id])  # :144
            generate_labels.append(generate_label)                       # :145
    return np.array(generate_bboxes), np.array(generate_labels)


def main():
    install_stubs()
    from math import exp, fabs
    from model.Reinforcement.action import Action
    sys.path.insert(0, os.path.join(os.environ.get("RLOD_REFERENCE", "/root/reference"), "lib", "datasets"))
    from RL_coco_loader import COCODataLoader

    action = Action([0.5, 0.25], wtrans=lambda x: exp(fabs(x)))           # config.py:48-51
    rng = np.random.RandomState(5)
    cat_ids = [1, 3, 7]
    B, pos_wratio, neg_wratio = 3, 2.5, 0.7
    batch, raw = [], []
    for b in range(B):
        dts, gts = {}, {}
        for c in cat_ids:
            nd = int(rng.randint(0, 6))
            ng = int(rng.randint(0, 4)) if not (b == 1 and c == 3) else 0
            dts[c] = [{"bbox": [float(v) for v in np.r_[rng.rand(2) * 300, rng.rand(2) * 150 + 5].astype(np.float32)],
                       "score": float(np.float32(rng.rand()))} for _ in range(nd)]
            gts[c] = [{"bbox": [float(v) for v in np.r_[rng.rand(2) * 300, rng.rand(2) * 150 + 5].astype(np.float32)],
                       "iscrowd": int(rng.rand() < 0.25)} for _ in range(ng)]
        if b == 0 and not dts[1]:
            dts[1] = [{"bbox": [10.0, 20.0, 50.0, 60.0], "score": 0.5}]
        bb, ll = getitem(dts, gts, cat_ids, action, pos_wratio, neg_wratio, 100 + b)
        img = torch.zeros(3, 8 + b, 9)
        batch.append((img, torch.from_numpy(bb).float(), torch.from_numpy(ll).float(), (8 + b, 9, 1.0)))
        raw.append((dts, gts))
    _, padded_bboxes, padded_labels, _ = COCODataLoader._collate_fn(None, batch)   # RL_coco_loader.py:19-76
    N = padded_bboxes.shape[1]
    G = max(sum(len(g[c]) for c in cat_ids) for _, g in raw)
    dets = np.zeros((B, N, 4), np.float32); det_cat = np.zeros((B, N), np.int32); det_score = np.zeros((B, N), np.float32)
    det_img = np.zeros((B, N), np.float32); ndet = np.zeros(B, np.int32)
    gt = np.zeros((B, max(G, 1), 4), np.float32); gt_cat = np.full((B, max(G, 1)), -1, np.int32)
    crowd = np.zeros((B, max(G, 1)), np.uint8); ngt = np.zeros(B, np.int32)
    for b, (dts, gts) in enumerate(raw):
        i = 0
        for c in cat_ids:
            for d in dts[c]:
                dets[b, i], det_cat[b, i], det_score[b, i], det_img[b, i] = d["bbox"], c, d["score"], 100 + b
                i += 1
        ndet[b] = i
        j = 0
        for c in cat_ids:
            for g in gts[c]:
                gt[b, j], gt_cat[b, j], crowd[b, j] = g["bbox"], c, g["iscrowd"]
                j += 1
        ngt[b] = j
    np.savez_compressed(os.path.join(HERE, "reference_rl.npz"), dets=dets, det_cat=det_cat, det_score=det_score,
                        det_img=det_img, ndet=ndet, gt=gt, gt_cat=gt_cat, crowd=crowd, ngt=ngt,
                        act=action.actDeltas.astype(np.float32), wratio=np.array([pos_wratio, neg_wratio]),
                        padded_bboxes=padded_bboxes.numpy(), padded_labels=padded_labels.numpy())
    print("wrote reference_rl.npz: bboxes", tuple(padded_bboxes.shape), "labels", tuple(padded_labels.shape), "ndet", ndet)


if __name__ == "__main__":
    main()
