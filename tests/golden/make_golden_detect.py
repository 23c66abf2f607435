"""Generate tests/golden/reference_detect.npz: the test-time post-processing loop of the
reference (RCNN_bases/test_net.py:244-307) EXECUTED with the reference's own bbox_transform_inv
and clip_boxes (imported unmodified from /root/reference/lib) and torch.sort / np.sort exactly as
written there; nms() is the reference wrapper over the stubbed cffi extension (oracle greedy NMS,
itself pinned to the reference's legacy CUDA kernel on the GPU).  The loop lives in a script in
the reference (not importable), so its lines are transcribed here with their line numbers.

    python tests/golden/make_golden_detect.py      # authoring container only
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_stubs  # noqa: E402


def reference_loop(rois, cls_prob, bbox_pred, im_info, num_classes, thresh, max_per_image, class_agnostic, cfg,
                   bbox_transform_inv, clip_boxes, nms):
    scores = cls_prob                                               # :244
    boxes = rois[:, :, 1:5]                                         # :245
    box_deltas = bbox_pred                                          # :249
    stds = torch.FloatTensor(cfg.TRAIN.BBOX_NORMALIZE_STDS)
    means = torch.FloatTensor(cfg.TRAIN.BBOX_NORMALIZE_MEANS)
    if class_agnostic:                                              # :252-255
        box_deltas = box_deltas.view(-1, 4) * stds + means
        box_deltas = box_deltas.view(1, -1, 4)
    else:                                                           # :256-259
        box_deltas = box_deltas.view(-1, 4) * stds + means
        box_deltas = box_deltas.view(1, -1, 4 * num_classes)
    pred_boxes = bbox_transform_inv(boxes, box_deltas, 1)           # :261
    pred_boxes = clip_boxes(pred_boxes, im_info, 1)                 # :262
    pred_boxes /= im_info[0][2]                                     # :267
    scores = scores.squeeze()                                       # :269
    pred_boxes = pred_boxes.squeeze()                               # :270
    all_boxes = [np.zeros((0, 5), np.float32) for _ in range(num_classes)]
    for j in range(1, num_classes):                                 # :277
        inds = torch.nonzero(scores[:, j] > thresh).view(-1)        # :278
        if inds.numel() > 0:                                        # :280
            cls_scores = scores[:, j][inds]                         # :281
            _, order = torch.sort(cls_scores, 0, True)              # :282
            if class_agnostic:
                cls_boxes = pred_boxes[inds, :]                     # :284
            else:
                cls_boxes = pred_boxes[inds][:, j * 4:(j + 1) * 4]  # :286
            cls_dets = torch.cat((cls_boxes, cls_scores.unsqueeze(1)), 1)  # :288
            cls_dets = cls_dets[order]                              # :290
            keep = nms(cls_dets, cfg.TEST.NMS)                      # :291
            cls_dets = cls_dets[keep.view(-1).long()]               # :292
            all_boxes[j] = cls_dets.cpu().numpy()                   # :295
    if max_per_image > 0:                                           # :300
        image_scores = np.hstack([all_boxes[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > max_per_image:                       # :303
            image_thresh = np.sort(image_scores)[-max_per_image]    # :304
            for j in range(1, num_classes):                         # :305-307
                keep = np.where(all_boxes[j][:, -1] >= image_thresh)[0]
                all_boxes[j] = all_boxes[j][keep, :]
    return all_boxes


def main():
    install_stubs()
    from model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    from model.nms.nms_wrapper import nms as ref_nms
    from model.utils.config import cfg

    def nms(dets, thresh):  # nms_wrapper.nms needs CUDA tensors only for its .is_cuda branch: call nms_gpu's cffi path
        from model.nms.nms_gpu import nms_gpu
        return nms_gpu(dets, thresh)

    out = {}
    for tag, (N, K, agn, thresh, cap, seed) in {"c21": (300, 21, False, 0.05, 100, 1), "c9": (97, 9, False, 0.0, 40, 2),
                                                "agn": (120, 6, True, 0.1, 30, 3)}.items():
        g = torch.Generator().manual_seed(50 + seed)
        im_info = torch.tensor([[600.0, 901.0, 1.5]])
        cx, cy = torch.rand(N, generator=g) * 900, torch.rand(N, generator=g) * 600
        w, h = torch.rand(N, generator=g) * 250 + 8, torch.rand(N, generator=g) * 250 + 8
        # clustered rois so that NMS has work: snap centres to a coarse lattice + jitter
        cx, cy = (cx / 120).round() * 120 + torch.randn(N, generator=g) * 6, (cy / 120).round() * 120 + torch.randn(N, generator=g) * 6
        b = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
        b[:, 0::2] = b[:, 0::2].clamp(0, 900)
        b[:, 1::2] = b[:, 1::2].clamp(0, 599)
        rois = torch.cat([torch.zeros(N, 1), b], 1)[None].contiguous()
        logits = torch.randn(N, K, generator=g) * 2.0
        cls_prob = torch.softmax(logits, 1)[None].contiguous()
        bbox_pred = (torch.randn(N, 4 if agn else 4 * K, generator=g) * 1.5)[None].contiguous()
        ab = reference_loop(rois, cls_prob, bbox_pred.clone(), im_info, K, thresh, cap, agn, cfg, bbox_transform_inv,
                            clip_boxes, nms)
        out.update({f"{tag}_rois": rois.numpy(), f"{tag}_cls_prob": cls_prob.numpy(), f"{tag}_bbox_pred": bbox_pred.numpy(),
                    f"{tag}_im_info": im_info.numpy(),
                    f"{tag}_cfg": np.array([N, K, int(agn), cap], np.int32), f"{tag}_thresh": np.array(thresh, np.float32),
                    f"{tag}_counts": np.array([len(a) for a in ab], np.int32),
                    f"{tag}_dets": np.concatenate(ab, 0).astype(np.float32)})
        print(tag, "kept per class", [len(a) for a in ab])
    np.savez_compressed(os.path.join(HERE, "reference_detect.npz"), **out)
    print("wrote reference_detect.npz,", os.path.getsize(os.path.join(HERE, "reference_detect.npz")), "bytes")


if __name__ == "__main__":
    main()
