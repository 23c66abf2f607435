"""Box algebra on the device (reference: lib/model/rpn/bbox_transform.py).  Same names and
argument order; each function is ONE sm_100a launch (rlod_bbox_transform_inv, rlod_clip_boxes,
rlod_bbox_overlaps, rlod_bbox_overlaps_batch) instead of ~15 / 4*B / ~20 eager ops."""
import torch

from .. import _backend as be


def bbox_transform_inv(boxes, deltas, batch_size):
    """boxes (B,N,4) + deltas (B,N,4k) -> (B,N,4k)   [reference :77-103]."""
    be.require_cuda("bbox_transform_inv", boxes, deltas)
    boxes, deltas = be.f32c(boxes), be.f32c(deltas)
    B, N = boxes.shape[0], boxes.shape[1]
    k = deltas.shape[2] // 4
    out = torch.empty_like(deltas)
    with torch.cuda.device(boxes.device):
        be.check(be.lib().rlod_bbox_transform_inv(be.ptr(boxes), be.ptr(deltas), B * N, k,
                                                  be.ptr(out), be.stream_of(boxes)),
                 "rlod_bbox_transform_inv")
    return out


def clip_boxes(boxes, im_shape, batch_size):
    """In-place clamp of (B,N,4k) boxes to [0, w-1] x [0, h-1]; im_shape (B,3) = [h, w, scale]
    [reference :125-133].  Returns `boxes`."""
    be.require_cuda("clip_boxes", boxes, im_shape)
    if not (boxes.is_contiguous() and boxes.dtype == torch.float32):
        raise ValueError("clip_boxes works in place: boxes must be contiguous fp32")
    im_shape = be.f32c(im_shape)
    B, N = boxes.shape[0], boxes.shape[1]
    k = boxes.shape[2] // 4
    with torch.cuda.device(boxes.device):
        be.check(be.lib().rlod_clip_boxes(be.ptr(boxes), be.ptr(im_shape), B, N, k,
                                          be.stream_of(boxes)), "rlod_clip_boxes")
    return boxes


def bbox_overlaps(anchors, gt_boxes):
    """(N,4) x (K,4) -> (N,K) IoU, +1 convention [reference :136-166]."""
    be.require_cuda("bbox_overlaps", anchors, gt_boxes)
    anchors, gt_boxes = be.f32c(anchors), be.f32c(gt_boxes)
    N, K = anchors.size(0), gt_boxes.size(0)
    out = torch.empty(N, K, dtype=torch.float32, device=anchors.device)
    with torch.cuda.device(anchors.device):
        be.check(be.lib().rlod_bbox_overlaps(be.ptr(anchors), be.ptr(gt_boxes), N, K, be.ptr(out),
                                             be.stream_of(anchors)), "rlod_bbox_overlaps")
    return out


def bbox_overlaps_batch(anchors, gt_boxes):
    """anchors (N,4) | (B,N,4) | (B,N,5: columns 1:5); gt_boxes (B,K,>=4) -> (B,N,K) with the
    reference's sentinels: zero-area gt -> 0, zero-area anchor -> -1 [reference :168-257]."""
    be.require_cuda("bbox_overlaps_batch", anchors, gt_boxes)
    anchors, gt_boxes = be.f32c(anchors), be.f32c(gt_boxes)
    B, K, gs = gt_boxes.shape
    if anchors.dim() == 2:
        N, row, bstride, skip = anchors.size(0), anchors.size(1), 0, 0
    elif anchors.dim() == 3:
        N, row = anchors.size(1), anchors.size(2)
        bstride, skip = N * row, (0 if row == 4 else 1)
    else:
        raise ValueError("anchors input dimension is not correct.")
    out = torch.empty(B, N, K, dtype=torch.float32, device=anchors.device)
    import ctypes
    a_ptr = ctypes.c_void_p(anchors.data_ptr() + 4 * skip)
    with torch.cuda.device(anchors.device):
        be.check(be.lib().rlod_bbox_overlaps_batch(a_ptr, bstride, row, be.ptr(gt_boxes), gs, B, N,
                                                   K, be.ptr(out), be.stream_of(anchors)),
                 "rlod_bbox_overlaps_batch")
    return out
