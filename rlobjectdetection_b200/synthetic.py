"""Seeded synthetic inputs of the BASELINE.json configs, shared by tests and
bench.py.  Everything is generated on the CPU with torch.Generator so that the CPU oracle and
the GPU path see identical bits; callers move tensors to the device."""
import math

import torch


def distinct_scores(gen, shape):
    """A random permutation of (1..n)/(n+1): all scores distinct -> tie-free ordering."""
    n = 1
    for s in shape:
        n *= s
    perm = torch.randperm(n, generator=gen).float() + 1.0
    return (perm / (n + 1)).reshape(shape)


def rpn_outputs(seed, B, A, H, W, im_h, im_w, im_scale=1.0, delta_std=0.2):
    """scores (B,2A,H,W) with distinct fg scores, deltas (B,4A,H,W), im_info (B,3)."""
    g = torch.Generator().manual_seed(seed)
    bg = torch.rand(B, A, H, W, generator=g)
    fg = torch.stack([distinct_scores(g, (A, H, W)) for _ in range(B)], 0)
    scores = torch.cat([bg, fg], 1).contiguous()
    deltas = torch.randn(B, 4 * A, H, W, generator=g) * delta_std
    im_info = torch.tensor([[float(im_h), float(im_w), float(im_scale)]] * B)
    return scores, deltas, im_info


def random_boxes(gen, n, im_h, im_w, smin=16.0, smax=512.0):
    """x1y1x2y2 boxes: centre uniform in the image, log-uniform side in [smin, smax], clipped."""
    cx = torch.rand(n, generator=gen) * (im_w - 1)
    cy = torch.rand(n, generator=gen) * (im_h - 1)
    w = torch.exp(torch.rand(n, generator=gen) * (math.log(smax) - math.log(smin)) + math.log(smin))
    h = torch.exp(torch.rand(n, generator=gen) * (math.log(smax) - math.log(smin)) + math.log(smin))
    x1 = (cx - 0.5 * w).clamp(0, im_w - 1)
    y1 = (cy - 0.5 * h).clamp(0, im_h - 1)
    x2 = (cx + 0.5 * w).clamp(0, im_w - 1)
    y2 = (cy + 0.5 * h).clamp(0, im_h - 1)
    return torch.stack([x1, y1, x2, y2], 1)


def rois_for_batch(seed, B, n_per_image, im_h, im_w, edge_cases=True):
    """(B*n, 5) rois grouped by image.  edge_cases replaces the first rows of image 0 with
    hand-made rois: border-touching, one-pixel, fully outside, inverted (x2 < x1)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        bx = random_boxes(g, n_per_image, im_h, im_w)
        rows.append(torch.cat([torch.full((n_per_image, 1), float(b)), bx], 1))
    rois = torch.cat(rows, 0)
    if edge_cases and n_per_image >= 8:
        e = torch.tensor([
            [0, 0, 0, im_w - 1, im_h - 1],                 # whole image
            [0, im_w - 40, im_h - 40, im_w - 1, im_h - 1],  # touches the far border
            [0, 17, 23, 17, 23],                            # one pixel
            [0, -200, -200, -50, -50],                      # fully outside (negative)
            [0, im_w + 50, im_h + 50, im_w + 300, im_h + 300],  # fully outside (beyond)
            [0, 300, 200, 100, 80],                         # inverted: x2 < x1, y2 < y1
            [0, -30, 40, 90, 700],                          # straddles top-left / bottom
            [0, 5.5, 7.25, 300.75, 11.125],                 # thin, fractional
        ], dtype=torch.float32)
        rois[:8] = e
    return rois.contiguous()


def gt_boxes(seed, B, G, im_h, im_w, crowd_frac=0.1):
    """gt (B,G,4) x1y1x2y2 and iscrowd (B,G) uint8."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.stack([random_boxes(g, G, im_h, im_w, 32.0, 400.0) for _ in range(B)], 0)
    crowd = (torch.rand(B, G, generator=g) < crowd_frac).to(torch.uint8)
    return gt.contiguous(), crowd


def to_xywh(boxes):
    """x1y1x2y2 -> COCO xywh (w = x2 - x1, no +1: the json convention)."""
    out = boxes.clone()
    out[..., 2] = boxes[..., 2] - boxes[..., 0]
    out[..., 3] = boxes[..., 3] - boxes[..., 1]
    return out


def clustered_dets(seed, n_images, n_classes, per_seg, im_h=600, im_w=1000, centres=30, jitter=8.0):
    """Per-class test-time NMS input (config 5): for every (image, class) segment `per_seg`
    boxes = one of `centres` cluster boxes + N(0, jitter) noise, distinct scores sorted
    descending inside the segment.  Returns dets (n_images*n_classes*per_seg, 5), seg_offsets."""
    g = torch.Generator().manual_seed(seed)
    segs = n_images * n_classes
    base = torch.stack([random_boxes(g, centres, im_h, im_w, 30.0, 300.0) for _ in range(n_images)], 0)
    pick = torch.randint(0, centres, (n_images, n_classes, per_seg), generator=g)
    boxes = torch.gather(base[:, None].expand(n_images, n_classes, centres, 4), 2,
                         pick[..., None].expand(n_images, n_classes, per_seg, 4))
    boxes = boxes + torch.randn(n_images, n_classes, per_seg, 4, generator=g) * jitter
    x1 = torch.minimum(boxes[..., 0], boxes[..., 2])
    x2 = torch.maximum(boxes[..., 0], boxes[..., 2])
    y1 = torch.minimum(boxes[..., 1], boxes[..., 3])
    y2 = torch.maximum(boxes[..., 1], boxes[..., 3])
    scores = torch.stack([distinct_scores(g, (per_seg,)).sort(descending=True).values
                          for _ in range(segs)], 0).reshape(n_images, n_classes, per_seg)
    dets = torch.stack([x1, y1, x2, y2, scores], -1).reshape(segs * per_seg, 5).contiguous()
    seg_offsets = torch.arange(0, segs + 1, dtype=torch.int32) * per_seg
    return dets, seg_offsets
