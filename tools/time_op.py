#!/usr/bin/env python
"""Median cold-L2 time of one op: python tools/time_op.py {align_fwd|align_bwd|pool_fwd|pool_bwd|crop_fwd|crop_bwd} {C2|C4} [iters]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402

op, cfg = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "C4")
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda", 0)
# P2: an FPN P2 level of 800 x 1216 (stride 4, rois of 16-112 pixels); BIG / BIGD: a stride-16 map of 1600 x 2400 with 300 / 2000 rois per image
B, C, H, W, n_per = {"C2": (4, 1024, 38, 63, 256), "C4": (24, 1024, 50, 75, 300), "P2": (2, 256, 200, 304, 512),
                     "BIG": (2, 1024, 100, 150, 300), "BIGD": (2, 1024, 100, 150, 2000)}[cfg]
stride = 4.0 if cfg == "P2" else 16.0
g = torch.Generator().manual_seed(1)
feat = torch.randn(B, C, H, W, generator=g).to(dev)
if cfg == "P2":
    rois = torch.cat([torch.cat([torch.full((n_per, 1), float(b)), syn.random_boxes(g, n_per, H * stride, W * stride, 16.0, 112.0)], 1)
                      for b in range(B)]).to(dev)
else:
    rois = syn.rois_for_batch(2, B, n_per, H * stride, W * stride).to(dev)
gout = torch.randn(rois.size(0), C, 7, 7, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if op == "pool_bwd":
    _, am = be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0)
if op.startswith("crop"):
    grid = be.affine_grid(rois, (H, W), 14)
    gyx = torch.stack([grid[..., 1], grid[..., 0]], 3).contiguous()
    gout14 = torch.randn(rois.size(0), C, 14, 14, generator=g).to(dev)
fns = {
    "crop_fwd": lambda: be.roi_crop_forward(feat, gyx),
    "crop_bwd": lambda: be.roi_crop_backward(gout14, gyx, (B, C, H, W)),
    "align_fwd": lambda: be.roi_align_forward(feat, rois, 7, 7, 1 / stride, be.POOL_AVG),
    "align_bwd": lambda: be.roi_align_backward(gout, rois, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG),
    "pool_fwd": lambda: be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0),
    "pool_bwd": lambda: be.roi_pool_backward(gout, am, rois, (B, C, H, W), 7, 7, 1 / 16.0),
}
fn = fns[op]
for _ in range(3):
    fn()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
alg = 4 * (B * C * H * W + 5 * rois.size(0) + rois.size(0) * C * 49) * (2 if op.startswith("pool") else 1) - (4 * B * C * H * W if op.startswith("pool") else 0)
print(f"{op} {cfg} env={' '.join(k + '=' + v for k, v in sorted(os.environ.items()) if k.startswith('RLOD_'))}: median {statistics.median(ts):.1f} us  min {min(ts):.1f} us (whole call incl. plan launches)")
