#!/usr/bin/env python
"""Two RoIAlignAvg forward calls on the same feature map against ONE call on the concatenated roi sets
(python tools/time_merge.py {C4|C4xN}): what sharing the plane fill and the plan launches is worth."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
dev = torch.device("cuda", 0)
# C4xN: N images of the C4 shape (what a rank holds when the batch of 24 is sharded: C4x3 at 8 GPUs, C4x6 at 4)
B, C, H, W, n_per = (24, 1024, 50, 75, 300) if cfg == "C4" else (int(cfg[3:]), 1024, 50, 75, 300)
g = torch.Generator().manual_seed(1)
feat = torch.randn(B, C, H, W, generator=g).to(dev)
ra = syn.rois_for_batch(2, B, n_per, H * 16.0, W * 16.0, edge_cases=False).to(dev)
rb = syn.rois_for_batch(3, B, n_per, H * 16.0, W * 16.0, edge_cases=False).to(dev)
cat = torch.cat([ra, rb]).contiguous()                                           # [set A; set B]: not grouped by image
inter = torch.cat([ra.view(B, n_per, 5), rb.view(B, n_per, 5)], 1).reshape(-1, 5).contiguous()  # grouped by image
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


two = timeit(lambda: (be.roi_align_forward(feat, ra, 7, 7, 1 / 16.0, be.POOL_AVG), be.roi_align_forward(feat, rb, 7, 7, 1 / 16.0, be.POOL_AVG)))
one_cat = timeit(lambda: be.roi_align_forward(feat, cat, 7, 7, 1 / 16.0, be.POOL_AVG))
one_int = timeit(lambda: be.roi_align_forward(feat, inter, 7, 7, 1 / 16.0, be.POOL_AVG))
o2 = torch.cat([be.roi_align_forward(feat, ra, 7, 7, 1 / 16.0, be.POOL_AVG), be.roi_align_forward(feat, rb, 7, 7, 1 / 16.0, be.POOL_AVG)])
o1 = be.roi_align_forward(feat, cat, 7, 7, 1 / 16.0, be.POOL_AVG)
print(f"{cfg}: two calls {two:.1f} us; one call [A;B] {one_cat:.1f} us; one call interleaved per image {one_int:.1f} us; "
      f"same bits: {bool(torch.equal(o1, o2))}")
