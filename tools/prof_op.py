#!/usr/bin/env python
"""Run one op of the path a few times (for ncu): python tools/prof_op.py {align_fwd|align_bwd|pool_fwd|pool_bwd|nms|proposal} {C2|C4|BIGD}"""  # BIGD: 2 x 1024 x 100 x 150, 2000 rois per image (tiled planes)
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402

op, cfg = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "C4")
dev = torch.device("cuda", 0)
B, C, H, W, n_per = {"C2": (4, 1024, 38, 63, 256), "BIGD": (2, 1024, 100, 150, 2000)}.get(cfg, (24, 1024, 50, 75, 300))
g = torch.Generator().manual_seed(1)
if op.startswith("align") or op.startswith("pool"):
    feat = torch.randn(B, C, H, W, generator=g).to(dev)
    rois = syn.rois_for_batch(2, B, n_per, H * 16.0, W * 16.0).to(dev)
    gout = torch.randn(rois.size(0), C, 7, 7, generator=g).to(dev)
    for _ in range(3):
        if op == "align_fwd":
            be.roi_align_forward(feat, rois, 7, 7, 1 / 16.0, be.POOL_AVG)
        elif op == "align_bwd":
            be.roi_align_backward(gout, rois, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG)
        elif op == "pool_fwd":
            out, am = be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0)
        elif op == "pool_bwd":
            out, am = be.roi_pool_forward(feat, rois, 7, 7, 1 / 16.0)
            be.roi_pool_backward(gout, am, rois, (B, C, H, W), 7, 7, 1 / 16.0)
elif op == "nms":
    n = 12000
    bx = syn.random_boxes(g, n, 600, 1000, 8.0, 300.0)
    sc = syn.distinct_scores(g, (n,)).sort(descending=True).values
    d = torch.cat([bx, sc[:, None]], 1).contiguous().to(dev)
    for _ in range(3):
        be.nms_padded(d, 0.7)
elif op == "proposal":
    import numpy as np
    from rlobjectdetection_b200.model.rpn.generate_anchors import generate_anchors
    Bp, A, Hh, Ww, imh, imw, pre, post, scales = ((1, 9, 37, 62, 600, 1000, 12000, 2000, (8, 16, 32)) if cfg == "C1"
                                                  else (24, 12, 50, 75, 800, 1200, 6000, 300, (4, 8, 16, 32)))
    scores, deltas, im_info = syn.rpn_outputs(5, Bp, A, Hh, Ww, imh, imw, imh / 600.0)
    anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array([0.5, 1, 2])).astype(np.float32)).to(dev)
    for _ in range(3):
        be.proposal_forward(scores.to(dev), deltas.to(dev), im_info.to(dev), anchors, 16, pre, post, 0.7)
torch.cuda.synchronize()
print("ok", op, cfg)
