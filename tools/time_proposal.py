#!/usr/bin/env python
"""Median cold-L2 time of the proposal layer: python tools/time_proposal.py {C1|C4|C4x3} [iters]
(RLOD_PROPOSAL_V1=1 selects round 1's one-CTA-per-image select/sort kernel)."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402
from rlobjectdetection_b200.model.rpn.generate_anchors import generate_anchors  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
B, A, H, W, imh, imw, pre, post, scales = {
    "C1": (1, 9, 37, 62, 600, 1000, 12000, 2000, (8, 16, 32)),
    "C4": (24, 12, 50, 75, 800, 1200, 6000, 300, (4, 8, 16, 32)),
    "C4x3": (3, 12, 50, 75, 800, 1200, 6000, 300, (4, 8, 16, 32)),
}[cfg]
scores, deltas, im_info = syn.rpn_outputs(5, B, A, H, W, imh, imw, imh / 600.0)
anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array([0.5, 1, 2]))).float().to(dev)
sd, dd, ii = scores.to(dev), deltas.to(dev), im_info.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: be.proposal_forward(sd, dd, ii, anchors, 16, pre, post, 0.7)  # noqa: E731
for _ in range(3):
    fn()
prof_on = os.environ.get("RLOD_TIME_NO_PROFILE") is None  # event brackets between launches defeat programmatic dependent launches
be.lib().rlod_profile_only(-1)
be.lib().rlod_profile_enable(1 if prof_on else 0)
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
be.lib().rlod_profile_enable(0)
prof = {k: round(1e3 * v[0] / v[1], 1) for k, v in be.profile_collect().items()}
print(f"proposal {cfg} v1={os.environ.get('RLOD_PROPOSAL_V1', '')}: median {statistics.median(ts):.1f} us  min {min(ts):.1f} us; "
      f"per launch (us, bracketed by events): {prof}")
