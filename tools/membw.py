#!/usr/bin/env python
"""HBM write / read / copy rates with plain torch ops at the forward kernel's volumes (context for the roofline)."""
import torch
dev = torch.device("cuda", 0)
out = torch.empty(7200 * 1024 * 49, device=dev)          # 1.445 GB, the C4 output
feat = torch.randn(24 * 1024 * 50 * 75, device=dev)      # 369 MB, the C4 features
half = torch.empty(out.numel() // 2, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)
us = t(lambda: out.fill_(1.0)); print(f"fill 1.445 GB: {us:.1f} us = {out.numel()*4/us/1e3:.0f} GB/s")
us = t(lambda: feat.sum()); print(f"sum 369 MB: {us:.1f} us = {feat.numel()*4/us/1e3:.0f} GB/s")
us = t(lambda: half.copy_(out[:half.numel()])); print(f"copy 0.72 GB -> 0.72 GB: {us:.1f} us = {half.numel()*8/us/1e3:.0f} GB/s")
s2 = torch.cuda.Stream()
def both():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2): feat.sum()
    out.fill_(1.0)
    torch.cuda.current_stream().wait_stream(s2)
us = t(both); print(f"fill 1.445 GB || sum 369 MB: {us:.1f} us = {(out.numel()+feat.numel())*4/us/1e3:.0f} GB/s")
