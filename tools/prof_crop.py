import os, sys, torch
sys.path.insert(0, "/root/repo")
from rlobjectdetection_b200 import synthetic as syn
from rlobjectdetection_b200.model import _backend as be
dev = torch.device("cuda", 0)
B, C, H, W, n_per = 4, 1024, 38, 63, 256
g = torch.Generator().manual_seed(1)
rois = syn.rois_for_batch(2, B, n_per, H * 16.0, W * 16.0).to(dev)
grid = be.affine_grid(rois, (H, W), 14)
gyx = torch.stack([grid[..., 1], grid[..., 0]], 3).contiguous()
gout = torch.randn(rois.size(0), C, 14, 14, generator=g).to(dev)
for _ in range(3):
    be.roi_crop_backward(gout, gyx, (B, C, H, W))
torch.cuda.synchronize(); print("ok")
