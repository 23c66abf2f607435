"""GPU: pins the CPU oracle (and therefore every parity claim) against the REFERENCE ITSELF --
its legacy CUDA kernels compiled UNCHANGED for sm_100a into oracle/_ref/libref_legacy.so
(oracle/Makefile; sources stay under /root/reference) -- and checks the new kernels against
the same legacy outputs directly.  Skipped when the prebuilt library did not travel."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402

DEV = "cuda:0"
P = ctypes.c_void_p


@pytest.fixture(scope="module")
def legacy(orc):
    lib = orc.ref_legacy()
    if lib is None:
        pytest.skip("oracle/_ref/libref_legacy.so not present (built only where /root/reference exists)")
    lib.nms_cuda_compute.restype = None
    lib.nms_cuda_compute.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_float]
    for name in ("ROIAlignForwardLaucher", "ROIPoolForwardLaucher"):
        getattr(lib, name).restype = ctypes.c_int
    lib.ROIAlignForwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 6 + [P, P, P]
    lib.ROIAlignBackwardLaucher.restype = ctypes.c_int
    lib.ROIAlignBackwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 7 + [P, P, P]
    lib.ROIPoolForwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 6 + [P, P, P, P]
    lib.ROIPoolBackwardLaucher.restype = ctypes.c_int
    lib.ROIPoolBackwardLaucher.argtypes = [P, ctypes.c_float] + [ctypes.c_int] * 7 + [P, P, P, P]
    return lib


def dp(t):
    return P(t.data_ptr())


def legacy_nms(lib, dets_np, thresh):
    """nms_cuda_compute copies `boxes` with cudaMemcpyHostToDevice and writes keep/num with
    cudaMemcpyHostToDevice (nms_cuda_kernel.cu:97-100, 147, 154): host boxes, device outputs."""
    n = dets_np.shape[0]
    host = np.ascontiguousarray(dets_np, dtype=np.float32)
    keep = torch.zeros(n, dtype=torch.int32, device=DEV)
    num = torch.zeros(1, dtype=torch.int32, device=DEV)
    torch.cuda.synchronize()
    lib.nms_cuda_compute(dp(keep), dp(num), P(host.ctypes.data), n, host.shape[1], thresh)
    torch.cuda.synchronize()
    k = int(num.item())
    return keep[:k].cpu().numpy()


@pytest.mark.parametrize("n,thresh", [(65, 0.5), (700, 0.3), (3000, 0.7), (12000, 0.7)])
def test_nms_oracle_and_kernel_vs_legacy(orc, legacy, n, thresh):
    g = torch.Generator().manual_seed(n)
    bx = syn.random_boxes(g, n, 600, 1000, 8.0, 300.0)
    sc = syn.distinct_scores(g, (n,)).sort(descending=True).values
    dets = torch.cat([bx, sc[:, None]], 1).contiguous()
    ref = legacy_nms(legacy, dets.numpy(), thresh)
    assert np.array_equal(orc.nms(dets.numpy(), thresh), ref)          # oracle == reference
    keep, num = be.nms_padded(dets.to(DEV), thresh)
    assert np.array_equal(keep[: int(num.item())].cpu().numpy(), ref)  # new kernel == reference


def test_nms_near_ties_vs_legacy(orc, legacy):
    rows = []
    for s in range(1, 60):
        rows += [[0, 0, 9, 9], [0, 0, 9, 4 + (s % 3)], [s, 0, 9 + s, 9], [0, s, 9, 9 + s], [0, 0, 2 * s, s]]
    boxes = np.array(rows, dtype=np.float32)
    dets = np.concatenate([boxes, np.linspace(1, 0, len(boxes), dtype=np.float32)[:, None]], 1)
    for thresh in (0.5, 0.25, 1.0 / 3.0, 0.7):
        ref = legacy_nms(legacy, dets, thresh)
        assert np.array_equal(orc.nms(dets, thresh), ref)
        keep, num = be.nms_padded(torch.from_numpy(dets).to(DEV), thresh)
        assert np.array_equal(keep[: int(num.item())].cpu().numpy(), ref)


def _case(seed, B, C, H, W, n_per):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(seed + 1, B, n_per, H * 16.0, W * 16.0)
    return feat, rois


@pytest.mark.parametrize("shape", [(2, 8, 20, 31, 16), (4, 64, 38, 63, 64)])
def test_roi_align_vs_legacy(orc, legacy, shape):
    B, C, H, W, n_per = shape
    feat, rois = _case(21, B, C, H, W, n_per)
    R = rois.size(0)
    f, r = feat.to(DEV), rois.to(DEV)
    top = torch.zeros(R, C, 8, 8, device=DEV)
    torch.cuda.synchronize()
    legacy.ROIAlignForwardLaucher(dp(f), 1 / 16.0, R, H, W, C, 8, 8, dp(r), dp(top), None)
    torch.cuda.synchronize()
    grid_ref = top.cpu().numpy()
    # the oracle restates the legacy kernel's arithmetic operation for operation
    grid_orc = orc.roi_align_grid(feat.numpy(), rois.numpy(), 8, 8, 1 / 16.0)
    assert np.array_equal(grid_orc, grid_ref), np.abs(grid_orc - grid_ref).max()
    # reference RoIAlignAvg = legacy grid + F.avg_pool2d(2, 1)  (modules/roi_align.py:26-29)
    avg_ref = torch.nn.functional.avg_pool2d(top, kernel_size=2, stride=1).cpu().numpy()
    assert np.array_equal(orc.pool2x2(grid_ref, False), avg_ref)
    out = be.roi_align_forward(f, r, 7, 7, 1 / 16.0, be.POOL_AVG).cpu().numpy()
    scale = np.abs(avg_ref).max()
    np.testing.assert_allclose(out, avg_ref, rtol=1e-5, atol=1e-5 * scale)
    mx_ref = torch.nn.functional.max_pool2d(top, kernel_size=2, stride=1).cpu().numpy()
    out = be.roi_align_forward(f, r, 7, 7, 1 / 16.0, be.POOL_MAX).cpu().numpy()
    np.testing.assert_allclose(out, mx_ref, rtol=1e-5, atol=1e-5 * scale)
    # backward: legacy atomics (order-nondeterministic) after autograd's avg_pool2d backward
    g = torch.Generator().manual_seed(5)
    gout = torch.randn(R, C, 7, 7, generator=g).to(DEV)
    x = top.clone().requires_grad_(True)
    torch.nn.functional.avg_pool2d(x, kernel_size=2, stride=1).backward(gout)
    ggrid = x.grad.contiguous()
    bottom = torch.zeros(B, C, H, W, device=DEV)
    torch.cuda.synchronize()
    legacy.ROIAlignBackwardLaucher(dp(ggrid), 1 / 16.0, B, R, H, W, C, 8, 8, dp(r), dp(bottom), None)
    torch.cuda.synchronize()
    bref = bottom.cpu().numpy()
    bscale = np.abs(bref).max()
    borc = orc.roi_align_bwd(gout.cpu().numpy(), feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    np.testing.assert_allclose(borc, bref, rtol=1e-5, atol=1e-5 * bscale)
    gin = be.roi_align_backward(gout, r, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG).cpu().numpy()
    np.testing.assert_allclose(gin, bref, rtol=1e-5, atol=1e-5 * bscale)


@pytest.mark.parametrize("shape", [(2, 8, 20, 31, 16), (3, 32, 38, 63, 50)])
def test_roi_pool_vs_legacy(orc, legacy, shape):
    B, C, H, W, n_per = shape
    feat, rois = _case(31, B, C, H, W, n_per)
    R = rois.size(0)
    f, r = feat.to(DEV), rois.to(DEV)
    top = torch.zeros(R, C, 7, 7, device=DEV)
    arg = torch.zeros(R, C, 7, 7, dtype=torch.int32, device=DEV)
    torch.cuda.synchronize()
    legacy.ROIPoolForwardLaucher(dp(f), 1 / 16.0, R, H, W, C, 7, 7, dp(r), dp(top), dp(arg), None)
    torch.cuda.synchronize()
    ro, ra = orc.roi_pool(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0)
    assert np.array_equal(ro, top.cpu().numpy()) and np.array_equal(ra, arg.cpu().numpy())
    out, am = be.roi_pool_forward(f, r, 7, 7, 1 / 16.0)
    assert torch.equal(out, top) and torch.equal(am, arg)
    gout = torch.randn(R, C, 7, 7, generator=torch.Generator().manual_seed(6)).to(DEV)
    bottom = torch.zeros(B, C, H, W, device=DEV)
    torch.cuda.synchronize()
    legacy.ROIPoolBackwardLaucher(dp(gout), 1 / 16.0, B, R, H, W, C, 7, 7, dp(r), dp(bottom), dp(arg), None)
    torch.cuda.synchronize()
    bref = bottom.cpu().numpy()
    bscale = np.abs(bref).max()
    np.testing.assert_allclose(orc.roi_pool_bwd(gout.cpu().numpy(), ra, (B, C, H, W), rois.numpy(), 1 / 16.0),
                               bref, rtol=1e-5, atol=1e-5 * bscale)
    gin = be.roi_pool_backward(gout, am, r, (B, C, H, W), 7, 7, 1 / 16.0).cpu().numpy()
    np.testing.assert_allclose(gin, bref, rtol=1e-5, atol=1e-5 * bscale)


def test_roi_crop_vs_legacy(orc, legacy):
    """BilinearSamplerBHWD launchers of the reference (roi_crop_cuda_kernel.cu:205-330), compiled
    unchanged: oracle and new kernels against them."""
    if not hasattr(legacy, "BilinearSamplerBHWD_updateOutput_cuda_kernel"):
        pytest.skip("libref_legacy.so was built without roi_crop")
    I = ctypes.c_int
    fwd = legacy.BilinearSamplerBHWD_updateOutput_cuda_kernel
    fwd.restype = I
    fwd.argtypes = [I] * 8 + [P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P]
    bwd = legacy.BilinearSamplerBHWD_updateGradInput_cuda_kernel
    bwd.restype = I
    bwd.argtypes = [I] * 8 + [P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P, I, I, I, I, P]
    B, C, H, W, n_per, gs = 2, 12, 20, 31, 11, 14
    feat, rois = _case(41, B, C, H, W, n_per)
    R = rois.size(0)
    grid_xy = orc.affine_grid(rois.numpy(), H, W, gs, True)
    grid_yx = torch.from_numpy(np.ascontiguousarray(np.stack([grid_xy[..., 1], grid_xy[..., 0]], 3)))
    f, gy = feat.to(DEV), grid_yx.to(DEV)
    out = torch.zeros(R, C, gs, gs, device=DEV)
    torch.cuda.synchronize()
    # argument order of roi_crop_cuda.c:21-44: sizes, then (data, stride0, stride1, stride2, stride3) of
    # input / grids (stride 0, 3, 1, 2) / output
    assert fwd(C, gs, gs, R, C, H, W, B, dp(f), *f.stride(), dp(gy), gy.stride(0), gy.stride(3), gy.stride(1), gy.stride(2),
               dp(out), *out.stride(), None) == 1
    torch.cuda.synchronize()
    ref = out.cpu().numpy()
    scale = np.abs(ref).max()
    np.testing.assert_allclose(orc.roi_crop(feat.numpy(), grid_yx.numpy()), ref, rtol=1e-5, atol=1e-6 * scale)
    ours = be.roi_crop_forward(f, gy).cpu().numpy()
    np.testing.assert_allclose(ours, ref, rtol=1e-6, atol=1e-6 * scale)
    gout = torch.randn(R, C, gs, gs, generator=torch.Generator().manual_seed(7)).to(DEV)
    gin_ref = torch.zeros(B, C, H, W, device=DEV)
    ggrid = torch.zeros_like(gy)
    torch.cuda.synchronize()
    assert bwd(C, gs, gs, R, C, H, W, B, dp(f), *f.stride(), dp(gy), gy.stride(0), gy.stride(3), gy.stride(1), gy.stride(2),
               dp(gin_ref), *gin_ref.stride(), dp(ggrid), ggrid.stride(0), ggrid.stride(3), ggrid.stride(1), ggrid.stride(2),
               dp(gout), *gout.stride(), None) == 1
    torch.cuda.synchronize()
    assert float(ggrid.abs().max()) == 0.0  # the reference stores no grid gradient
    bref = gin_ref.cpu().numpy()
    bscale = np.abs(bref).max()
    gin = be.roi_crop_backward(gout, gy, (B, C, H, W)).cpu().numpy()
    np.testing.assert_allclose(gin, bref, rtol=1e-5, atol=1e-5 * bscale)
    np.testing.assert_allclose(orc.roi_crop_bwd(gout.cpu().numpy(), grid_yx.numpy(), (B, C, H, W)), bref, rtol=1e-5,
                               atol=1e-5 * bscale)
