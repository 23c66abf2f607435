"""GPU parity: the sm_100a kernels (through the C ABI / the reference-shaped modules) against the
CPU oracle on identical seeded inputs.  Integer / index results are bit-exact; pooled features,
gradients and rewards are within 1e-5 relative (tolerance written at each assert)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from rlobjectdetection_b200 import synthetic as syn  # noqa: E402
from rlobjectdetection_b200.model import _backend as be  # noqa: E402

DEV = "cuda:0"
RTOL = 1e-5


def close(a, ref, rtol=RTOL, what=""):
    """|a - ref| <= rtol*|ref| + rtol*max|ref|  (fp32 outputs of O(1) data)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=rtol * max(scale, 1e-30), err_msg=what)


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)) if not torch.is_tensor(x) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


# ------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------
def _sorted_dets(seed, n, im=600.0, smin=8.0, smax=200.0):
    g = torch.Generator().manual_seed(seed)
    bx = syn.random_boxes(g, n, im, im, smin, smax)
    sc = syn.distinct_scores(g, (n,)).sort(descending=True).values
    return torch.cat([bx, sc[:, None]], 1).contiguous()


@pytest.mark.parametrize("force_large", [0, 1, 2])  # 2: the scan kernel of segments beyond 57 344 boxes
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 300, 511, 513, 1000, 3000])
@pytest.mark.parametrize("thresh", [0.3, 0.7])
def test_nms_bit_exact(orc, n, thresh, force_large):
    dets = _sorted_dets(100 + n, n)
    ref = orc.nms(dets.numpy(), thresh)
    prev = be.lib().rlod_debug_nms_force_large(force_large)
    try:
        keep, num = be.nms_padded(cu(dets), thresh)
    finally:
        be.lib().rlod_debug_nms_force_large(prev)
    k = int(num.item())
    assert k == len(ref)
    assert np.array_equal(keep[:k].cpu().numpy(), ref)
    assert (keep[k:].cpu().numpy() == -1).all()


@pytest.mark.parametrize("max_keep", [1, 7, 64, 65, 300])
def test_nms_max_keep(orc, max_keep):
    dets = _sorted_dets(7, 2000)
    ref = orc.nms(dets.numpy(), 0.5, max_keep=max_keep)
    keep, num = be.nms_padded(cu(dets), 0.5, max_keep=max_keep)
    k = int(num.item())
    assert k == len(ref) == min(max_keep, len(orc.nms(dets.numpy(), 0.5)))
    assert np.array_equal(keep[:k].cpu().numpy(), ref)


@pytest.mark.parametrize("max_keep", [0, 600, 777, 1500])
@pytest.mark.parametrize("smax", [60.0, 400.0])
def test_nms_decoupled_scan_early_stop(orc, max_keep, smax):
    """Row-major mask + decoupled scan (n > 512, more than 512 keeps allowed): light and heavy suppression,
    stops in the middle of a block and of the far warps' row-block schedule, ragged last block."""
    dets = _sorted_dets(31, 5003, smax=smax)
    ref = orc.nms(dets.numpy(), 0.6, max_keep=max_keep)
    keep, num = be.nms_padded(cu(dets), 0.6, max_keep=max_keep)
    k = int(num.item())
    assert k == len(ref)
    assert np.array_equal(keep[:k].cpu().numpy(), ref)
    assert (keep[k:].cpu().numpy() == -1).all()


def test_nms_near_threshold_and_degenerate(orc):
    # integer boxes whose IoU sits exactly on / one ulp around the threshold, duplicates,
    # inverted and zero-size boxes: decisions must equal the IEEE-division reference
    rows = []
    for s in range(1, 40):
        rows += [[0, 0, 9, 9], [0, 0, 9, 4 + (s % 3)], [s, 0, 9 + s, 9], [0, s, 9, 9 + s],
                 [3, 3, 3, 3], [10, 10, 5, 5], [0, 0, 2 * s, s], [s, s, 3 * s, 2 * s]]
    base = np.array(rows, dtype=np.float32)
    rng = np.random.default_rng(5)
    jit = base + rng.integers(0, 2, base.shape).astype(np.float32)
    boxes = np.concatenate([base, jit, base + 0.5], 0)
    dets = np.concatenate([boxes, np.linspace(1, 0, len(boxes), dtype=np.float32)[:, None]], 1)
    for thresh in (0.5, 0.25, 1.0 / 3.0, 0.7, 0.0, 0.9999999, -0.5, 1.5):
        ref = orc.nms(dets, thresh)
        for fl in (0, 1, 2):
            prev = be.lib().rlod_debug_nms_force_large(fl)
            try:
                keep, num = be.nms_padded(cu(dets), thresh)
            finally:
                be.lib().rlod_debug_nms_force_large(prev)
            k = int(num.item())
            assert k == len(ref), (thresh, fl)
            assert np.array_equal(keep[:k].cpu().numpy(), ref), (thresh, fl)


def test_nms_stride4_and_wrapper(orc):
    from rlobjectdetection_b200.model.nms.nms_wrapper import nms
    dets = _sorted_dets(11, 700)
    ref = orc.nms(dets.numpy(), 0.7)
    out = nms(cu(dets), 0.7)
    assert out.dtype == torch.int32 and out.shape == (len(ref), 1)
    assert np.array_equal(out.view(-1).cpu().numpy(), ref)
    keep, num = be.nms_padded(cu(dets[:, :4].contiguous()), 0.7)  # float4 path
    assert np.array_equal(keep[: int(num.item())].cpu().numpy(), ref)
    assert nms(torch.zeros(0, 5, device=DEV), 0.7) == []  # reference: nms_wrapper.py:13-14


def test_nms_c1_size(orc):
    # config 1 size: 12000 boxes, thr 0.7 -- full keep list and the post-NMS 2000 cut
    dets = _sorted_dets(1, 12000, im=1000.0, smin=16.0, smax=400.0)
    ref = orc.nms(dets.numpy(), 0.7)
    keep, num = be.nms_padded(cu(dets), 0.7)
    k = int(num.item())
    assert k == len(ref) and np.array_equal(keep[:k].cpu().numpy(), ref)
    keep2, num2 = be.nms_padded(cu(dets), 0.7, max_keep=2000)
    k2 = int(num2.item())
    assert k2 == min(2000, len(ref)) and np.array_equal(keep2[:k2].cpu().numpy(), ref[:k2])


@pytest.mark.parametrize("shape", [(3, 5, 300), (64, 81, 300)])
def test_nms_batched_per_class(orc, shape):
    # config 5: images x classes segments of 300 clustered boxes, thr 0.3, one launch
    ni, nc, per = shape
    dets, seg = syn.clustered_dets(4, ni, nc, per)
    rk, rn = orc.nms_batched(dets.numpy(), seg.numpy(), 0.3)
    keep, num = be.nms_batched(cu(dets), cu(seg), 0.3, max_seg=per)
    assert np.array_equal(num.cpu().numpy(), rn)
    assert np.array_equal(keep.cpu().numpy(), rk)


def test_nms_batched_ragged(orc):
    g = torch.Generator().manual_seed(9)
    lens = [0, 1, 64, 0, 129, 700, 5, 513, 0]
    parts = [_sorted_dets(50 + i, n) for i, n in enumerate(lens) if n > 0]
    dets = torch.cat(parts, 0)
    seg = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    rk, rn = orc.nms_batched(dets.numpy(), seg.numpy(), 0.5)
    for max_seg in (max(lens), None):
        keep, num = be.nms_batched(cu(dets), cu(seg), 0.5, max_seg=max_seg)
        assert np.array_equal(num.cpu().numpy(), rn)
        assert np.array_equal(keep.cpu().numpy(), rk)
    del g


# ------------------------------------------------------------------------------------------
# RoIAlign
# ------------------------------------------------------------------------------------------
def _align_case(seed, B, C, H, W, n_per, stride=16.0, shuffle=False):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(seed + 1, B, n_per, H * stride, W * stride)
    if shuffle:
        rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    return feat, rois


@pytest.mark.parametrize("mode", [be.POOL_NONE, be.POOL_AVG, be.POOL_MAX])
@pytest.mark.parametrize("case", [
    dict(B=1, C=2, H=5, W=6, n_per=8, p=7),          # tiny map, generic path (R small? no: fast needs C%4)
    dict(B=2, C=8, H=20, W=31, n_per=16, p=7),        # fast path, edge rois
    dict(B=3, C=12, H=38, W=63, n_per=40, p=7),       # fast path
    dict(B=2, C=6, H=19, W=23, n_per=12, p=7),        # C % 4 != 0 -> generic
    dict(B=2, C=8, H=16, W=16, n_per=9, p=3),         # other pooled size -> generic
    dict(B=2, C=4, H=12, W=40, n_per=10, p=7, shuffle=True),  # rois not grouped by image
    dict(B=2, C=8, H=21, W=31, n_per=12, p=7),        # H*W odd: planes are 4-byte aligned only -> LDGSTS fill
    dict(B=1, C=4, H=96, W=90, n_per=16, p=7),        # (H+2)*pitch > 8192 pixels -> generic kernels
])
def test_roi_align_forward(orc, case, mode):
    p = case["p"]
    feat, rois = _align_case(3, case["B"], case["C"], case["H"], case["W"], case["n_per"],
                             shuffle=case.get("shuffle", False))
    ah = p + 1 if mode == be.POOL_NONE else p  # NONE: sample an (p+1)x(p+1) grid directly
    ref = orc.roi_align(feat.numpy(), rois.numpy(), ah, ah, 1 / 16.0, pool_mode=mode)
    out = be.roi_align_forward(cu(feat), cu(rois), ah, ah, 1 / 16.0, mode)
    close(out.cpu().numpy(), ref, what=f"roi_align fwd {case} mode {mode}")


def test_roi_align_forward_bad_batch_index(orc):
    feat, rois = _align_case(5, 2, 8, 20, 30, 8)
    rois[3, 0] = 7.0   # image 7 of 2
    rois[5, 0] = -1.0
    out = be.roi_align_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0, be.POOL_AVG).cpu().numpy()
    assert (out[3] == 0).all() and (out[5] == 0).all()
    good = [i for i in range(rois.size(0)) if i not in (3, 5)]
    ref = orc.roi_align(feat.numpy(), rois[good].numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    close(out[good], ref)


def test_roi_align_images_without_rois_and_empty_input(orc):
    # images 0 and 2 of 3 have no roi at all (their CTAs return at once); R == 0 is a no-op
    feat, rois = _align_case(6, 3, 8, 20, 30, 10)
    only1 = rois[rois[:, 0] == 1].contiguous()
    ref = orc.roi_align(feat.numpy(), only1.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    out = be.roi_align_forward(cu(feat), cu(only1), 7, 7, 1 / 16.0, be.POOL_AVG)
    close(out.cpu().numpy(), ref)
    g = torch.Generator().manual_seed(9)
    gout = torch.randn(only1.size(0), 8, 7, 7, generator=g)
    gref = orc.roi_align_bwd(gout.numpy(), feat.numpy(), only1.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    gin = be.roi_align_backward(cu(gout), cu(only1), None, tuple(feat.shape), 7, 7, 1 / 16.0, be.POOL_AVG)
    close(gin.cpu().numpy(), gref)
    assert (gin[0] == 0).all() and (gin[2] == 0).all()
    empty = be.roi_align_forward(cu(feat), cu(torch.zeros(0, 5)), 7, 7, 1 / 16.0, be.POOL_AVG)
    assert tuple(empty.shape) == (0, 8, 7, 7)


@pytest.mark.parametrize("case", [
    dict(B=3, C=12, H=38, W=63, n_per=40, mode=be.POOL_AVG),   # plane kernel
    dict(B=2, C=6, H=19, W=23, n_per=12, mode=be.POOL_MAX),    # generic kernels (C % 4 != 0)
    dict(B=2, C=8, H=16, W=16, n_per=9, mode=be.POOL_NONE),    # plane kernel, 8 x 8 output
])
def test_roi_align_planned_forward_equals_forward(case):
    """rlod_roi_align_plan + rlod_roi_align_forward_planned (the plan made on ANOTHER stream, ordered by an event)
    == rlod_roi_align_forward, bit for bit, for rois grouped by image and for a concatenation of two roi sets."""
    from rlobjectdetection_b200.model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    feat, rois = _align_case(21, case["B"], case["C"], case["H"], case["W"], case["n_per"])
    _, rois2 = _align_case(22, case["B"], case["C"], case["H"], case["W"], case["n_per"])
    mod = {be.POOL_AVG: RoIAlignAvg, be.POOL_MAX: RoIAlignMax, be.POOL_NONE: RoIAlign}[case["mode"]]
    p = 8 if case["mode"] == be.POOL_NONE else 7
    layer = mod(p, p, 1 / 16.0)
    f = cu(feat)
    for r in (cu(rois), cu(torch.cat([rois, rois2]))):
        ref = layer(f, r)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            plan = layer.plan(r, tuple(f.shape))
            done = torch.cuda.Event()
            done.record(side)
        torch.cuda.current_stream().wait_event(done)
        with torch.no_grad():
            out = layer.forward_planned(f, plan)
        assert torch.equal(out, ref)
    with pytest.raises(ValueError):
        layer.forward_planned(f[:, :, :-1], plan)


@pytest.mark.parametrize("grouped", [True, False])
def test_roi_align_long_roi_lists(orc, grouped):
    """More rois per image than the list kernels hold in shared memory (2048): the lists are collected but not
    partitioned by walk mode; grouped by image and as a concatenation of two roi sets; forward and backward."""
    B, C, H, W, n_per = 2, 8, 24, 31, 1300
    feat, ra = _align_case(31, B, C, H, W, n_per)
    _, rb = _align_case(32, B, C, H, W, n_per)
    if grouped:
        rois = torch.cat([ra.view(B, n_per, 5), rb.view(B, n_per, 5)], 1).reshape(-1, 5).contiguous()
    else:
        rois = torch.cat([ra, rb]).contiguous()
    ref = orc.roi_align(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    out = be.roi_align_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0, be.POOL_AVG)
    close(out.cpu().numpy(), ref, what="long roi lists fwd")
    g = torch.Generator().manual_seed(5)
    gout = torch.randn(rois.size(0), C, 7, 7, generator=g)
    gref = orc.roi_align_bwd(gout.numpy(), feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    gin = be.roi_align_backward(cu(gout), cu(rois), None, tuple(feat.shape), 7, 7, 1 / 16.0, be.POOL_AVG)
    close(gin.cpu().numpy(), gref, what="long roi lists bwd")


@pytest.mark.parametrize("mode", [be.POOL_NONE, be.POOL_AVG, be.POOL_MAX])
@pytest.mark.parametrize("shuffle", [False, True])
def test_roi_align_forward_split_tail(orc, mode, shuffle):
    """Few (image, 4-channel) items and long roi lists: the pooling launch serves every item of its last wave
    with two or three CTAs, each with a contiguous share of the image's roi groups (fwd_tail_split).  Ragged
    lists: an image with fewer groups than CTAs (empty shares), an image without rois, shares of unequal size."""
    B, C, H, W = 4, 8, 30, 41
    counts = [3, 0, 1021, 517]
    g = torch.Generator().manual_seed(77)
    feat = torch.randn(B, C, H, W, generator=g)
    parts = []
    for b, n in enumerate(counts):
        r = syn.rois_for_batch(100 + b, 1, max(n, 1), H * 16.0, W * 16.0)[:n]
        r[:, 0] = b
        parts.append(r)
    rois = torch.cat(parts).contiguous()
    if shuffle:
        rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    ah = 8 if mode == be.POOL_NONE else 7
    ref = orc.roi_align(feat.numpy(), rois.numpy(), ah, ah, 1 / 16.0, pool_mode=mode)
    out = be.roi_align_forward(cu(feat), cu(rois), ah, ah, 1 / 16.0, mode)
    close(out.cpu().numpy(), ref, what=f"split tail mode {mode} shuffle {shuffle}")


@pytest.mark.parametrize("mode", [be.POOL_NONE, be.POOL_AVG, be.POOL_MAX])
@pytest.mark.parametrize("case", [
    dict(B=2, C=8, H=200, W=304, n_per=1500, stride=4.0),               # FPN P2 of 800 x 1216: 6 x 10 tiles
    dict(B=3, C=4, H=50, W=200, n_per=130, stride=16.0, shuffle=True),  # tiled along x only, rois not grouped
    dict(B=1, C=12, H=260, W=40, n_per=130, stride=8.0),                # tiled along y only, last tile hangs over
    dict(B=2, C=4, H=100, W=167, n_per=380, stride=8.0, small=True),    # every roi fits a tile
    dict(B=2, C=4, H=100, W=167, n_per=40, stride=8.0),                 # too few rois per tile: generic kernel
])
def test_roi_align_forward_tiled_large_maps(orc, case, mode):
    """Maps beyond one CTA's shared memory: the plane kernel runs over overlapping tiles of the map (every roi in
    the tile that holds all of its taps), rois larger than half a tile go through the generic kernel; the result is
    the oracle's whatever the route.  Rois touch the borders, lie outside, are inverted (rois_for_batch)."""
    B, C, H, W, stride = case["B"], case["C"], case["H"], case["W"], case["stride"]
    g = torch.Generator().manual_seed(41)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(42, B, case["n_per"], H * stride, W * stride)
    if case.get("small"):
        ctr = 0.5 * (rois[:, 1:3] + rois[:, 3:5])
        half = (0.5 * (rois[:, 3:5] - rois[:, 1:3])).clamp(-12 * stride, 12 * stride)
        rois[:, 1:3], rois[:, 3:5] = ctr - half, ctr + half
    if case.get("shuffle"):
        rois = rois[torch.randperm(rois.size(0), generator=g)].contiguous()
    ah = 8 if mode == be.POOL_NONE else 7
    ref = orc.roi_align(feat.numpy(), rois.numpy(), ah, ah, 1 / stride, pool_mode=mode)
    f, r = cu(feat), cu(rois)
    n0 = be.lib().rlod_launch_count()
    out = be.roi_align_forward(f, r, ah, ah, 1 / stride, mode)
    # tiled: plan + lists + plane kernel + generic kernel for the rois beyond a tile; else plan + fix-up + generic
    assert be.lib().rlod_launch_count() - n0 == (3 if case["n_per"] == 40 else 4)
    close(out.cpu().numpy(), ref, what=f"tiled {case} mode {mode}")


def test_roi_align_tiled_planned_forward_equals_forward():
    from rlobjectdetection_b200.model.roi_align.modules.roi_align import RoIAlignAvg
    g = torch.Generator().manual_seed(43)
    feat = cu(torch.randn(2, 8, 120, 150, generator=g))
    rois = cu(syn.rois_for_batch(44, 2, 320, 120 * 8.0, 150 * 8.0))  # 3 x 4 tiles, 26 rois per tile
    layer = RoIAlignAvg(7, 7, 1 / 8.0)
    ref = layer(feat, rois)
    plan = layer.plan(rois, tuple(feat.shape))
    with torch.no_grad():
        out = layer.forward_planned(feat, plan)
    assert torch.equal(out, ref)


def test_roi_align_c2_full_size(orc):
    # config 2: Res-101 C4 at 600x1000 -> (4,1024,38,63), 4 x 256 rois, 7x7
    feat, rois = _align_case(1, 4, 1024, 38, 63, 256)
    ref = orc.roi_align(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    out = be.roi_align_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0, be.POOL_AVG)
    close(out.cpu().numpy(), ref, what="C2 fwd")


@pytest.mark.parametrize("mode", [be.POOL_NONE, be.POOL_AVG, be.POOL_MAX])
@pytest.mark.parametrize("case", [
    dict(B=2, C=8, H=20, W=31, n_per=16, p=7),
    dict(B=3, C=12, H=38, W=63, n_per=40, p=7),
    dict(B=2, C=6, H=19, W=23, n_per=12, p=7),
    dict(B=2, C=8, H=16, W=16, n_per=9, p=3),
    dict(B=2, C=4, H=50, W=75, n_per=30, p=7, shuffle=True),
    dict(B=2, C=8, H=21, W=31, n_per=12, p=7),
])
def test_roi_align_backward(orc, case, mode):
    p = case["p"]
    feat, rois = _align_case(4, case["B"], case["C"], case["H"], case["W"], case["n_per"],
                             shuffle=case.get("shuffle", False))
    ah = p + 1 if mode == be.POOL_NONE else p
    g = torch.Generator().manual_seed(44)
    gout = torch.randn(rois.size(0), case["C"], ah, ah, generator=g)
    ref = orc.roi_align_bwd(gout.numpy(), feat.numpy(), rois.numpy(), ah, ah, 1 / 16.0, pool_mode=mode)
    gin = be.roi_align_backward(cu(gout), cu(rois), cu(feat), tuple(feat.shape), ah, ah, 1 / 16.0, mode)
    close(gin.cpu().numpy(), ref, what=f"roi_align bwd {case} mode {mode}")
    # accumulate form
    base = torch.randn(feat.shape, generator=g)
    acc = cu(base.clone())
    be.roi_align_backward(cu(gout), cu(rois), cu(feat), tuple(feat.shape), ah, ah, 1 / 16.0, mode,
                          grad_in=acc)
    close(acc.cpu().numpy(), ref + base.numpy().astype(np.float64), what="bwd accumulate")


def test_roi_align_backward_c2_full_size(orc):
    feat, rois = _align_case(1, 4, 1024, 38, 63, 256)
    g = torch.Generator().manual_seed(2)
    gout = torch.randn(rois.size(0), 1024, 7, 7, generator=g)
    ref = orc.roi_align_bwd(gout.numpy(), feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0,
                            pool_mode=orc.POOL_AVG)
    gin = be.roi_align_backward(cu(gout), cu(rois), None, tuple(feat.shape), 7, 7, 1 / 16.0,
                                be.POOL_AVG)
    close(gin.cpu().numpy(), ref, what="C2 bwd")
    # run to run: the flush token ring of k_align8_bwd_own fixes the order in which rois add into a pixel
    # (the reference's atomicAdd order is not fixed), so repeats are bit-identical
    gin2 = be.roi_align_backward(cu(gout), cu(rois), None, tuple(feat.shape), 7, 7, 1 / 16.0,
                                 be.POOL_AVG)
    assert torch.equal(gin2, gin), "RoIAlign backward is not bit-reproducible run to run"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [be.POOL_NONE, be.POOL_AVG])
def test_roi_align_backward_tiny_and_clamped_rois(orc, mode):
    """Rois whose 8 sample rows / columns fall into one or two pixels (every tap merged: up to 7 pooled
    columns per pixel column), rois hanging over every border (clamped indices, ratios > 1, invalid
    samples at either end), many rois on the same pixels (the token ring orders them), an image with a
    single roi and an image with none."""
    B, C, H, W = 4, 8, 20, 27
    ah = 8 if mode == be.POOL_NONE else 7
    g = torch.Generator().manual_seed(77)
    rows = []
    for k in range(40):  # sub-pixel to 3-pixel rois around the same spot of image 0
        x1 = 100.0 + float(torch.rand(1, generator=g)) * 8
        y1 = 80.0 + float(torch.rand(1, generator=g)) * 8
        w = float(torch.rand(1, generator=g)) * 48
        h = float(torch.rand(1, generator=g)) * 48
        rows.append([0, x1, y1, x1 + w, y1 + h])
    for k in range(24):  # over the borders of image 1
        cx = float(torch.rand(1, generator=g)) * W * 16
        cy = float(torch.rand(1, generator=g)) * H * 16
        side = k % 4
        if side == 0: rows.append([1, -60.0 - k, cy - 30, 50.0 + k, cy + 30])
        if side == 1: rows.append([1, cx - 40, -35.0 - k, cx + 40, 20.0 + k])
        if side == 2: rows.append([1, W * 16 - 50.0 - k, cy - 20, W * 16 + 80.0, cy + 25])
        if side == 3: rows.append([1, cx - 25, H * 16 - 30.0 - k, cx + 25, H * 16 + 90.0])
    rows.append([3, 17.0, 23.0, 17.0, 23.0])  # image 3: one roi of one pixel; image 2: none
    rois = torch.tensor(rows, dtype=torch.float32)
    feat = torch.randn(B, C, H, W, generator=g)
    gout = torch.randn(rois.size(0), C, ah, ah, generator=g)
    ref = orc.roi_align_bwd(gout.numpy(), feat.numpy(), rois.numpy(), ah, ah, 1 / 16.0, pool_mode=mode)
    gin = be.roi_align_backward(cu(gout), cu(rois), cu(feat), tuple(feat.shape), ah, ah, 1 / 16.0, mode)
    close(gin.cpu().numpy(), ref, what=f"roi_align bwd tiny / clamped rois, mode {mode}")
    assert not gin[2].any(), "an image without rois must get a zero gradient"


def test_roi_align_module_autograd(orc):
    from rlobjectdetection_b200.model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    feat, rois = _align_case(6, 2, 8, 24, 30, 12)
    g = torch.Generator().manual_seed(3)
    for mod, mode, ah in ((RoIAlignAvg(7, 7, 1 / 16.0), orc.POOL_AVG, 7),
                          (RoIAlignMax(7, 7, 1 / 16.0), orc.POOL_MAX, 7),
                          (RoIAlign(8, 8, 1 / 16.0), orc.POOL_NONE, 8)):
        x = cu(feat).requires_grad_(True)
        y = mod(x, cu(rois))
        gout = torch.randn(y.shape, generator=g)
        y.backward(cu(gout))
        close(y.detach().cpu().numpy(), orc.roi_align(feat.numpy(), rois.numpy(), ah, ah, 1 / 16.0, pool_mode=mode))
        close(x.grad.cpu().numpy(), orc.roi_align_bwd(gout.numpy(), feat.numpy(), rois.numpy(), ah, ah,
                                                      1 / 16.0, pool_mode=mode))
    with pytest.raises(NotImplementedError):
        RoIAlignAvg(7, 7, 1 / 16.0)(feat, rois)  # CPU tensors: no fallback


def test_roi_align_properties_c3_size():
    # config 3 size (8 images 800x1200 -> 50x75x1024, 300 rois each): size-independent checks.
    # (1) constant map -> every fully-inside sample equals the constant; (2) linearity of the
    # backward in grad_out; (3) <fwd(x), g> == <x, bwd(g)> (adjointness) within fp32 error.
    B, C, H, W, n = 8, 1024, 50, 75, 300
    rois = cu(syn.rois_for_batch(2, B, n, 800, 1200))
    ones = torch.full((B, C, H, W), 2.5, device=DEV)
    y = be.roi_align_forward(ones, rois, 7, 7, 1 / 16.0, be.POOL_AVG)
    inside = (rois[:, 1] >= 0) & (rois[:, 2] >= 0) & (rois[:, 3] <= 1199 - 16) & (rois[:, 4] <= 799 - 16) \
        & (rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])
    yi = y[inside]
    assert yi.numel() > 0 and (yi - 2.5).abs().max().item() <= 2.5 * 1e-6
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, C, H, W, generator=g).to(DEV)
    g1 = torch.randn(B * n, C, 7, 7, generator=g).to(DEV)
    fx = be.roi_align_forward(x, rois, 7, 7, 1 / 16.0, be.POOL_AVG)
    b1 = be.roi_align_backward(g1, rois, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG)
    b2 = be.roi_align_backward(2.0 * g1, rois, None, (B, C, H, W), 7, 7, 1 / 16.0, be.POOL_AVG)
    # linear in grad_out (exact per roi; the order rois add into a pixel is not fixed run to run)
    assert (b2 - 2.0 * b1).abs().max().item() <= 1e-5 * b2.abs().max().item()
    lhs = (fx.double() * g1.double()).sum().item()
    rhs = (x.double() * b1.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)


# ------------------------------------------------------------------------------------------
# RoIPool
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [dict(B=2, C=8, H=20, W=31, n_per=16), dict(B=3, C=5, H=38, W=63, n_per=33)])
def test_roi_pool(orc, case):
    feat, rois = _align_case(12, case["B"], case["C"], case["H"], case["W"], case["n_per"])
    ro, ra = orc.roi_pool(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0)
    out, arg = be.roi_pool_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0)
    assert np.array_equal(out.cpu().numpy(), ro)     # max of fp32 values: bit-exact
    assert np.array_equal(arg.cpu().numpy(), ra)     # flat NCHW argmax, -1 for empty bins
    # inference form (no argmax buffer: the kernel tracks maxima only): same values; the module picks it under no_grad
    out_inf, none = be.roi_pool_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0, want_argmax=False)
    assert none is None and torch.equal(out_inf, out)
    from rlobjectdetection_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    layer = _RoIPooling(7, 7, 1 / 16.0)
    with torch.no_grad():
        assert torch.equal(layer(cu(feat), cu(rois)), out)
    g = torch.Generator().manual_seed(13)
    gout = torch.randn(out.shape, generator=g)
    ref = orc.roi_pool_bwd(gout.numpy(), ra, tuple(feat.shape), rois.numpy(), 1 / 16.0)
    gin = be.roi_pool_backward(cu(gout), arg, cu(rois), tuple(feat.shape), 7, 7, 1 / 16.0)
    close(gin.cpu().numpy(), ref, what="roi_pool bwd")


def test_roi_pool_module_autograd(orc):
    from rlobjectdetection_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    feat, rois = _align_case(14, 2, 8, 24, 30, 12)
    x = cu(feat).requires_grad_(True)
    y = _RoIPooling(7, 7, 1 / 16.0)(x, cu(rois))
    ro, ra = orc.roi_pool(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0)
    assert np.array_equal(y.detach().cpu().numpy(), ro)
    gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    y.backward(cu(gout))
    close(x.grad.cpu().numpy(), orc.roi_pool_bwd(gout.numpy(), ra, tuple(feat.shape), rois.numpy(), 1 / 16.0))


# ------------------------------------------------------------------------------------------
# box algebra + proposal layer
# ------------------------------------------------------------------------------------------
def test_box_algebra_vs_golden(orc, golden):
    from rlobjectdetection_b200.model.rpn import bbox_transform as bt
    dec = bt.bbox_transform_inv(cu(golden["dec_boxes"]), cu(golden["dec_deltas"]), 2).cpu().numpy()
    ref = golden["dec_out"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(dec), fin)
    # only exp() may differ (libdevice expf vs torch CPU): a few ulp on the exp-scaled terms
    np.testing.assert_allclose(dec[fin], ref[fin], rtol=3e-6, atol=1e-4)
    clp = bt.clip_boxes(cu(golden["dec_out"]).clone(), cu(golden["clip_im_info"]), 2).cpu().numpy()
    assert np.array_equal(clp, golden["clip_out"])
    o = bt.bbox_overlaps(cu(golden["ovl_anchors"]), cu(golden["ovl_gt"])).cpu().numpy()
    assert np.array_equal(o, golden["ovl_out"])
    o3 = bt.bbox_overlaps_batch(cu(golden["ovlb_anchors"]), cu(golden["ovlb_gt"])).cpu().numpy()
    assert np.array_equal(o3, golden["ovlb_out3"])
    o2 = bt.bbox_overlaps_batch(cu(golden["ovlb_anchors"][0]), cu(golden["ovlb_gt"])).cpu().numpy()
    assert np.array_equal(o2, golden["ovlb_out2"])
    # (B,N,5) anchors: columns 1:5 are the box
    a5 = np.concatenate([np.zeros((2, golden["ovlb_anchors"].shape[1], 1), np.float32),
                         golden["ovlb_anchors"]], 2)
    o5 = bt.bbox_overlaps_batch(cu(a5), cu(golden["ovlb_gt"])).cpu().numpy()
    assert np.array_equal(o5, golden["ovlb_out3"])


def _proposal_parity(orc, scores, deltas, im_info, anchors, stride, pre, post, thresh):
    rois, order, props, nkeep = be.proposal_forward(cu(scores), cu(deltas), cu(im_info), cu(anchors),
                                                    stride, pre, post, thresh, return_taps=True)
    rois, order, props, nkeep = (t.cpu().numpy() for t in (rois, order, props, nkeep))
    # stage 1: sort order bit-exact, decoded boxes equal up to the expf ulp
    o_rois, o_order, o_props, o_keep, o_nkeep = orc.proposal_layer(
        scores.numpy(), deltas.numpy(), im_info.numpy(), anchors.numpy(), stride, pre, post, thresh,
        return_taps=True)
    assert np.array_equal(order, o_order)
    np.testing.assert_allclose(props, o_props, rtol=3e-6, atol=2e-4)
    # stage 2: NMS + padding bit-exact when the oracle runs on the GPU's decoded boxes
    g_rois, _, _, g_keep, g_nkeep = orc.proposal_layer(
        scores.numpy(), deltas.numpy(), im_info.numpy(), anchors.numpy(), stride, pre, post, thresh,
        boxes_override=props, return_taps=True)
    assert np.array_equal(nkeep, g_nkeep)
    assert np.array_equal(rois, g_rois)
    return rois


@pytest.mark.parametrize("tag", ["test", "train", "all"])
def test_proposal_layer_golden(orc, golden, tag):
    stride, pre, post, A = [int(v) for v in golden[f"prop_{tag}_cfg"]]
    t = lambda k: torch.from_numpy(golden[f"prop_{tag}_{k}"])  # noqa: E731
    rois = _proposal_parity(orc, t("scores"), t("deltas"), t("im_info"), t("anchors"), stride, pre, post, 0.7)
    ref = golden[f"prop_{tag}_rois"]  # produced by the reference's own _ProposalLayer
    assert np.array_equal(rois[:, :, 0], ref[:, :, 0])
    assert np.array_equal(rois[:, :, 1:].any(axis=2), ref[:, :, 1:].any(axis=2))
    np.testing.assert_allclose(rois, ref, rtol=3e-6, atol=2e-4)


@pytest.mark.parametrize("cfgname", ["c1_train", "c4_test", "c4_train", "ties"])
def test_proposal_layer_full_size(orc, cfgname):
    from rlobjectdetection_b200.model.rpn.generate_anchors import generate_anchors
    if cfgname == "c1_train":   # VGG-16 600x1000: 37x62, A=9, 12000 -> 2000
        B, H, W, scales, pre, post, imh, imw = 1, 37, 62, [8, 16, 32], 12000, 2000, 600, 1000
    elif cfgname == "c4_test":  # COCO 800x1200: 50x75, A=12, 6000 -> 300
        B, H, W, scales, pre, post, imh, imw = 3, 50, 75, [4, 8, 16, 32], 6000, 300, 800, 1200
    elif cfgname == "c4_train":
        B, H, W, scales, pre, post, imh, imw = 2, 50, 75, [4, 8, 16, 32], 12000, 2000, 800, 1200
    else:                       # heavy score ties: order must be "lower anchor index first"
        B, H, W, scales, pre, post, imh, imw = 2, 20, 30, [8, 16, 32], 1000, 100, 320, 480
    anchors = torch.from_numpy(generate_anchors(scales=np.array(scales), ratios=np.array([0.5, 1, 2]))).float()
    A = anchors.size(0)
    scores, deltas, im_info = syn.rpn_outputs(3, B, A, H, W, imh, imw)
    if cfgname == "ties":
        scores[:, A:] = (scores[:, A:] * 16).floor() / 16  # 16 distinct values over 5400 anchors
    _proposal_parity(orc, scores, deltas, im_info, anchors, 16, pre, post, 0.7)


def test_proposal_module_and_cfg(orc):
    from rlobjectdetection_b200.model.rpn.proposal_layer import _ProposalLayer
    from rlobjectdetection_b200.model.utils.config import cfg
    layer = _ProposalLayer(16, [8, 16, 32], [0.5, 1, 2])
    scores, deltas, im_info = syn.rpn_outputs(5, 2, 9, 14, 21, 224, 336)
    old = (cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N)
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 600, 50
    try:
        rois = layer((cu(scores), cu(deltas), cu(im_info), "TEST"))
    finally:
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = old
    assert rois.shape == (2, 50, 5)
    assert (rois[1, :, 0] == 1).all()  # image index on every row, padding included
    ref = orc.proposal_layer(scores.numpy(), deltas.numpy(), im_info.numpy(),
                             layer._anchors.cpu().numpy(), 16, 600, 50, 0.7)
    np.testing.assert_allclose(rois.cpu().numpy(), ref, rtol=3e-6, atol=2e-4)
    with pytest.raises(NotImplementedError):
        layer((scores, deltas, im_info, "TEST"))


# ------------------------------------------------------------------------------------------
# RL refinement
# ------------------------------------------------------------------------------------------
def test_reward_vs_golden(orc, golden):
    r, l, w = be.action_reward(cu(golden["iou_dt"][None]), cu(golden["iou_gt"][None]), cu(golden["act16"]),
                               crowd=cu(golden["iou_crowd"][None]), mode=be.IOU_COCO, iou_thres=0.0,
                               pos_wratio=2.0, neg_wratio=0.5)
    assert np.array_equal(r[0].cpu().numpy(), golden["reward_out"].astype(np.float32))  # bit-exact
    assert np.array_equal(l[0].cpu().numpy(), golden["reward_label"].astype(np.float32))
    np.testing.assert_allclose(w[0].cpu().numpy(), golden["reward_weight"], rtol=1e-6)


@pytest.mark.parametrize("mode", [be.IOU_COCO, be.IOU_RCNN])
@pytest.mark.parametrize("nact", [16, 56])
def test_reward_c3_size(orc, mode, nact):
    # config 3: 8 images x 300 boxes x 16 (and the reference default 56) actions x 20 gt
    B, N, G = 8, 300, 20
    g = torch.Generator().manual_seed(2)
    boxes = torch.stack([syn.random_boxes(g, N, 800, 1200) for _ in range(B)], 0)
    gt, crowd = syn.gt_boxes(21, B, G, 800, 1200)
    ngt = torch.tensor([20, 0, 7, 20, 1, 20, 13, 20], dtype=torch.int32)
    if mode == be.IOU_COCO:
        boxes, gt = syn.to_xywh(boxes), syn.to_xywh(gt)
    delta = [.5, .25] if nact == 16 else [.5, .25, .125, .0625, .03125, .015625, .008]
    act = orc.action_table(delta)
    cr = crowd if mode == be.IOU_COCO else None
    rr, rl, rw = orc.action_reward(boxes.numpy(), gt.numpy(), act, crowd=None if cr is None else cr.numpy(),
                                   ngt=ngt.numpy(), mode=mode, iou_thres=0.0, pos_wratio=1.5, neg_wratio=0.75)
    r, l, w = be.action_reward(cu(boxes), cu(gt), cu(act), crowd=None if cr is None else cu(cr), ngt=cu(ngt),
                               mode=mode, iou_thres=0.0, pos_wratio=1.5, neg_wratio=0.75)
    assert np.array_equal(r.cpu().numpy(), rr)  # same IEEE ops in the same order: bit-exact
    assert np.array_equal(l.cpu().numpy(), rl)
    np.testing.assert_allclose(w.cpu().numpy(), rw, rtol=1e-6)


@pytest.mark.parametrize("kind", ["identity", "exp_abs", "custom_tensor", "custom_scalar"])
def test_action_wtrans_is_honoured(orc, kind):
    # weight = action.wtrans(delta_iou) * ratio (RL_coco_dataset.py:128-135): Action's default is the
    # identity (action.py:7-10), Config.act_wtrans is exp(|x|) (config.py:48-51), anything else is custom
    import math
    from rlobjectdetection_b200.model.Reinforcement.action import Action, exp_abs
    from rlobjectdetection_b200.model.Reinforcement.reward import action_rewards
    from rlobjectdetection_b200.rl_step import generate_labels
    B, N, G = 2, 40, 5
    g = torch.Generator().manual_seed(77)
    boxes = syn.to_xywh(torch.stack([syn.random_boxes(g, N, 300, 400, 16, 200) for _ in range(B)], 0))
    gt, crowd = syn.gt_boxes(78, B, G, 300, 400)
    gt = syn.to_xywh(gt)
    wt = {"identity": None, "exp_abs": exp_abs, "custom_tensor": lambda x: x * x + 0.5,
          "custom_scalar": lambda x: math.sqrt(math.fabs(x)) + 1.0}[kind]
    action = Action([0.5, 0.25]) if wt is None else Action([0.5, 0.25], wtrans=wt)
    r, l, w = action_rewards(action, cu(boxes), cu(gt), iscrowd=cu(crowd), pos_wratio=1.5, neg_wratio=0.75)
    code = {"identity": orc.WTRANS_IDENTITY, "exp_abs": orc.WTRANS_EXP_ABS}.get(kind)
    rr, rl, rw = orc.action_reward(boxes.numpy(), gt.numpy(), action.actDeltas, crowd=crowd.numpy(), pos_wratio=1.5,
                                   neg_wratio=0.75, wtrans=orc.WTRANS_IDENTITY if code is None else code)
    assert np.array_equal(r.cpu().numpy(), rr) and np.array_equal(l.cpu().numpy(), rl)
    if code is None:  # custom: the callable on the fp32 reward, times the ratio
        f = (lambda x: x * x + 0.5) if kind == "custom_tensor" else np.vectorize(wt)
        rw = (f(rr.astype(np.float64)) * np.where(rl > 0, 1.5, 0.75)).astype(np.float32)
    np.testing.assert_allclose(w.cpu().numpy(), rw, rtol=1e-6, atol=1e-7)
    if kind == "identity":
        assert (w.cpu().numpy() < 0).any()  # identity weights follow the sign of delta_iou
    # the collated label tensor takes the same transform
    det_cat = torch.zeros(B, N, dtype=torch.int32)
    _, labels = generate_labels(action, cu(boxes), cu(torch.zeros(B, N)), cu(det_cat), cu(torch.zeros(B, N)),
                                cu(torch.full((B,), N, dtype=torch.int32)), cu(gt), cu(torch.zeros(B, G, dtype=torch.int32)),
                                iscrowd=cu(crowd), pos_wratio=1.5, neg_wratio=0.75)
    np.testing.assert_allclose(labels[..., 2].cpu().numpy(), rw, rtol=1e-6, atol=1e-7)


def test_reward_float64_boxes_bit_exact(orc):
    # COCO json boxes are float64 (e.g. 123.45 is not an fp32 number): double rows go through in fp64 and match
    # the reference loop bit for bit, labels included, where the fp32-rounded boxes would not
    rng = np.random.default_rng(3)
    B, N, G = 2, 30, 4
    boxes = np.round(np.concatenate([rng.uniform(0, 300, (B, N, 2)), rng.uniform(5, 150, (B, N, 2))], 2), 2)
    gt = np.round(np.concatenate([rng.uniform(0, 300, (B, G, 2)), rng.uniform(5, 150, (B, G, 2))], 2), 2)
    crowd = (rng.uniform(size=(B, G)) < 0.25).astype(np.uint8)
    act = orc.action_table([0.5, 0.25])
    rr, rl, rw = orc.action_reward_f64(boxes, gt, act, crowd=crowd, pos_wratio=2.0, neg_wratio=0.5)
    r, l, w = be.action_reward(cu(boxes), cu(gt), cu(act), crowd=cu(crowd), mode=be.IOU_COCO, pos_wratio=2.0,
                               neg_wratio=0.5)
    assert np.array_equal(r.cpu().numpy(), rr) and np.array_equal(l.cpu().numpy(), rl)
    np.testing.assert_allclose(w.cpu().numpy(), rw, rtol=2e-7)
    lab = be.rl_labels(cu(boxes), cu(gt), cu(act), crowd=cu(crowd), pos_wratio=2.0, neg_wratio=0.5).cpu().numpy()
    assert np.array_equal(lab[..., 1], rl)
    np.testing.assert_allclose(lab[..., 2], rw, rtol=2e-7)
    with pytest.raises(RuntimeError):  # fp64 rows exist in COCO mode only
        be.lib()  # keep the library loaded
        from rlobjectdetection_b200.model._backend import check, ptr, stream_of
        t = cu(boxes)
        check(be.lib().rlod_action_reward(ptr(t), ptr(cu(gt)), 1, None, None, ptr(cu(act)), B, N, 16, G, be.IOU_RCNN,
                                          1, 0.0, 1.0, 1.0, ptr(torch.empty(B, N, 16, device=DEV)), None, None,
                                          stream_of(t)), "rlod_action_reward")


def test_move_from_act_tied_preds(orc, golden):
    # saturated / equal predictions: the visit order is np.flip(np.argsort(pred)) = higher flat index first
    # among equals; pinned on the reference's own move_from_act (tests/golden/make_golden.py, tied cases)
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    act = Action([0.5, 0.25])
    for tag in ("tie_small", "tie_large", "tie_all"):  # tie_all: the reference's unpinned run (all preds equal)
        for k in (1, 3):
            bb = golden[f"{tag}_boxes"].copy()
            moved, prec = act.move_from_act(bb, golden[f"{tag}_preds"], golden[f"{tag}_targets"], k, device=DEV)
            assert np.array_equal(moved, golden[f"{tag}_k{k}_out"]), (tag, k)
            assert prec == float(golden[f"{tag}_k{k}_prec"])
            rb, rp = orc.move_from_act(golden[f"{tag}_boxes"], golden[f"{tag}_preds"], golden[f"{tag}_targets"],
                                       golden["act16"], k)
            assert np.array_equal(rb, moved) and rp == prec


def test_move_from_act_vs_golden(orc, golden):
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    act = Action([0.5, 0.25])
    for k in (1, 5, 40, 1000):
        bb = golden["move_in_boxes"].copy()
        moved, prec = act.move_from_act(bb, golden["move_preds"], golden["move_targets"], k, device=DEV)
        rb, rp = orc.move_from_act(golden["move_in_boxes"], golden["move_preds"], golden["move_targets"],
                                   golden["act16"], k)
        assert np.array_equal(moved, rb) and prec == rp
        if k in (1, 5):  # the reference's own outputs
            assert np.array_equal(moved, golden[f"move_k{k}_boxes"]) and prec == float(golden[f"move_k{k}_prec"])


def test_move_from_act_corners_matches_reward():
    # refining an x1y1x2y2 roi tensor in place: the moved box's IoU gain equals the reward
    B, N, G = 2, 64, 5
    g = torch.Generator().manual_seed(31)
    boxes = torch.stack([syn.random_boxes(g, N, 600, 1000) for _ in range(B)], 0)
    gt, _ = syn.gt_boxes(32, B, G, 600, 1000)
    rois = torch.cat([torch.arange(B).float()[:, None, None].expand(B, N, 1), boxes], 2).contiguous()
    act = cu(np.asarray([[0.5, 0, 0, 0], [-0.5, 0, 0, 0], [0, 0.25, 0, 0], [0, 0, 0.25, 0], [0, 0, 0, -0.25]],
                        dtype=np.float32))
    r, l, _ = be.action_reward(cu(boxes), cu(gt), act, mode=be.IOU_RCNN)
    refined = cu(rois).clone()
    moved = be.move_from_act(refined, r, l, act, N, corners=True)
    zero_act = torch.zeros(1, 4, device=DEV)
    before = be.action_reward(cu(boxes), cu(gt), zero_act, mode=be.IOU_RCNN, want_labels=False)
    # IoU of refined boxes vs gt, via a zero action on the refined boxes: reward 0, so use overlaps
    from rlobjectdetection_b200.model.rpn.bbox_transform import bbox_overlaps_batch
    iou_new = bbox_overlaps_batch(refined[:, :, 1:5].contiguous(), cu(gt)).max(dim=2).values
    iou_old = bbox_overlaps_batch(cu(boxes), cu(gt)).max(dim=2).values
    best = r.max(dim=2).values
    gain = torch.where(best > 0, best, torch.zeros_like(best))
    assert torch.equal(iou_new - iou_old, gain)
    assert int(moved.item()) == int((best > 0).sum().item())
    assert (before == 0).all()
    assert torch.equal(refined[:, :, 0], cu(rois)[:, :, 0])


@pytest.mark.parametrize("pool", ["avg", "max", "none"])
@pytest.mark.parametrize("shape", [(2, 64, 25, 38), (3, 32, 50, 75), (2, 16, 13, 21)])
def test_roi_align_channels_last_input(orc, pool, shape):
    # a channels-last (NHWC) feature map is read in place by the plane kernel: same values as the NCHW run,
    # bit for bit (same taps, same arithmetic), and within 1e-5 of the oracle
    B, C, H, W = shape
    g = torch.Generator().manual_seed(61)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(62, B, 40, H * 16.0, W * 16.0)
    mode = {"avg": be.POOL_AVG, "max": be.POOL_MAX, "none": be.POOL_NONE}[pool]
    a = 8 if pool == "none" else 7
    nchw = be.roi_align_forward(cu(feat), cu(rois), a, a, 1 / 16.0, mode)
    f_cl = cu(feat).contiguous(memory_format=torch.channels_last)
    assert not f_cl.is_contiguous()
    nhwc = be.roi_align_forward(f_cl, cu(rois), a, a, 1 / 16.0, mode)
    assert torch.equal(nchw, nhwc)
    ref = orc.roi_align(feat.numpy(), rois.numpy(), a, a, 1 / 16.0, pool_mode={"avg": orc.POOL_AVG, "max": orc.POOL_MAX,
                                                                               "none": orc.POOL_NONE}[pool])
    close(nhwc.cpu().numpy(), ref)
    # the module surface takes it too, forward and backward (the gradient comes back NCHW-dense)
    from rlobjectdetection_b200.model.roi_align.modules.roi_align import RoIAlignAvg
    if pool == "avg":
        x = f_cl.clone().requires_grad_(True)
        y = RoIAlignAvg(7, 7, 1 / 16.0)(x, cu(rois))
        y.backward(torch.ones_like(y))
        x2 = cu(feat).clone().requires_grad_(True)
        y2 = RoIAlignAvg(7, 7, 1 / 16.0)(x2, cu(rois))
        y2.backward(torch.ones_like(y2))
        assert torch.equal(y, y2)
        close(x.grad.cpu().numpy(), x2.grad.cpu().numpy())


def test_roi_pool_channels_last_input(orc):
    B, C, H, W = 2, 64, 38, 63
    g = torch.Generator().manual_seed(63)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(64, B, 64, H * 16.0, W * 16.0)
    o1, a1 = be.roi_pool_forward(cu(feat), cu(rois), 7, 7, 1 / 16.0)
    o2, a2 = be.roi_pool_forward(cu(feat).contiguous(memory_format=torch.channels_last), cu(rois), 7, 7, 1 / 16.0)
    assert torch.equal(o1, o2) and torch.equal(a1, a2)   # argmax keeps its NCHW meaning
    ro, ra = orc.roi_pool(feat.numpy(), rois.numpy(), 7, 7, 1 / 16.0)
    assert np.array_equal(o2.cpu().numpy(), ro) and np.array_equal(a2.cpu().numpy(), ra)


def test_feature_layouts_are_not_silently_copied():
    feat = torch.randn(2, 8, 50, 75, device=DEV)
    rois = cu(syn.rois_for_batch(65, 2, 16, 800.0, 1200.0))
    with pytest.raises(ValueError, match="dense NCHW or channels-last"):
        be.roi_align_forward(feat[:, :, ::2], rois, 7, 7, 1 / 16.0, be.POOL_AVG)       # strided view
    with pytest.raises(TypeError):
        be.roi_align_forward(feat.double(), rois, 7, 7, 1 / 16.0, be.POOL_AVG)
    # channels-last + a shape the plane kernel does not cover (C % 4 != 0): refused, not copied
    odd = torch.randn(2, 6, 20, 30, device=DEV).contiguous(memory_format=torch.channels_last)
    with pytest.raises(RuntimeError, match="unsupported|not supported"):
        be.roi_align_forward(odd, rois, 7, 7, 1 / 16.0, be.POOL_AVG)


@pytest.mark.parametrize("nact,ties", [(16, False), (56, False), (16, True)])
def test_reward_refine_equals_unfused(orc, nact, ties):
    # rlod_reward_refine == rlod_action_reward(RCNN) -> copy -> rlod_move_from_act(maxk=N, rewards as preds) -> pack,
    # bit for bit (also with heavily tied rewards: coarse integer boxes), and == the oracle's restatement
    B, N, G = 3, 300, 20
    g = torch.Generator().manual_seed(51)
    boxes = torch.stack([syn.random_boxes(g, N, 800, 1200) for _ in range(B)], 0)
    gt, _ = syn.gt_boxes(52, B, G, 800, 1200)
    if ties:
        boxes, gt = (boxes / 64).round() * 64, (gt / 64).round() * 64
    ngt = torch.tensor([20, 0, 7], dtype=torch.int32)
    rois = torch.cat([torch.arange(B).float()[:, None, None].expand(B, N, 1), boxes], 2).contiguous()
    delta = [.5, .25] if nact == 16 else [.5, .25, .125, .0625, .03125, .015625, .008]
    act = cu(orc.action_table(delta))
    r, l, w = be.action_reward(cu(boxes), cu(gt), act, ngt=cu(ngt), mode=be.IOU_RCNN, pos_wratio=1.5, neg_wratio=0.75)
    refined = cu(rois).clone()
    moved = be.move_from_act(refined, r, l, act, N, corners=True)
    f = be.reward_refine(cu(rois), cu(gt), act, ngt=cu(ngt), pos_wratio=1.5, neg_wratio=0.75, first_image=7)
    for name, ref in (("reward", r), ("label", l), ("weight", w), ("refined", refined)):
        assert torch.equal(f[name], ref), name
    assert int(f["moved"].item()) == int(moved.item())
    packed = torch.cat([refined, r], 2)
    packed[:, :, 0] += 7
    assert torch.equal(f["packed"], packed)
    o_ref, o_moved = orc.refine_best_action(rois.numpy(), r.cpu().numpy(), l.cpu().numpy(), act.cpu().numpy())
    assert np.array_equal(f["refined"].cpu().numpy(), o_ref) and o_moved == int(moved.item())
    if ties:
        rr = r.cpu().numpy()
        assert ((rr == rr.max(axis=2, keepdims=True)).sum(axis=2) > 1).any()  # the tie rule is exercised
    only = be.reward_refine(cu(rois), cu(gt), act, ngt=cu(ngt), want=("packed",))
    assert list(only) == ["packed"] and torch.equal(only["packed"][:, :, 1:], packed[:, :, 1:])


# ------------------------------------------------------------------------------------------
# whole step
# ------------------------------------------------------------------------------------------
def test_c4_bench_workload_matches_oracle(orc):
    """The EXACT workload bench.py's headline number is quoted on -- bench.make_inputs(100, 24): 24 images,
    50x75x1024 features, 45 000 anchors each, TEST cfg 6000 -> 300, 7 200 rois, 16 actions x 20 gt -- through
    DetectRefineStep and against the CPU oracle: sort order / keep decisions / rewards / labels / refined boxes
    bit-exact (the oracle's NMS and rewards run on the GPU's decoded boxes: device expf vs libm differ in the last
    ulp, bounded separately at 3e-6), pooled features of both RoIAlign passes within 1e-5."""
    import bench
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    from rlobjectdetection_b200.model.utils.config import cfg
    inputs = bench.make_inputs(100, bench.IMAGES)
    scores, deltas, im_info, feat, gt = inputs
    step = DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA, backward=False,
                            outputs=("rois", "reward", "label", "refined", "moved", "packed"), first_image=0)
    old = (cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH)
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
    try:
        out = step(*(cu(t) for t in inputs))
        torch.cuda.synchronize()
    finally:
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = old
    anchors = step.proposal._anchors.cpu()
    rois = _proposal_parity(orc, scores, deltas, im_info, anchors, bench.STRIDE, bench.PRE, bench.POST, bench.NMS_T)
    assert np.array_equal(out["rois"].cpu().numpy(), rois)           # the step's rois are the layer's
    assert rois.shape == (24, 300, 5) and (rois[:, :, 1:].any(axis=2)).all()  # 300 kept per image, none padded
    act = step.action.actDeltas
    rr, rl, _ = orc.action_reward(rois[:, :, 1:5], gt.numpy(), act, mode=orc.MODE_RCNN)
    assert np.array_equal(out["reward"].cpu().numpy(), rr) and np.array_equal(out["label"].cpu().numpy(), rl)
    refined, moved = orc.refine_best_action(rois, rr, rl, act)
    assert np.array_equal(out["refined"].cpu().numpy(), refined) and int(out["moved"].item()) == moved
    assert np.array_equal(out["packed"].cpu().numpy(), np.concatenate([refined, rr], 2))
    for name, boxes in (("pooled", rois), ("pooled_refined", refined)):
        got = out[name].cpu().numpy()
        del out[name]
        ref = orc.roi_align(feat.numpy(), boxes.reshape(-1, 5), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
        close(got, ref, what=name)
        del got, ref


@pytest.mark.parametrize("pipelined", [False, True])
def test_graphed_step_equals_eager(pipelined):
    # the CUDA-graph form of the step (serial, and pipelined over the next step's inputs) against eager launches
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    B, C, H, W, G = 3, 32, 25, 38, 6
    step = DetectRefineStep(cfg_key="TEST", backward=False, outputs=("rois", "reward", "refined", "packed", "moved"),
                            first_image=5)
    A = step.proposal._num_anchors
    g = torch.Generator().manual_seed(12)
    feat = cu(torch.randn(B, C, H, W, generator=g))
    scores, deltas, im_info = (cu(t) for t in syn.rpn_outputs(20, B, A, H, W, H * 16, W * 16))
    gt = cu(syn.gt_boxes(40, B, G, H * 16, W * 16)[0])
    keys = ("rois", "reward", "refined", "packed", "moved", "pooled", "pooled_refined")
    ref = {k: v.clone() for k, v in step(scores, deltas, im_info, feat, gt).items()}
    gs = step.capture(scores, deltas, im_info, feat, gt, next_inputs=(scores, deltas, im_info, gt) if pipelined else None)
    gs.prime()
    for it in range(3):
        out = gs.replay()
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(out[k], ref[k]), (k, it)
    # new contents in the same buffers: the replay follows them (pipelined: one step later, or after prime())
    s2, d2, _ = syn.rpn_outputs(21, B, A, H, W, H * 16, W * 16)
    scores.copy_(cu(s2)), deltas.copy_(cu(d2))
    ref2 = {k: v.clone() for k, v in step(scores, deltas, im_info, feat, gt).items()}
    assert not torch.equal(ref2["rois"], ref["rois"])
    if pipelined:
        gs.replay()   # pools the old rois, prepares the new ones
    out = gs.replay()
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(out[k], ref2[k]), k

@pytest.mark.parametrize("graphed", [False, True])
def test_step_merged_repool_equals_separate(graphed):
    """repool="merged" (one RoIAlign call pools the proposals and the refined boxes) == repool="separate" (two
    calls), bit for bit, eager and graphed, with and without the step-ahead light work."""
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    B, C, H, W, G = 3, 32, 25, 38, 6
    outs = ("rois", "reward", "label", "refined", "packed", "moved")
    mk = lambda form: DetectRefineStep(cfg_key="TEST", backward=False, outputs=outs, first_image=2, repool=form)  # noqa: E731
    merged, separate = mk("merged"), mk("separate")
    A = merged.proposal._num_anchors
    g = torch.Generator().manual_seed(13)
    feat = cu(torch.randn(B, C, H, W, generator=g))
    scores, deltas, im_info = (cu(t) for t in syn.rpn_outputs(23, B, A, H, W, H * 16, W * 16))
    gt = cu(syn.gt_boxes(41, B, G, H * 16, W * 16)[0])
    keys = outs + ("pooled", "pooled_refined")
    ref = {k: v.clone() for k, v in separate(scores, deltas, im_info, feat, gt).items()}
    assert not torch.equal(ref["pooled"], ref["pooled_refined"])
    if graphed:
        gs = merged.capture(scores, deltas, im_info, feat, gt, next_inputs=(scores, deltas, im_info, gt))
        gs.prime()
        for _ in range(2):
            out = gs.replay()
    else:
        nxt = (scores, deltas, im_info, gt)
        merged(scores, deltas, im_info, feat, gt, inputs_ready=True, next_inputs=nxt, next_ready=True)
        out = merged(scores, deltas, im_info, feat, gt, inputs_ready=True, next_inputs=nxt, next_ready=True)
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(out[k], ref[k]), k
    with pytest.raises(ValueError):
        mk("both")


def test_hotpath_step_matches_oracle_chain(orc):
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    B, C, H, W, G = 2, 16, 25, 38, 6
    step = DetectRefineStep(cfg_key="TEST", backward=True)
    A = step.proposal._num_anchors
    scores, deltas, im_info = syn.rpn_outputs(9, B, A, H, W, H * 16, W * 16)
    g = torch.Generator().manual_seed(10)
    feat = torch.randn(B, C, H, W, generator=g)
    gt, _ = syn.gt_boxes(11, B, G, H * 16, W * 16)
    gp = torch.randn(B * 300, C, 7, 7, generator=g)
    out = step(cu(scores), cu(deltas), cu(im_info), cu(feat), cu(gt), cu(gp))
    rois = out["rois"].cpu().numpy()
    pooled_ref = orc.roi_align(feat.numpy(), rois.reshape(-1, 5), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG)
    close(out["pooled"].cpu().numpy(), pooled_ref)
    rr, rl, _ = orc.action_reward(rois[:, :, 1:5], gt.numpy(), step.action.actDeltas, mode=orc.MODE_RCNN)
    assert np.array_equal(out["reward"].cpu().numpy(), rr) and np.array_equal(out["label"].cpu().numpy(), rl)
    refined = out["refined"].cpu().numpy()
    close(out["pooled_refined"].cpu().numpy(),
          orc.roi_align(feat.numpy(), refined.reshape(-1, 5), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG))
    close(out["grad_feat"].cpu().numpy(),
          orc.roi_align_bwd(gp.numpy(), feat.numpy(), refined.reshape(-1, 5), 7, 7, 1 / 16.0, pool_mode=orc.POOL_AVG))


def test_hotpath_step_run_ahead_streams():
    """inputs_ready=True / an Event lets the step's light stream run ahead of the caller's stream: many
    calls back to back on changing inputs, with and without run-ahead, must give identical results
    (the cross-stream tensors are event-ordered and record_stream'ed), also under allocator churn."""
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    B, C, H, W, G = 3, 32, 25, 38, 6
    step = DetectRefineStep(cfg_key="TEST", backward=False)
    A = step.proposal._num_anchors
    g = torch.Generator().manual_seed(12)
    feat = cu(torch.randn(B, C, H, W, generator=g))
    ins = []
    for k in range(6):
        scores, deltas, im_info = syn.rpn_outputs(20 + k, B, A, H, W, H * 16, W * 16)
        gt, _ = syn.gt_boxes(40 + k, B, G, H * 16, W * 16)
        ins.append((cu(scores), cu(deltas), cu(im_info), feat, cu(gt)))
    torch.cuda.synchronize()
    keys = ("rois", "reward", "refined", "pooled", "pooled_refined")
    ref = []
    for a in ins:
        o = step(*a)                       # serial: the light stream waits for the caller's stream
        ref.append({k: o[k].clone() for k in keys})
    torch.cuda.synchronize()
    for mode in ("resident", "event", "prefetch", "prefetch_wrong"):
        outs = []
        for i, a in enumerate(ins):
            if mode == "event":
                ev = torch.cuda.Event()
                ev.record()
                o = step(*a, inputs_ready=ev)
            elif mode.startswith("prefetch"):
                # announce the next step's inputs (or, "wrong", other ones: they must be ignored)
                n = ins[(i + (1 if mode == "prefetch" else 3)) % len(ins)]
                o = step(*a, inputs_ready=True, next_inputs=(n[0], n[1], n[2], n[4]), next_ready=True)
            else:
                o = step(*a, inputs_ready=True)
            outs.append({k: o[k] for k in keys})
            del o
            junk = torch.empty(1 << 20, device=DEV).normal_()   # allocator churn on the caller's stream
            del junk
        torch.cuda.synchronize()
        for r, o in zip(ref, outs):
            for k in keys:
                assert torch.equal(r[k], o[k]), (mode, k)



# ------------------------------------------------------------------------------------------
# test-time post-processing (f1): threshold / decode / clip / sort / per-class NMS / cap
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["c21", "c9", "agn"])
def test_detect_postprocess_golden(orc, tag):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_detect.npz"))
    N, K, agn, cap = (int(v) for v in g[f"{tag}_cfg"])
    dets, counts = be.detect_postprocess(cu(g[f"{tag}_rois"]), cu(g[f"{tag}_cls_prob"]), cu(g[f"{tag}_bbox_pred"]),
                                         cu(g[f"{tag}_im_info"]), thresh=float(g[f"{tag}_thresh"]), nms_thresh=0.3,
                                         max_per_image=cap, stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0),
                                         class_agnostic=bool(agn))
    c = counts.cpu().numpy()[0]
    assert list(c) == list(g[f"{tag}_counts"])  # keep decisions bit-exact
    d = dets.cpu().numpy()[0]
    got = np.concatenate([d[j, :c[j]] for j in range(K)], 0)
    ref = g[f"{tag}_dets"]
    assert np.array_equal(got[:, 4], ref[:, 4])
    np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=3e-6, atol=2e-4)


def test_detect_postprocess_c5_size(orc):
    # config 5: 64 images x 81 classes x 300 boxes, thr 0.3: keep lists vs the oracle
    from rlobjectdetection_b200.detections import postprocess_detections, to_all_boxes
    B, N, K = 64, 300, 81
    g = torch.Generator().manual_seed(77)
    rois = torch.cat([torch.arange(B).float()[:, None, None].expand(B, N, 1),
                      torch.stack([syn.random_boxes(g, N, 600, 1000, 24.0, 300.0) for _ in range(B)], 0)], 2).contiguous()
    cls_prob = torch.softmax(torch.randn(B, N, K, generator=g) * 3.0, 2).contiguous()
    bbox_pred = (torch.randn(B, N, 4 * K, generator=g) * 0.5).contiguous()
    im_info = torch.tensor([[600.0, 1000.0, 1.6]]).repeat(B, 1)
    dets, counts = postprocess_detections(cu(rois), cu(cls_prob), cu(bbox_pred), cu(im_info), thresh=0.05)
    ab = to_all_boxes(dets, counts)
    sub = [0, 17, 63]
    ref = orc.detect_postprocess(rois[sub].numpy(), cls_prob[sub].numpy(), bbox_pred[sub].numpy(), im_info[sub].numpy(),
                                 thresh=0.05, nms_thresh=0.3, max_per_image=100, stds=(0.1, 0.1, 0.2, 0.2),
                                 means=(0.0, 0.0, 0.0, 0.0))
    for k, b in enumerate(sub):
        for j in range(K):
            assert ab[j][b].shape == ref[k][j].shape, (b, j)
            assert np.array_equal(ab[j][b][:, 4], ref[k][j][:, 4])
            np.testing.assert_allclose(ab[j][b][:, :4], ref[k][j][:, :4], rtol=3e-6, atol=2e-4)
    c = counts.cpu().numpy()
    assert (c[:, 0] == 0).all() and (c.sum(1) >= 100).all()  # the cap keeps ties with >=


# ------------------------------------------------------------------------------------------
# RL batch generation + eval glue (f2)
# ------------------------------------------------------------------------------------------
def test_rl_generate_labels_vs_reference_collate(orc):
    import os
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    from rlobjectdetection_b200.rl_step import generate_labels
    from math import exp, fabs
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_rl.npz"))
    # the golden was made with the reference's Config.act_wtrans, passed the way a user of the
    # reference passes it: as a plain callable (recognised by probing, action.wtrans_code)
    action = Action([0.5, 0.25], wtrans=lambda x: exp(fabs(x)))
    bboxes, labels = generate_labels(action, cu(g["dets"]), cu(g["det_score"]), cu(g["det_cat"]), cu(g["det_img"]),
                                     cu(g["ndet"]), cu(g["gt"]), cu(g["gt_cat"]), iscrowd=cu(g["crowd"]), ngt=cu(g["ngt"]),
                                     pos_wratio=float(g["wratio"][0]), neg_wratio=float(g["wratio"][1]))
    lab, ref = labels.cpu().numpy(), g["padded_labels"]
    assert np.array_equal(lab[..., :2], ref[..., :2])                      # act ids and +-1 labels bit-exact
    np.testing.assert_allclose(lab[..., 2], ref[..., 2], rtol=2e-7)        # fp64 weight rounded to fp32
    np.testing.assert_allclose(bboxes.cpu().numpy(), g["padded_bboxes"], rtol=0, atol=0)


def test_rl_labels_c3_size(orc):
    # config 3: 8 images x 300 boxes x 16 actions x 20 gt, 5 categories
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    B, N, G = 8, 300, 20
    g = torch.Generator().manual_seed(31)
    dets = syn.to_xywh(torch.stack([syn.random_boxes(g, N, 800, 1200) for _ in range(B)], 0))
    gt, crowd = syn.gt_boxes(32, B, G, 800, 1200)
    gt = syn.to_xywh(gt)
    det_cat = torch.randint(0, 5, (B, N), generator=g, dtype=torch.int32)
    gt_cat = torch.randint(0, 5, (B, G), generator=g, dtype=torch.int32)
    ndet = torch.tensor([300, 299, 0, 150, 300, 1, 300, 300], dtype=torch.int32)
    ngt = torch.tensor([20, 0, 5, 20, 20, 20, 3, 20], dtype=torch.int32)
    act = Action([0.5, 0.25])
    lab = be.rl_labels(cu(dets), cu(gt), act.table(DEV), det_cat=cu(det_cat), ndet=cu(ndet), gt_cat=cu(gt_cat),
                       crowd=cu(crowd), ngt=cu(ngt), pos_wratio=1.5, neg_wratio=0.5).cpu().numpy()
    sub = [0, 1, 2, 5]
    ref = orc.rl_labels(dets[sub].numpy(), det_cat[sub].numpy(), ndet[sub].numpy(), gt[sub].numpy(), gt_cat[sub].numpy(),
                        crowd[sub].numpy(), ngt[sub].numpy(), act.actDeltas, 0.0, 1.5, 0.5)
    assert np.array_equal(lab[sub][..., :2], ref[..., :2])
    np.testing.assert_allclose(lab[sub][..., 2], ref[..., 2], rtol=2e-7)


def test_rl_eval_step_matches_reference_sequence(orc, golden):
    # trainval_net.py:202-221 = xyxy -> xywh, Action.move_from_act(maxk=1), / scale; move_from_act itself
    # is pinned by the golden vectors of the reference's Action (test_move_from_act_vs_golden)
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    from rlobjectdetection_b200.rl_step import eval_step
    act = Action([0.5, 0.25])
    xywh = golden["move_in_boxes"]
    b, n, _ = xywh.shape
    rows = np.zeros((b, n, 8), np.float32)
    rows[:, :, 0] = np.arange(b)[:, None]
    rows[:, :, 1:3] = xywh[:, :, :2]
    rows[:, :, 3:5] = xywh[:, :, :2] + xywh[:, :, 2:]
    scale = np.array([1.0, 1.6, 0.5], np.float32)
    out, moved = eval_step(act, cu(rows), cu(golden["move_preds"]), cu(golden["move_targets"]), cu(scale), maxk=1)
    ref_in = rows[:, :, 1:5].copy()
    ref_in[:, :, 2] -= ref_in[:, :, 0]
    ref_in[:, :, 3] -= ref_in[:, :, 1]
    ref, prec = orc.move_from_act(ref_in, golden["move_preds"], golden["move_targets"], act.actDeltas, 1)
    np.testing.assert_allclose(out.cpu().numpy()[:, :, 1:5], ref / scale[:, None, None], rtol=1e-6, atol=1e-5)
    assert int(moved.item()) == int(round(prec * b * 1 / 100.0))


# ------------------------------------------------------------------------------------------
# RoICrop, POOLING_MODE 'crop' (f4)
# ------------------------------------------------------------------------------------------
def _crop_case(seed, B, C, H, W, n_per, gs):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, C, H, W, generator=g)
    rois = syn.rois_for_batch(seed + 1, B, n_per, H * 16.0, W * 16.0)
    return feat, rois


@pytest.mark.parametrize("case", [dict(B=2, C=6, H=20, W=31, n_per=9, gs=14), dict(B=3, C=16, H=38, W=63, n_per=20, gs=14),
                                  dict(B=1, C=3, H=5, W=6, n_per=8, gs=7)])
def test_roi_crop_forward_backward(orc, case):
    from rlobjectdetection_b200.model.utils.net_utils import _affine_grid_gen
    feat, rois = _crop_case(8, case["B"], case["C"], case["H"], case["W"], case["n_per"], case["gs"])
    H, W, gs = case["H"], case["W"], case["gs"]
    for ac in (True, False):
        grid = _affine_grid_gen(cu(rois), (H, W), gs, align_corners=ac).cpu().numpy()
        np.testing.assert_allclose(grid, orc.affine_grid(rois.numpy(), H, W, gs, ac), rtol=0, atol=2e-6)
    grid_xy = orc.affine_grid(rois.numpy(), H, W, gs, True)
    grid_yx = np.ascontiguousarray(np.stack([grid_xy[..., 1], grid_xy[..., 0]], 3))
    out = be.roi_crop_forward(cu(feat), cu(grid_yx)).cpu().numpy()
    ref = orc.roi_crop(feat.numpy(), grid_yx)
    close(out, ref, what="roi_crop fwd")
    g = torch.Generator().manual_seed(5)
    gout = torch.randn(*out.shape, generator=g)
    gin = be.roi_crop_backward(cu(gout), cu(grid_yx), tuple(feat.shape)).cpu().numpy()
    close(gin, orc.roi_crop_bwd(gout.numpy(), grid_yx, tuple(feat.shape)), what="roi_crop bwd")


def test_roi_crop_backward_general_grids(orc):
    """The backward gathers per patch pixel when a roi's grid is separable and monotone (what
    _affine_grid_gen emits) and scatters otherwise: jittered (non-separable), mirrored (x2 < x1:
    decreasing columns), far-outside and mixed batches must all match the oracle's scatter."""
    B, C, H, W, n_per, gs = 2, 40, 20, 31, 12, 14
    feat, rois = _crop_case(21, B, C, H, W, n_per, gs)
    rois = rois.clone()
    rois[1, [1, 3]] = rois[1, [3, 1]]            # mirrored in x
    rois[2, [2, 4]] = rois[2, [4, 2]]            # mirrored in y
    rois[3, 1:5] = torch.tensor([-300.0, -200.0, -120.0, -90.0])   # every tap outside the map
    rois[4, 1:5] = torch.tensor([-40.0, -30.0, 60.0, 50.0])        # straddles the top-left corner
    rois[5, 1:5] = torch.tensor([W * 16.0 - 50, H * 16.0 - 40, W * 16.0 + 80, H * 16.0 + 90])
    grid_xy = orc.affine_grid(rois.numpy(), H, W, gs, True)
    grid_yx = np.ascontiguousarray(np.stack([grid_xy[..., 1], grid_xy[..., 0]], 3)).astype(np.float32)
    rng = np.random.default_rng(3)
    grid_yx[6] += rng.normal(0, 0.02, grid_yx[6].shape).astype(np.float32)       # not separable
    grid_yx[7, 3, 5, 1] = np.nextafter(grid_yx[7, 3, 5, 1], np.float32(2.0))      # one sample off by an ulp
    g = torch.Generator().manual_seed(6)
    gout = torch.randn(rois.size(0), C, gs, gs, generator=g)
    gin = be.roi_crop_backward(cu(gout), cu(grid_yx), tuple(feat.shape)).cpu().numpy()
    close(gin, orc.roi_crop_bwd(gout.numpy(), grid_yx, tuple(feat.shape)), what="roi_crop bwd general")
    # accumulate: adds to the caller's gradient
    base = torch.randn(*feat.shape, generator=g)
    acc = cu(base).clone()
    be.roi_crop_backward(cu(gout), cu(grid_yx), tuple(feat.shape), grad_in=acc)
    close(acc.cpu().numpy(), base.numpy() + gin, what="roi_crop bwd accumulate")


def test_roi_crop_forward_plane_kernel_general_grids(orc):
    """C % 4 == 0, equal rois per image and gh * gw % 4 == 0 take the plane-resident forward (k_roi_crop_planes):
    zero-framed planes instead of per-tap masks.  Mirrored, far-outside, corner-straddling, non-separable and
    exactly-on-the-border grids must equal the oracle's masked taps (within 1e-5: the plane kernel sums the taps
    with FMAs, as the reference's own CUDA build does; exact zeros where no tap is inside)."""
    B, C, H, W, n_per, gs = 2, 40, 20, 31, 12, 14
    feat, rois = _crop_case(31, B, C, H, W, n_per, gs)
    rois = rois.clone()
    rois[1, [1, 3]] = rois[1, [3, 1]]
    rois[2, [2, 4]] = rois[2, [4, 2]]
    rois[3, 1:5] = torch.tensor([-300.0, -200.0, -120.0, -90.0])
    rois[4, 1:5] = torch.tensor([-40.0, -30.0, 60.0, 50.0])
    rois[5, 1:5] = torch.tensor([W * 16.0 - 50, H * 16.0 - 40, W * 16.0 + 80, H * 16.0 + 90])
    rois[8, 1:5] = torch.tensor([0.0, 0.0, W * 16.0, H * 16.0])    # the whole map: samples on both borders
    grid_xy = orc.affine_grid(rois.numpy(), H, W, gs, True)
    grid_yx = np.ascontiguousarray(np.stack([grid_xy[..., 1], grid_xy[..., 0]], 3)).astype(np.float32)
    rng = np.random.default_rng(4)
    grid_yx[6] += rng.normal(0, 0.02, grid_yx[6].shape).astype(np.float32)
    grid_yx[9, 0, 0] = (-1.0, -1.0)                                # exactly the first pixel
    grid_yx[9, 0, 1] = (1.0, 1.0)                                  # exactly the last pixel
    grid_yx[9, 0, 2] = (1e30, -1e30)                               # saturating coordinates
    grid_yx[9, 0, 3] = (np.nextafter(np.float32(-1.0), np.float32(-2.0)), 0.0)   # one ulp above the map
    out = be.roi_crop_forward(cu(feat), cu(grid_yx)).cpu().numpy()
    ref = orc.roi_crop(feat.numpy(), grid_yx)
    close(out, ref, what="roi_crop fwd planes")
    assert not out[3].any() and not out[9, :, 0, 2].any()          # no tap inside: exact zeros
    np.testing.assert_array_equal(out[9, :, 0, 0], feat[0, :, 0, 0].numpy())      # weight 1 on one pixel: exact
    np.testing.assert_array_equal(out[9, :, 0, 1], feat[0, :, H - 1, W - 1].numpy())
    # the same call with an odd channel count takes the gather kernel, 36 channels the 4-channel planes
    out39 = be.roi_crop_forward(cu(feat[:, :39].contiguous()), cu(grid_yx)).cpu().numpy()
    close(out39, ref[:, :39], what="roi_crop fwd gather")
    out36 = be.roi_crop_forward(cu(feat[:, :36].contiguous()), cu(grid_yx)).cpu().numpy()
    np.testing.assert_array_equal(out36, out[:, :36])               # G = 1 and G = 2: the same arithmetic
    # fewer rois per image than warps, and a single image
    feat1, rois1 = _crop_case(32, 1, 4, 9, 11, 3, 6)
    g1 = orc.affine_grid(rois1.numpy(), 9, 11, 6, True)
    g1 = np.ascontiguousarray(np.stack([g1[..., 1], g1[..., 0]], 3)).astype(np.float32)
    close(be.roi_crop_forward(cu(feat1), cu(g1)).cpu().numpy(), orc.roi_crop(feat1.numpy(), g1), what="roi_crop fwd small")
    # more rois than warps with a point count that is no multiple of 32 nor of 4 (5 x 5), rois split unevenly over warps
    feat2, rois2 = _crop_case(33, 2, 8, 12, 17, 21, 5)
    g2 = orc.affine_grid(rois2.numpy(), 12, 17, 5, True)
    g2 = np.ascontiguousarray(np.stack([g2[..., 1], g2[..., 0]], 3)).astype(np.float32)
    close(be.roi_crop_forward(cu(feat2), cu(g2)).cpu().numpy(), orc.roi_crop(feat2.numpy(), g2), what="roi_crop fwd 5x5")


def test_roi_crop_module_and_crop_pool(orc):
    from rlobjectdetection_b200.model.roi_crop.modules.roi_crop import _RoICrop
    from rlobjectdetection_b200.model.utils.net_utils import crop_pool
    feat, rois = _crop_case(9, 2, 8, 20, 31, 10, 14)
    f = cu(feat).requires_grad_(True)
    pooled = crop_pool(f, cu(rois), pooling_size=7, max_pool=True)
    assert tuple(pooled.shape) == (20, 8, 7, 7)
    grid_xy = orc.affine_grid(rois.numpy(), 20, 31, 14, True)
    grid_yx = np.ascontiguousarray(np.stack([grid_xy[..., 1], grid_xy[..., 0]], 3))
    ref14 = orc.roi_crop(feat.numpy(), grid_yx)
    ref = ref14.reshape(20, 8, 7, 2, 7, 2).max(axis=(3, 5))
    close(pooled.detach().cpu().numpy(), ref, what="crop_pool")
    pooled.sum().backward()
    assert f.grad is not None and torch.isfinite(f.grad).all()
    out = _RoICrop()(cu(feat), cu(grid_yx))
    close(out.cpu().numpy(), ref14)


# ------------------------------------------------------------------------------------------
# training-target layers (f3)
# ------------------------------------------------------------------------------------------
def test_proposal_target_layer_vs_reference(orc):
    import os
    from rlobjectdetection_b200.model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    from rlobjectdetection_b200.model.utils.config import cfg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_targets.npz"))
    old = cfg.TRAIN.BATCH_SIZE
    cfg.TRAIN.BATCH_SIZE = int(g["pt_cfg"][0])
    try:
        layer = _ProposalTargetLayer(21)
        ro, lab, tg, iw, ow = layer(cu(g["pt_rois"]), cu(g["pt_gt"]), None, fg_keys=cu(g["pt_fg_keys"]), bg_u=cu(g["pt_bg_u"]))
    finally:
        cfg.TRAIN.BATCH_SIZE = old
    assert int(layer.last_status.sum().item()) == 0
    assert np.array_equal(ro.cpu().numpy(), g["pt_out_rois"])          # same rois in the same order
    assert np.array_equal(lab.cpu().numpy(), g["pt_out_labels"])
    np.testing.assert_allclose(tg.cpu().numpy(), g["pt_out_targets"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(iw.cpu().numpy(), g["pt_out_inside"]) and np.array_equal(ow.cpu().numpy(), g["pt_out_outside"])


def test_proposal_target_layer_train_size_and_corner_cases(orc):
    # TRAIN size: 2000 rois + 20 gt, 128 per image; image 1 has no gt at all (only bg), image 2 only fg-able rois
    B, N, G, R = 3, 2000, 20, 128
    g = torch.Generator().manual_seed(13)
    gt = torch.zeros(B, G, 5)
    gt[0, :7, :4] = syn.random_boxes(g, 7, 600, 1000, 40.0, 300.0)
    gt[0, :7, 4] = torch.arange(1, 8).float()
    gt[2, :2, :4] = torch.tensor([[100.0, 100.0, 300.0, 300.0], [400.0, 50.0, 700.0, 350.0]])
    gt[2, :2, 4] = torch.tensor([3.0, 9.0])
    rois = torch.zeros(B, N, 5)
    for b in range(B):
        rois[b, :, 0] = b
        rois[b, :, 1:] = syn.random_boxes(g, N, 600, 1000, 16.0, 400.0)
    rois[2, :, 1:] = gt[2, torch.randint(0, 2, (N,), generator=g), :4] + torch.randn(N, 4, generator=g) * 3.0
    fg_keys, bg_u = torch.rand(B, N + G, generator=g), torch.rand(B, R, generator=g)
    out = be.proposal_target(cu(rois), cu(gt), cu(fg_keys), cu(bg_u), R, 32, 0.5, 0.5, 0.1, means=(0, 0, 0, 0),
                             stds=(0.1, 0.1, 0.2, 0.2))
    ref = orc.proposal_target(rois.numpy(), gt.numpy(), fg_keys.numpy(), bg_u.numpy(), R, 32)
    names = ["rois", "labels", "targets", "inside", "outside", "status"]
    for o, r, nm in zip(out, ref, names):
        if nm == "targets":
            np.testing.assert_allclose(o.cpu().numpy(), r, rtol=1e-5, atol=1e-6, err_msg=nm)
        else:
            assert np.array_equal(o.cpu().numpy(), r), nm


def test_anchor_target_layer_vs_reference(orc):
    import os
    from rlobjectdetection_b200.model.rpn.anchor_target_layer import _AnchorTargetLayer
    from rlobjectdetection_b200.model.utils.config import cfg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_targets.npz"))
    H, W, stride, bs = (int(v) for v in g["at_cfg"])
    old = (cfg.TRAIN.RPN_BATCHSIZE, cfg.TRAIN.RPN_FG_FRACTION)
    try:
        cfg.TRAIN.RPN_BATCHSIZE = bs
        layer = _AnchorTargetLayer(stride, [8, 16, 32], [0.5, 1, 2])
        score = torch.zeros(2, 18, H, W, device=DEV)
        for tag, frac in (("at", 0.5), ("at2", 0.1)):
            cfg.TRAIN.RPN_FG_FRACTION = frac
            L, T, IW, OW = layer((score, cu(g["at_gt"]), cu(g["at_im_info"]), None), keys=cu(g["at_keys"]))
            assert np.array_equal(L.cpu().numpy(), g[f"{tag}_labels"]), tag
            np.testing.assert_allclose(T.cpu().numpy(), g[f"{tag}_targets"], rtol=1e-5, atol=1e-6)
            assert np.array_equal(IW.cpu().numpy(), g[f"{tag}_inside"])
            np.testing.assert_allclose(OW.cpu().numpy(), g[f"{tag}_outside"], rtol=1e-6)
    finally:
        cfg.TRAIN.RPN_BATCHSIZE, cfg.TRAIN.RPN_FG_FRACTION = old


def test_anchor_target_layer_c1_size(orc):
    # VGG-16 600x1000: 37 x 62 map, 9 anchors, 20 gt, RPN batch 256
    B, H, W, G, A = 2, 37, 62, 20, 9
    g = torch.Generator().manual_seed(21)
    gt = torch.zeros(B, G, 5)
    for b, ng in enumerate((20, 4)):
        gt[b, :ng, :4] = syn.random_boxes(g, ng, 600, 1000, 48.0, 400.0)
        gt[b, :ng, 4] = torch.randint(1, 21, (ng,), generator=g).float()
    im_info = torch.tensor([[600.0, 1000.0, 1.6]] * B)
    keys = torch.rand(B, H * W * A, generator=g)
    anchors = orc.generate_anchors(16, (0.5, 1, 2), (8, 16, 32)).astype(np.float32)
    out = be.anchor_target(cu(gt), cu(im_info), cu(anchors), cu(keys), A, H, W, 16, 0.7, 0.3, False, 0.5, 256, 1.0, -1.0)
    ref = orc.anchor_target(gt.numpy(), im_info.numpy(), anchors, keys.numpy(), H, W, 16)
    assert np.array_equal(out[0].cpu().numpy(), ref[0])
    np.testing.assert_allclose(out[1].cpu().numpy(), ref[1], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out[2].cpu().numpy(), ref[2])
    np.testing.assert_allclose(out[3].cpu().numpy(), ref[3], rtol=1e-6)
    lab = out[0].cpu().numpy()
    assert ((lab == 1).sum(axis=(1, 2, 3)) <= 128).all() and ((lab >= 0).sum(axis=(1, 2, 3)) <= 256).all()
