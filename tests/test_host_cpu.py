"""CPU-only checks: the C-ABI library loads and exports exactly what include/rlod.h declares,
the host-side mirror of the reference interface behaves (anchors, action table, cfg), the
product refuses to run without a GPU (no fallback), and the multi-GPU plumbing (sharding +
packed all-gather) works at world_size 2 over gloo."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from rlobjectdetection_b200.model import _backend as be
    header = open(os.path.join(ROOT, "include", "rlod.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rlod_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = be.lib()  # loads without a GPU: no CUDA call happens at load time
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rlod.h but not exported"
    assert declared == set(be.SIGNATURES), declared ^ set(be.SIGNATURES)
    assert lib.rlod_version() >= 100
    assert lib.rlod_error_string(-2) == b"workspace too small"


def test_workspace_queries_are_pure_host_arithmetic():
    from rlobjectdetection_b200.model import _backend as be
    lib = be.lib()
    assert lib.rlod_nms_workspace_bytes(1, 300) == 0                      # shared-memory path
    assert lib.rlod_nms_workspace_bytes(1, 12000) == 188 * 188 * 64 * 8   # [row][col_block] mask, rows of 4-word sectors
    assert lib.rlod_nms_workspace_bytes(24, 6000) == 24 * 94 * 96 * 64 * 8
    small = lib.rlod_roi_align_workspace_bytes(4, 1024, 7, 7, be.POOL_AVG)
    # per roi (8x8 sample grid): 128-byte plan + 128-byte forward-kernel record + 768-byte backward-kernel record
    # + order entries
    assert 1024 * (260 + 768) <= small < 1024 * (260 + 768) + 64 * 1024
    # post_nms_topN <= 512: kept-list NMS, no n x n mask in the proposal workspace; beyond: the mask
    assert 12000 * 16 <= lib.rlod_proposal_workspace_bytes(1, 9, 37, 62, 12000, 300) < 12000 * 16 + 188 * 188 * 512
    assert lib.rlod_proposal_workspace_bytes(1, 9, 37, 62, 12000, 2000) >= 12000 * 16 + 188 * 188 * 512


def _route(lib, B, C, H, W, R, pool_mode, channels_last=0):
    import ctypes
    info = (ctypes.c_int * 8)()
    rc = lib.rlod_roi_align_forward_route(B, C, H, W, R, 7, 7, pool_mode, channels_last, ctypes.cast(info, ctypes.c_void_p))
    return rc, list(info)


def test_roi_align_forward_route_is_host_arithmetic():
    """Which kernels the RoIAlign forward takes: plane kernel for the reference's maps, tiles for dense roi sets on
    maps beyond one CTA's shared memory (tiles cover the map, overlap by half, stay within the plane kernel's limits),
    the gather kernel for sparse ones; the last wave of a short launch is split."""
    from rlobjectdetection_b200.model import _backend as be
    lib = be.lib()
    AVG = be.POOL_AVG
    # the reference's configurations: plane kernel, whole map
    for (B, C, H, W, R, split) in ((24, 1024, 50, 75, 14400, 1), (4, 1024, 38, 63, 1024, 2), (3, 1024, 50, 75, 1800, 3),
                                  (6, 1024, 50, 75, 3600, 2), (12, 1024, 50, 75, 7200, 1)):
        rc, info = _route(lib, B, C, H, W, R, AVG)
        assert rc == 0 and info[0] == 1 and info[1:3] == [1, 1] and info[7] == split, (B, C, H, W, R, info)
    assert _route(lib, 2, 6, 20, 30, 64, AVG)[1][0] == 0          # C % 4 != 0: gather kernel
    # large maps: tiles for dense roi sets
    for (B, C, H, W, R) in ((2, 1024, 100, 150, 4000), (2, 256, 200, 304, 2 * 60 * 16), (1, 12, 260, 40, 130),
                            (3, 4, 50, 200, 390), (2, 256, 200, 336, 2 * 128 * 16), (1, 64, 300, 90, 4000)):
        rc, info = _route(lib, B, C, H, W, R, AVG)
        kind, ny, nx, th, tw, sy, sx, _ = info
        assert rc == 0 and kind == 2, (B, C, H, W, R, info)
        assert ny * nx >= 2 and ny * nx <= 128 and R >= 16 * B * ny * nx
        assert (ny - 1) * sy + th >= H and (nx - 1) * sx + tw >= W        # the tiles cover the map
        assert (ny == 1 and th == H) or (sy <= th // 2 and (ny - 2) * sy + th < H)   # half-tile overlap, no spare tile
        assert (nx == 1 and tw == W) or (sx <= tw // 2 and (nx - 2) * sx + tw < W)
        assert (th + 2) * ((tw + 1) | 1) <= 4030                          # two CTAs per SM
        assert W % 4 != 0 or nx == 1 or sx % 4 == 0                       # 16-byte aligned window rows
    # sparse roi sets and maps beyond 128 tiles: gather kernel
    assert _route(lib, 2, 256, 200, 304, 1024, AVG)[1][0] == 0
    assert _route(lib, 1, 64, 600, 600, 100000, AVG)[1][0] == 0
    # channels-last is the plane kernel's only
    assert _route(lib, 2, 64, 25, 38, 64, AVG, channels_last=1)[0] == 0
    assert _route(lib, 2, 1024, 100, 150, 4000, AVG, channels_last=1)[0] == -3   # RLOD_EUNSUPPORTED
    assert _route(lib, 0, 64, 25, 38, 64, AVG)[0] == -1


def test_invalid_arguments_are_rejected_without_a_gpu():
    from rlobjectdetection_b200.model import _backend as be
    lib = be.lib()
    # argument validation happens before any CUDA call
    assert lib.rlod_nms(None, -1, 5, 0.5, 0, None, None, None, 0, None) == -1
    assert lib.rlod_roi_align_forward(None, None, 1, 4, 1, 5, 3, 7, 7, 0.0625, 1, 0, None, None, 0, None) == -1
    assert lib.rlod_roi_align_forward(None, None, 1, 4, 5, 5, 3, 7, 7, 0.0625, 9, 0, None, None, 0, None) == -1
    assert lib.rlod_proposal_forward(None, None, None, None, 1, 9, 4, 4, 16, 100, 10, 0.7, None, None,
                                     None, None, None, 0, None) == -1
    with pytest.raises(RuntimeError, match="invalid argument"):
        be.check(-1, "x")


def test_no_cpu_fallback():
    from rlobjectdetection_b200.model.nms.nms_wrapper import nms
    from rlobjectdetection_b200.model.roi_align.modules.roi_align import RoIAlignAvg
    from rlobjectdetection_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    from rlobjectdetection_b200.model.rpn.bbox_transform import bbox_overlaps
    from rlobjectdetection_b200.model.rpn.proposal_layer import _ProposalLayer
    feat, rois = torch.randn(1, 4, 8, 8), torch.tensor([[0, 0, 0, 50, 50.]])
    with pytest.raises(NotImplementedError):
        RoIAlignAvg(7, 7, 1 / 16.)(feat, rois)
    with pytest.raises(NotImplementedError):
        _RoIPooling(7, 7, 1 / 16.)(feat, rois)
    with pytest.raises(NotImplementedError):
        nms(torch.rand(4, 5), 0.5)
    with pytest.raises(NotImplementedError):
        bbox_overlaps(torch.rand(4, 4), torch.rand(2, 4))
    with pytest.raises(NotImplementedError):
        _ProposalLayer(16, [8, 16, 32], [0.5, 1, 2])((torch.rand(1, 18, 4, 4), torch.rand(1, 36, 4, 4),
                                                     torch.tensor([[64., 64., 1.]]), "TEST"))
    assert nms(torch.zeros(0, 5), 0.5) == []  # the reference returns [] before touching the device


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rlobjectdetection_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("the CPU oracle and", ""), os.path.join(d, f)


def test_anchor_and_action_tables(golden):
    from rlobjectdetection_b200.model.Reinforcement.action import Action
    from rlobjectdetection_b200.model.rpn.generate_anchors import generate_anchors
    from rlobjectdetection_b200.model.rpn.proposal_layer import _ProposalLayer
    a9 = generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    assert np.array_equal(a9, golden["anchors9"])
    assert np.array_equal(a9[0], [-84, -40, 99, 55]) and np.array_equal(a9[8], [-168, -344, 183, 359])
    a12 = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    assert np.array_equal(a12, golden["anchors12"])
    layer = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    assert layer._num_anchors == 12 and layer._anchors.dtype == torch.float32
    assert len(list(layer.parameters())) == 0  # stateless: checkpoints stay compatible
    act = Action([0.5, 0.25])
    assert act.num_acts == 16 and np.array_equal(act.actDeltas, golden["act16"])
    assert np.array_equal(Action([.5, .25, .125, .0625, .03125, .015625, .008]).actDeltas, golden["act56"])
    assert np.array_equal(Action([0.5], alpha=2.0).actDeltas[:2], [[1, 0, 0, 0], [-1, 0, 0, 0]])


def test_cfg_keys_and_overrides():
    from rlobjectdetection_b200.model.utils.config import cfg, cfg_from_list
    assert cfg["TRAIN"].RPN_PRE_NMS_TOP_N == 12000 and cfg.TRAIN.RPN_POST_NMS_TOP_N == 2000
    assert cfg["TEST"].RPN_PRE_NMS_TOP_N == 6000 and cfg.TEST.RPN_POST_NMS_TOP_N == 300
    assert cfg.TEST.RPN_NMS_THRESH == 0.7 and cfg.TEST.NMS == 0.3 and cfg.POOLING_SIZE == 7
    old = cfg.TEST.RPN_POST_NMS_TOP_N
    try:
        cfg_from_list(["TEST.RPN_POST_NMS_TOP_N", "100", "ANCHOR_SCALES", "[4, 8, 16, 32]"])
        assert cfg.TEST.RPN_POST_NMS_TOP_N == 100 and cfg.ANCHOR_SCALES == [4, 8, 16, 32]
        with pytest.raises(KeyError):
            cfg_from_list(["TEST.NO_SUCH_KEY", "1"])
    finally:
        cfg.TEST.RPN_POST_NMS_TOP_N = old
        cfg.ANCHOR_SCALES = [8, 16, 32]


def test_synthetic_inputs_are_deterministic():
    from rlobjectdetection_b200 import synthetic as syn
    s1, d1, i1 = syn.rpn_outputs(3, 2, 9, 5, 7, 80, 112)
    s2, d2, i2 = syn.rpn_outputs(3, 2, 9, 5, 7, 80, 112)
    assert torch.equal(s1, s2) and torch.equal(d1, d2) and torch.equal(i1, i2)
    fg = s1[:, 9:].reshape(2, -1)
    assert all(len(torch.unique(fg[b])) == fg.shape[1] for b in range(2))  # tie-free scores
    dets, seg = syn.clustered_dets(4, 2, 3, 50)
    d = dets.reshape(6, 50, 5)
    assert (d[:, :-1, 4] > d[:, 1:, 4]).all() and seg.tolist() == [0, 50, 100, 150, 200, 250, 300]


def test_shard_bounds_cover_the_batch():
    from rlobjectdetection_b200.shard import shard_bounds
    for gb in (1, 7, 24, 25):
        for ws in (1, 2, 4, 8):
            spans = [shard_bounds(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_bounds(24, 3, 8) == (9, 12)


def _gloo_worker(rank, world, port, gb, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rlobjectdetection_b200.shard import gather_results, pack_results, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        N, A = 5, 3
        full_rois = torch.rand(gb, N, 5, generator=g)
        full_rois[:, :, 0] = torch.arange(gb).float()[:, None]
        full_reward = torch.rand(gb, N, A, generator=g)
        lo, hi = shard_bounds(gb, rank, world)
        local = full_rois[lo:hi].clone()
        local[:, :, 0] -= lo  # rank-local image indices, as the per-rank hot path emits
        out = gather_results(pack_results(local, full_reward[lo:hi], lo), gb)
        expect = torch.cat([full_rois, full_reward], 2)
        ok = out.shape == expect.shape and torch.equal(out, expect)
        # the pipelined form (staging copy + asynchronous collective, two steps in flight): step s gathers the
        # packed rows scaled by s + 1, the staging buffers are reused from step 2 on
        from rlobjectdetection_b200.shard import PipelinedGather
        mine = pack_results(local, full_reward[lo:hi], lo)
        pg = PipelinedGather(mine, gb)
        for s_ in range(5):
            pg.submit(mine * float(s_ + 1))
            if s_ in (0, 3, 4):
                got = pg.result()
                ok = ok and got.shape == expect.shape and torch.equal(got, expect * float(s_ + 1))
        pg.drain()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gb", [8, 7])
def test_two_rank_gather_equals_single_rank(tmp_path, gb):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + gb) % 2000
    mp.spawn(_gloo_worker, args=(2, port, gb, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]


def test_wtrans_code_recognises_the_reference_forms():
    # Action.wtrans (action.py:7-10) -> kernel selector: Action's default, the drop-in's exp_abs, the reference's own
    # Config.act_wtrans passed as a plain callable (config.py:48-51, recognised by probing), anything else = custom
    from math import exp, fabs
    from rlobjectdetection_b200.model import _backend as be
    from rlobjectdetection_b200.model.Reinforcement.action import Action, exp_abs, wtrans_code
    assert wtrans_code(Action([0.5])) == be.WTRANS_IDENTITY
    assert wtrans_code(Action([0.5], wtrans=exp_abs)) == be.WTRANS_EXP_ABS
    assert wtrans_code(Action([0.5], wtrans=lambda x: exp(fabs(x)))) == be.WTRANS_EXP_ABS
    assert wtrans_code(Action([0.5], wtrans=lambda x: x)) == be.WTRANS_IDENTITY
    assert wtrans_code(Action([0.5], wtrans=lambda x: x * x)) == be.WTRANS_RAW
    assert wtrans_code(Action([0.5], wtrans=lambda x: x.clamp(min=0))) == be.WTRANS_RAW  # tensor-only callable


def test_feature_layouts_without_a_gpu():
    # dense NCHW and channels-last maps are taken in place, everything else is refused (never silently copied)
    from rlobjectdetection_b200.model import _backend as be
    x = torch.randn(2, 8, 6, 5)
    assert be.feature_layout(x, "t") == (x, 0) or be.feature_layout(x, "t")[1] == 0
    cl = x.contiguous(memory_format=torch.channels_last)
    t, flag = be.feature_layout(cl, "t")
    assert flag == 1 and t.data_ptr() == cl.data_ptr()
    with pytest.raises(ValueError):
        be.feature_layout(x[:, ::2], "t")
    with pytest.raises(TypeError):
        be.feature_layout(x.double(), "t")


def test_reference_arm_runs_the_reference_python(orc):
    # bench.py --impl reference: the vendored reference modules (baseline/_ref, present after build()) imported
    # unmodified under the two stubs; its rois equal the oracle port's up to the expf ulp
    from baseline import ref_arm
    if not ref_arm.available():
        pytest.skip("baseline/_ref not vendored (no /root/reference at build time)")
    import bench
    inp = [t[:1].contiguous() for t in bench.make_inputs(3, 1)]
    rois, reward, refined, pooled, pooled2 = ref_arm.step(orc, inp, bench.STRIDE, bench.SCALES, bench.RATIOS, 600, 50, 0.7,
                                                          bench.POOL, bench.ACT_DELTA)
    anchors = orc.generate_anchors(16, bench.RATIOS, bench.SCALES).astype(np.float32)
    ref = orc.proposal_layer(inp[0].numpy(), inp[1].numpy(), inp[2].numpy(), anchors, 16, 600, 50, 0.7)
    np.testing.assert_allclose(rois, ref, rtol=3e-6, atol=2e-4)
    assert reward.shape == (1, 50, 16) and pooled.shape == (50, 1024, 7, 7) and refined.shape == rois.shape
    assert "baseline/_ref" in sys.modules["model.rpn.proposal_layer"].__file__.replace(os.sep, "/")
