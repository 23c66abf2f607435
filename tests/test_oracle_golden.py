"""CPU: the oracle (oracle/) against the golden vectors produced by executing the reference's own
Python and its vendored maskApi.c (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest


def test_anchor_known_answer(orc, golden):
    # the actual output of generate_anchors (0-based; the MATLAB comment at
    # generate_anchors.py:12-37 is the 1-based version), SURVEY.md section 4
    kat = np.array([[-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                    [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                    [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]], dtype=np.float64)
    a9 = orc.generate_anchors(16, (0.5, 1, 2), (8, 16, 32))
    assert np.array_equal(a9, kat)
    assert np.array_equal(a9, golden["anchors9"])
    assert np.array_equal(orc.generate_anchors(16, (0.5, 1, 2), (4, 8, 16, 32)), golden["anchors12"])


def test_decode_clip(orc, golden):
    dec = orc.bbox_transform_inv(golden["dec_boxes"], golden["dec_deltas"])
    ref = golden["dec_out"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(dec), fin)
    # torch's vectorised expf (sleef) and glibc expf may differ by an ulp
    np.testing.assert_allclose(dec[fin], ref[fin], rtol=3e-6, atol=1e-4)
    clp = orc.clip_boxes(golden["dec_out"], golden["clip_im_info"])
    assert np.array_equal(clp, golden["clip_out"])  # clamp is exact, inf -> border


def test_overlaps(orc, golden):
    o = orc.bbox_overlaps(golden["ovl_anchors"], golden["ovl_gt"])
    np.testing.assert_array_equal(o, golden["ovl_out"])
    o3 = orc.bbox_overlaps_batch(golden["ovlb_anchors"], golden["ovlb_gt"])
    np.testing.assert_array_equal(o3, golden["ovlb_out3"])
    assert (o3[0, 5] == -1).all() and (o3[:, :, 4:] [o3[:, :, 4:] != -1] == 0).all()
    o2 = orc.bbox_overlaps_batch(golden["ovlb_anchors"][0], golden["ovlb_gt"])
    np.testing.assert_array_equal(o2, golden["ovlb_out2"])


@pytest.mark.parametrize("tag", ["test", "train", "all"])
def test_proposal_layer(orc, golden, tag):
    stride, pre, post, A = [int(v) for v in golden[f"prop_{tag}_cfg"]]
    rois = orc.proposal_layer(golden[f"prop_{tag}_scores"], golden[f"prop_{tag}_deltas"],
                              golden[f"prop_{tag}_im_info"], golden[f"prop_{tag}_anchors"],
                              stride, pre, post, 0.7)
    ref = golden[f"prop_{tag}_rois"]
    assert rois.shape == ref.shape
    # identical keep sets and order; coordinates up to the expf ulp
    assert np.array_equal(rois[:, :, 0], ref[:, :, 0])
    assert np.array_equal(rois[:, :, 1:].any(axis=2), ref[:, :, 1:].any(axis=2))
    np.testing.assert_allclose(rois, ref, rtol=3e-6, atol=2e-4)


def test_action_table_and_move(orc, golden):
    assert np.array_equal(orc.action_table([0.5, 0.25]), golden["act16"])
    assert np.array_equal(orc.action_table([.5, .25, .125, .0625, .03125, .015625, .008]), golden["act56"])
    for k in (1, 5):
        moved, prec = orc.move_from_act(golden["move_in_boxes"], golden["move_preds"],
                                        golden["move_targets"], golden["act16"], k)
        np.testing.assert_array_equal(moved, golden[f"move_k{k}_boxes"])
        assert prec == float(golden[f"move_k{k}_prec"])


def test_move_from_act_tied_preds_vs_reference(orc, golden):
    # equal predictions: np.flip(np.argsort) visits the higher flat index first (reference run with the sort
    # pinned stable; "tie_all" = all preds equal, where the reference's unpinned run is the fixture)
    for tag in ("tie_small", "tie_large", "tie_all"):
        for k in (1, 3):
            moved, prec = orc.move_from_act(golden[f"{tag}_boxes"], golden[f"{tag}_preds"], golden[f"{tag}_targets"],
                                            golden["act16"], k)
            np.testing.assert_array_equal(moved, golden[f"{tag}_k{k}_out"])
            assert prec == float(golden[f"{tag}_k{k}_prec"])


def test_reward_weight_transforms_and_f64_rows(orc, golden):
    # weight = wtrans(delta_iou) * ratio: identity (Action's default) vs exp(|x|) (Config.act_wtrans); and the
    # fp64-row loop equals the C port on fp32-representable boxes
    a = (golden["iou_dt"][None], golden["iou_gt"][None], golden["act16"])
    r1, l1, w1 = orc.action_reward(*a, crowd=golden["iou_crowd"][None], pos_wratio=2.0, neg_wratio=0.5,
                                   wtrans=orc.WTRANS_EXP_ABS)
    r0, l0, w0 = orc.action_reward(*a, crowd=golden["iou_crowd"][None], pos_wratio=2.0, neg_wratio=0.5,
                                   wtrans=orc.WTRANS_IDENTITY)
    assert np.array_equal(r0, r1) and np.array_equal(l0, l1)
    np.testing.assert_allclose(w0[0], golden["reward_out"] * np.where(golden["reward_label"] > 0, 2.0, 0.5), rtol=1e-6)
    rf, lf, wf = orc.action_reward_f64(*a, crowd=golden["iou_crowd"][None], pos_wratio=2.0, neg_wratio=0.5)
    assert np.array_equal(rf, r1) and np.array_equal(lf, l1)
    np.testing.assert_allclose(wf, w1, rtol=1e-6)


def test_bbiou_and_reward(orc, golden):
    o = orc.bbiou(golden["iou_dt"], golden["iou_gt"], golden["iou_crowd"])
    np.testing.assert_array_equal(o, golden["iou_out"])  # bit-exact fp64 vs maskApi.c
    assert abs(orc.bbiou([[10, 10, 20, 20]], [[12, 12, 20, 20]], [0])[0, 0] - 324.0 / 476.0) < 1e-15
    assert orc.bbiou([[0, 0, 5, 5]], [[0, 0, 10, 10]], [1])[0, 0] == 1.0  # crowd: union = dt area
    r, l, w = orc.action_reward(golden["iou_dt"][None], golden["iou_gt"][None], golden["act16"],
                                crowd=golden["iou_crowd"][None], mode=orc.MODE_COCO, iou_thres=0.0,
                                pos_wratio=2.0, neg_wratio=0.5)
    np.testing.assert_array_equal(r[0], golden["reward_out"].astype(np.float32))
    np.testing.assert_array_equal(l[0], golden["reward_label"].astype(np.float32))
    np.testing.assert_allclose(w[0], golden["reward_weight"], rtol=1e-6)


def test_ref_maskapi_when_present(orc, golden):
    if orc.ref_maskapi() is None:
        pytest.skip("oracle/_ref/libmaskapi.so not built (needs /root/reference)")
    rng = np.random.default_rng(0)
    dt = np.concatenate([rng.random((200, 2)) * 300, rng.random((200, 2)) * 100], 1)
    gt = np.concatenate([rng.random((50, 2)) * 300, rng.random((50, 2)) * 100], 1)
    cr = (rng.random(50) < 0.2).astype(np.uint8)
    np.testing.assert_array_equal(orc.bbiou(dt, gt, cr), orc.ref_bbiou(dt, gt, cr))


@pytest.mark.parametrize("tag", ["c21", "c9", "agn"])
def test_detect_postprocess(orc, tag):
    """test_net.py:244-307 executed with the reference's own decode / clip / sort
    (tests/golden/make_golden_detect.py) vs the oracle restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_detect.npz"))
    N, K, agn, cap = (int(v) for v in g[f"{tag}_cfg"])
    ab = orc.detect_postprocess(g[f"{tag}_rois"], g[f"{tag}_cls_prob"], g[f"{tag}_bbox_pred"], g[f"{tag}_im_info"],
                                thresh=float(g[f"{tag}_thresh"]), nms_thresh=0.3, max_per_image=cap,
                                stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0), class_agnostic=bool(agn))[0]
    assert [len(a) for a in ab] == list(g[f"{tag}_counts"])
    got = np.concatenate(ab, 0)
    np.testing.assert_allclose(got, g[f"{tag}_dets"], rtol=1e-6, atol=1e-4)
    assert np.array_equal(got[:, 4], g[f"{tag}_dets"][:, 4])  # scores pass through untouched


def test_rl_labels_vs_reference_collate(orc):
    """One collated RL batch from the reference's own _collate_fn + label loop
    (tests/golden/make_golden_rl.py) vs the oracle restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_rl.npz"))
    lab = orc.rl_labels(g["dets"], g["det_cat"], g["ndet"], g["gt"], g["gt_cat"], g["crowd"], g["ngt"], g["act"],
                        iou_thres=0.0, pos_wratio=float(g["wratio"][0]), neg_wratio=float(g["wratio"][1]))
    assert np.array_equal(lab[..., :2], g["padded_labels"][..., :2])
    np.testing.assert_allclose(lab[..., 2], g["padded_labels"][..., 2], rtol=2e-7)


def test_affine_grid_and_roi_crop_vs_torch(orc):
    """POOLING_MODE 'crop': the grid is pinned to torch.nn.functional.affine_grid (the call the
    reference makes, net_utils.py:163) in both of its historical behaviours, and the sampler to
    grid_sample(align_corners=True, zeros padding), which is the arithmetic of
    roi_crop_cuda_kernel.cu:11-23,87-113."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    R, H, W, gs = 9, 20, 31, 14
    x1, y1 = torch.rand(R, generator=g) * 300, torch.rand(R, generator=g) * 200
    rois = torch.stack([torch.zeros(R), x1, y1, x1 + torch.rand(R, generator=g) * 250, y1 + torch.rand(R, generator=g) * 200], 1)
    rois[0] = torch.tensor([0, -40.0, -30.0, 600.0, 400.0])  # partly outside the map
    a, b, c, d = (rois[:, i:i + 1] / 16.0 for i in (1, 2, 3, 4))
    zero = torch.zeros_like(a)
    theta = torch.cat([(c - a) / (W - 1), zero, (a + c - W + 1) / (W - 1), zero, (d - b) / (H - 1),
                       (b + d - H + 1) / (H - 1)], 1).view(-1, 2, 3)          # net_utils.py:155-161
    for ac in (True, False):
        ref = F.affine_grid(theta, torch.Size((R, 1, gs, gs)), align_corners=ac).numpy()
        np.testing.assert_allclose(orc.affine_grid(rois.numpy(), H, W, gs, ac), ref, rtol=0, atol=2e-6)
    grid_xy = torch.from_numpy(orc.affine_grid(rois.numpy(), H, W, gs, True))
    feat = torch.randn(3, 5, H, W, generator=g)
    grid_yx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
    out = orc.roi_crop(feat.numpy(), grid_yx.numpy())
    per = R // 3
    ref = torch.cat([F.grid_sample(feat[r // per:r // per + 1], grid_xy[r:r + 1], mode="bilinear", padding_mode="zeros",
                                   align_corners=True) for r in range(R)], 0).numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_target_layers_vs_reference(orc):
    """_ProposalTargetLayer / _AnchorTargetLayer executed unmodified with np.random patched to
    recorded keys (tests/golden/make_golden_targets.py) vs the oracle ports."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_targets.npz"))
    R, fgp = (int(v) for v in g["pt_cfg"])
    ro, lab, tg, iw, ow, status = orc.proposal_target(g["pt_rois"], g["pt_gt"], g["pt_fg_keys"], g["pt_bg_u"], R, fgp)
    assert (status == 0).all()
    assert np.array_equal(ro, g["pt_out_rois"]) and np.array_equal(lab, g["pt_out_labels"])
    np.testing.assert_allclose(tg, g["pt_out_targets"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(iw, g["pt_out_inside"]) and np.array_equal(ow, g["pt_out_outside"])
    H, W, stride, bs = (int(v) for v in g["at_cfg"])
    for tag, frac in (("at", 0.5), ("at2", 0.1)):
        L, T, IW, OW = orc.anchor_target(g["at_gt"], g["at_im_info"], g["at_anchors"], g["at_keys"], H, W, stride,
                                         fg_fraction=frac, batchsize=bs)
        assert np.array_equal(L, g[f"{tag}_labels"])
        np.testing.assert_allclose(T, g[f"{tag}_targets"], rtol=1e-6, atol=1e-6)
        assert np.array_equal(IW, g[f"{tag}_inside"])
        np.testing.assert_allclose(OW, g[f"{tag}_outside"], rtol=1e-7)
