"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON (imported unmodified from
/root/reference/lib under two sys.modules stubs) on seeded synthetic inputs, and the reference's
vendored maskApi.c (compiled unchanged into oracle/_ref/libmaskapi.so).

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; tests never read /root/reference.

Stubs: `easydict` (pip package missing here) and `model.nms._ext.nms`
(the cffi extension cannot be built: torch.utils.ffi is gone).  The nms stub is backed by the
oracle's greedy NMS; the NMS arithmetic itself is pinned separately against the reference's
legacy CUDA kernel on the GPU (tests/test_gpu_legacy.py, oracle/_ref/libref_legacy.so).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RLOD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402


def install_stubs():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            dict.__setitem__(self, k, v)
            object.__setattr__(self, k, v)

        __setitem__ = __setattr__

    m = types.ModuleType("easydict")
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m

    def nms_cuda(keep, dets, num_out, thresh):
        k = orc.nms(dets.numpy(), float(thresh))
        keep[: len(k), 0] = torch.from_numpy(k)
        num_out[0] = len(k)
        return 1

    ext = types.ModuleType("model.nms._ext")
    nmsmod = types.ModuleType("model.nms._ext.nms")
    nmsmod.nms_cuda = nms_cuda
    ext.nms = nmsmod
    sys.modules["model.nms._ext"] = ext
    sys.modules["model.nms._ext.nms"] = nmsmod
    sys.path.insert(0, os.path.join(REF, "lib"))


def distinct_scores(gen, shape_fg):
    """fg scores = random permutation of (1..n)/(n+1): all distinct -> tie-free ordering."""
    n = int(np.prod(shape_fg))
    perm = torch.randperm(n, generator=gen).float() + 1.0
    return (perm / (n + 1)).reshape(shape_fg)


def main():
    install_stubs()
    from model.rpn.generate_anchors import generate_anchors
    from model.rpn.bbox_transform import (bbox_transform_inv, clip_boxes, bbox_overlaps,
                                          bbox_overlaps_batch)
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    from model.Reinforcement.action import Action

    out = {}
    # ---- anchors (generate_anchors.py:45-56) -------------------------------------------
    out["anchors9"] = generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    out["anchors12"] = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    assert np.array_equal(out["anchors9"], orc.generate_anchors(16, (0.5, 1, 2), (8, 16, 32)))
    assert np.array_equal(out["anchors12"], orc.generate_anchors(16, (0.5, 1, 2), (4, 8, 16, 32)))

    # ---- decode / clip / overlaps (bbox_transform.py) ----------------------------------
    g = torch.Generator().manual_seed(10)
    B, N, k = 2, 257, 3
    xy = torch.rand(B, N, 2, generator=g) * 500
    wh = torch.rand(B, N, 2, generator=g) * 300 + 1
    boxes = torch.cat([xy, xy + wh], 2)
    deltas = torch.randn(B, N, 4 * k, generator=g) * 0.5
    deltas[0, 0, 2] = 100.0  # exp overflow -> inf, clipped later (no clamp on dw/dh, :90-91)
    im_info = torch.tensor([[600.0, 1000.0, 1.6], [480.0, 640.0, 1.0]])
    dec = bbox_transform_inv(boxes, deltas, B)
    clp = clip_boxes(dec.clone(), im_info, B)
    out.update(dec_boxes=boxes.numpy(), dec_deltas=deltas.numpy(), dec_out=dec.numpy(),
               clip_im_info=im_info.numpy(), clip_out=clp.numpy())
    gt = torch.cat([torch.rand(13, 2, generator=g) * 500, torch.rand(13, 2, generator=g) * 300 + 520], 1)
    out["ovl_anchors"] = boxes[0].numpy()
    out["ovl_gt"] = gt.numpy()
    out["ovl_out"] = bbox_overlaps(boxes[0], gt).numpy()
    gtb = torch.zeros(B, 6, 5)
    gtb[:, :4, :4] = torch.cat([torch.rand(B, 4, 2, generator=g) * 300, torch.rand(B, 4, 2, generator=g) * 300 + 310], 2)
    anc = boxes.clone()
    anc[0, 5] = torch.tensor([7.0, 9.0, 7.0, 9.0])  # degenerate anchor -> -1
    out["ovlb_anchors"] = anc.numpy()
    out["ovlb_gt"] = gtb.numpy()
    out["ovlb_out3"] = bbox_overlaps_batch(anc, gtb).numpy()
    out["ovlb_out2"] = bbox_overlaps_batch(anc[0], gtb).numpy()

    # ---- _ProposalLayer (proposal_layer.py:49-161) -------------------------------------
    for tag, (Bp, A, H, W, key, scales, pre, post) in {
        "test": (2, 9, 14, 21, "TEST", [8, 16, 32], 600, 50),
        "train": (1, 12, 10, 16, "TRAIN", [4, 8, 16, 32], 1200, 200),
        "all": (2, 9, 5, 7, "TEST", [8, 16, 32], 6000, 300),  # pre_nms_topN > K*A
    }.items():
        g = torch.Generator().manual_seed(20 + len(tag))
        cfg[key].RPN_PRE_NMS_TOP_N = pre
        cfg[key].RPN_POST_NMS_TOP_N = post
        cfg[key].RPN_NMS_THRESH = 0.7
        scores = torch.cat([torch.rand(Bp, A, H, W, generator=g), distinct_scores(g, (Bp, A, H, W))], 1)
        deltas = torch.randn(Bp, 4 * A, H, W, generator=g) * 0.2
        im_info = torch.tensor([[H * 16.0, W * 16.0, 1.0]] * Bp)
        layer = _ProposalLayer(16, scales, [0.5, 1, 2])
        rois = layer((scores, deltas, im_info, key))
        out.update({f"prop_{tag}_scores": scores.numpy(), f"prop_{tag}_deltas": deltas.numpy(),
                    f"prop_{tag}_im_info": im_info.numpy(), f"prop_{tag}_rois": rois.numpy(),
                    f"prop_{tag}_cfg": np.array([16, pre, post, A], dtype=np.int32),
                    f"prop_{tag}_anchors": layer._anchors.numpy()})

    # ---- Action table / move_from_act (action.py) --------------------------------------
    act = Action([0.5, 0.25])
    act56 = Action([.5, .25, .125, .0625, .03125, .015625, .008])
    out["act16"] = act.actDeltas
    out["act56"] = act56.actDeltas
    g = torch.Generator().manual_seed(30)
    b, n = 3, 40
    bb = np.concatenate([torch.rand(b, n, 2, generator=g).numpy() * 400,
                         torch.rand(b, n, 2, generator=g).numpy() * 200 + 4], 2).astype(np.float32)
    preds = torch.randperm(b * n * 16, generator=g).float().reshape(b, n, 16).numpy() / 7.0
    targets = np.where(torch.rand(b, n, 16, generator=g).numpy() > 0.5, 1.0, -1.0).astype(np.float32)
    for maxk in (1, 5):
        moved, prec = act.move_from_act(bb.copy(), preds, targets, maxk)
        out[f"move_k{maxk}_boxes"] = moved
        out[f"move_k{maxk}_prec"] = np.array(prec)
    out.update(move_in_boxes=bb, move_preds=preds, move_targets=targets)
    # tied / saturated predictions: the visit order is np.flip(np.argsort(pred)) (action.py:44).  numpy's
    # default sort is stable only for short or presorted runs (and SIMD-dispatched beyond), so the fixture is
    # the reference's own move_from_act with np.argsort pinned to kind='stable' -- among equal preds the
    # HIGHER flat index is visited first.  Where the unpinned run agrees on this machine it is said so.
    import model.Reinforcement.action as ref_action
    real_argsort = np.argsort

    class _StableNp:
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def argsort(a, *args, **kw):
            return real_argsort(a, kind="stable")

    for tag, (tb, tn, levels) in {"tie_small": (2, 1, 3), "tie_large": (2, 40, 2), "tie_all": (2, 40, 1)}.items():
        tbx = np.concatenate([torch.rand(tb, tn, 2, generator=g).numpy() * 400,
                              torch.rand(tb, tn, 2, generator=g).numpy() * 200 + 4], 2).astype(np.float32)
        tpr = (torch.randint(0, levels, (tb, tn, 16), generator=g).float() / max(levels - 1, 1)).numpy()
        ttg = np.where(torch.rand(tb, tn, 16, generator=g).numpy() > 0.4, 1.0, -1.0).astype(np.float32)
        out.update({f"{tag}_boxes": tbx, f"{tag}_preds": tpr, f"{tag}_targets": ttg})
        for maxk in (1, 3):
            ref_action.np = _StableNp()
            try:
                moved, prec = act.move_from_act(tbx.copy(), tpr, ttg, maxk)
            finally:
                ref_action.np = np
            plain, plain_prec = act.move_from_act(tbx.copy(), tpr, ttg, maxk)
            print(f"{tag} maxk={maxk}: unpinned numpy sort agrees with the stable one here:",
                  bool(np.array_equal(plain, moved) and plain_prec == prec))
            if tag == "tie_all":  # all preds equal (an untrained head): the UNPINNED reference run is the fixture
                assert np.array_equal(plain, moved) and plain_prec == prec
            out[f"{tag}_k{maxk}_out"] = moved
            out[f"{tag}_k{maxk}_prec"] = np.array(prec)

    # ---- bbIou + reward loop (maskApi.c:98-109 compiled unchanged; RL_coco_dataset.py:119-137)
    g = torch.Generator().manual_seed(40)
    nb, ng = 12, 5
    dt = np.concatenate([torch.rand(nb, 2, generator=g).numpy() * 300,
                         torch.rand(nb, 2, generator=g).numpy() * 200 + 2], 1).astype(np.float32)
    gtx = np.concatenate([torch.rand(ng, 2, generator=g).numpy() * 300,
                          torch.rand(ng, 2, generator=g).numpy() * 200 + 2], 1).astype(np.float32)
    crowd = np.array([0, 1, 0, 0, 1], dtype=np.uint8)
    out["iou_dt"], out["iou_gt"], out["iou_crowd"] = dt, gtx, crowd
    out["iou_out"] = orc.ref_bbiou(dt, gtx, crowd)
    kat = orc.ref_bbiou([[10, 10, 20, 20]], [[12, 12, 20, 20]], [0])
    assert abs(kat[0, 0] - 324.0 / 476.0) < 1e-15
    rewards = np.zeros((nb, 16), dtype=np.float64)
    labels = np.zeros((nb, 16), dtype=np.float64)
    weights = np.zeros((nb, 16), dtype=np.float64)
    from math import exp, fabs
    for i in range(nb):
        bbox = [float(v) for v in dt[i]]
        w, h = bbox[2], bbox[3]
        origin_ious = orc.ref_bbiou([bbox], gtx, crowd)
        for act_id, act_delta in enumerate(act.actDeltas):
            new_bbox = bbox + act_delta * np.array([w, h, w, h])
            new_ious = orc.ref_bbiou([new_bbox], gtx, crowd)
            delta_iou = new_ious.max() - origin_ious.max()
            rewards[i, act_id] = delta_iou
            labels[i, act_id] = 1 if delta_iou > act.iou_thres else -1
            weights[i, act_id] = exp(fabs(delta_iou)) * (2.0 if delta_iou > act.iou_thres else 0.5)
    out.update(reward_out=rewards, reward_label=labels, reward_weight=weights)

    np.savez_compressed(os.path.join(HERE, "reference_python.npz"), **out)
    print("wrote reference_python.npz with", len(out), "arrays,",
          os.path.getsize(os.path.join(HERE, "reference_python.npz")), "bytes")


if __name__ == "__main__":
    main()
