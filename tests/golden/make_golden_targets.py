"""Generate tests/golden/reference_targets.npz by EXECUTING the reference's own
_ProposalTargetLayer and _AnchorTargetLayer (imported unmodified from /root/reference/lib) on
seeded inputs, with np.random patched so that the random draws are functions of recorded
tensors: permutation(n) := stable argsort of the recorded keys of the set being permuted,
rand(k) := the first k recorded uniforms of that image.  Everything else -- overlaps, thresholds,
label rules, target arithmetic, output layouts, the leaked loop variable of the weights -- is the
reference's own code.

One shim for torch >= 1.0: Tensor.index(idx) (removed) := self[idx], used at
proposal_target_layer_cascade.py:143.

    python tests/golden/make_golden_targets.py      # authoring container only
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_stubs  # noqa: E402

STATE = {}


def patched_permutation(n):
    f = sys._getframe(1)
    loc, fname, line = f.f_locals, os.path.basename(f.f_code.co_filename), f.f_lineno
    i = int(loc["i"])
    if fname == "anchor_target_layer.py":
        # the fg call sits at :133, the bg call at :143
        inds = loc["fg_inds"] if line < 138 else loc["bg_inds"]
        glob = loc["inds_inside"].numpy()[inds.numpy()]
        keys = STATE["anchor_keys"][i][glob]
    else:
        inds = loc["fg_inds"]
        keys = STATE["fg_keys"][i][inds.numpy()]
    assert len(keys) == n
    return np.argsort(keys, kind="stable")


def patched_rand(k):
    f = sys._getframe(1)
    i = int(f.f_locals["i"])
    return STATE["bg_u"][i][:k].astype(np.float64)


def main():
    install_stubs()
    torch.Tensor.index = lambda self, idx: self[idx]
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.utils.config import cfg
    np.random.permutation, np.random.rand = patched_permutation, patched_rand
    out = {}
    rng = np.random.RandomState(11)

    def boxes(n, w, h, smin, smax):
        cx, cy = rng.uniform(0, w, n), rng.uniform(0, h, n)
        bw, bh = np.exp(rng.uniform(np.log(smin), np.log(smax), n)), np.exp(rng.uniform(np.log(smin), np.log(smax), n))
        b = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        b[:, 0::2] = b[:, 0::2].clip(0, w - 1)
        b[:, 1::2] = b[:, 1::2].clip(0, h - 1)
        return b.astype(np.float32)

    # ---- proposal target layer: 3 images (normal, few fg, padded gt) -------------------------
    B, N, G, imw, imh = 3, 300, 6, 600, 400
    cfg.TRAIN.BATCH_SIZE = 64
    gt = np.zeros((B, G, 5), np.float32)
    rois = np.zeros((B, N, 5), np.float32)
    for b in range(B):
        ng = [6, 1, 3][b]
        gt[b, :ng, :4] = boxes(ng, imw, imh, 40, 250)
        gt[b, :ng, 4] = rng.randint(1, 21, ng)
        r = boxes(N, imw, imh, 16, 300)
        # a third of the rois jitter a gt box so that there are foreground candidates
        for j in range(0, N, 3):
            g = gt[b, rng.randint(0, ng), :4]
            r[j] = (g + rng.normal(0, 6, 4)).astype(np.float32)
        rois[b, :, 0], rois[b, :, 1:] = b, r
    STATE["fg_keys"] = rng.rand(B, N + G).astype(np.float32)
    STATE["bg_u"] = rng.rand(B, 64).astype(np.float32)
    layer = _ProposalTargetLayer(21)
    ro, lab, tg, iw, ow = layer(torch.from_numpy(rois), torch.from_numpy(gt), torch.tensor([6, 1, 3]))
    out.update(pt_rois=rois, pt_gt=gt, pt_fg_keys=STATE["fg_keys"], pt_bg_u=STATE["bg_u"], pt_out_rois=ro.numpy(),
               pt_out_labels=lab.numpy(), pt_out_targets=tg.numpy(), pt_out_inside=iw.numpy(), pt_out_outside=ow.numpy(),
               pt_cfg=np.array([64, int(np.round(cfg.TRAIN.FG_FRACTION * 64))], np.int32))
    print("proposal target: fg per image", [(lab.numpy()[b] > 0).sum() for b in range(B)])

    # ---- anchor target layer: 2 images, 14 x 21 map, 9 anchors --------------------------------
    B, H, W, G = 2, 14, 21, 5
    cfg.TRAIN.RPN_BATCHSIZE = 64
    gt = np.zeros((B, G, 5), np.float32)
    for b in range(B):
        ng = [5, 2][b]
        gt[b, :ng, :4] = boxes(ng, W * 16, H * 16, 48, 220)
        gt[b, :ng, 4] = rng.randint(1, 21, ng)
    im_info = np.array([[H * 16.0, W * 16.0, 1.0]] * B, np.float32)
    STATE["anchor_keys"] = rng.rand(B, H * W * 9).astype(np.float32)
    layer = _AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2])
    score = torch.zeros(B, 18, H, W)
    L, T, IW, OW = layer((score, torch.from_numpy(gt), torch.from_numpy(im_info), torch.tensor([5, 2])))
    out.update(at_gt=gt, at_im_info=im_info, at_keys=STATE["anchor_keys"], at_anchors=layer._anchors.numpy(),
               at_labels=L.numpy(), at_targets=T.numpy(), at_inside=IW.numpy(), at_outside=OW.numpy(),
               at_cfg=np.array([H, W, 16, 64], np.int32))
    print("anchor target: fg", [(L.numpy()[b] == 1).sum() for b in range(B)], "bg", [(L.numpy()[b] == 0).sum() for b in range(B)])
    # same inputs, fg fraction 0.1 -> the fg subsampling branch (:128-135) runs too
    cfg.TRAIN.RPN_FG_FRACTION = 0.1
    L, T, IW, OW = layer((score, torch.from_numpy(gt), torch.from_numpy(im_info), torch.tensor([5, 2])))
    out.update(at2_labels=L.numpy(), at2_targets=T.numpy(), at2_inside=IW.numpy(), at2_outside=OW.numpy())
    print("anchor target (fg fraction 0.1): fg", [(L.numpy()[b] == 1).sum() for b in range(B)], "bg",
          [(L.numpy()[b] == 0).sum() for b in range(B)])
    np.savez_compressed(os.path.join(HERE, "reference_targets.npz"), **out)
    print("wrote reference_targets.npz", os.path.getsize(os.path.join(HERE, "reference_targets.npz")), "bytes")


if __name__ == "__main__":
    main()
