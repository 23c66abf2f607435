"""The reference arm of bench.py: the reference's OWN Python for every stage of the path that exists
on the CPU in the reference, imported UNMODIFIED from baseline/_ref/lib (vendored there, git-ignored,
by __graft_entry__.build() from /root/reference -- never committed), plus the CPU port (oracle/) for
the stages the reference only has as CUDA:

  stage                         | code that runs                                           | kind
  ------------------------------|----------------------------------------------------------|----------
  anchors, decode, clip, sort,  | reference lib/model/rpn/proposal_layer.py:49-161 with     | reference
  pre/post top-k, padding       | bbox_transform.py:77-133, generate_anchors.py (torch CPU) |
  NMS inside the layer          | the layer's `nms()` -> nms_gpu -> `_ext.nms.nms_cuda`     | port
                                | (cffi, not buildable: torch.utils.ffi is gone) stubbed by |
                                | the greedy restatement of nms_cuda_kernel.cu:123-144      |
  RoIAlignAvg x2                | CUDA-only in the reference (functions/roi_align.py:28-29) | port
                                | -> oracle/rlod_oracle.c, OpenMP                           |
  action rewards                | reference lib/model/rpn/bbox_transform.py:136-166         | reference
                                | bbox_overlaps (torch CPU) on the moved boxes              |
  Action table                  | reference lib/model/Reinforcement/action.py:6-22          | reference

Only bench.py (cpu_baseline / --impl reference) imports this module; the product never does."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "lib")
VENDOR = [  # relative to <reference>/lib
    "model/__init__.py", "model/rpn/__init__.py", "model/rpn/proposal_layer.py", "model/rpn/bbox_transform.py",
    "model/rpn/generate_anchors.py", "model/utils/__init__.py", "model/utils/config.py", "model/nms/__init__.py",
    "model/nms/nms_wrapper.py", "model/nms/nms_gpu.py", "model/Reinforcement/action.py",
]


def vendor(reference_root):
    """Copy the reference files of the arm into baseline/_ref/lib (build time, authoring container)."""
    import shutil
    n = 0
    for rel in VENDOR:
        src = os.path.join(reference_root, "lib", rel)
        dst = os.path.join(REF_LIB, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


def available():
    return all(os.path.exists(os.path.join(REF_LIB, rel)) for rel in VENDOR)


_LOADED = {}


def load(orc):
    """Import the vendored reference modules under the two stubs of SURVEY appendix A."""
    if _LOADED:
        return _LOADED

    class EasyDict(dict):  # the pip package `easydict` is not installed here
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
    sys.modules.setdefault("easydict", m)

    def nms_cuda(keep, dets, num_out, thresh):  # signature of the cffi symbol (nms/src/nms_cuda.h:4-5)
        k = orc.nms(dets.numpy(), float(thresh))
        keep[: len(k), 0] = torch.from_numpy(k)
        num_out[0] = len(k)
        return 1

    ext = types.ModuleType("model.nms._ext")
    nmsmod = types.ModuleType("model.nms._ext.nms")
    nmsmod.nms_cuda = nms_cuda
    ext.nms = nmsmod
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["model.nms._ext"] = ext
    sys.modules["model.nms._ext.nms"] = nmsmod
    sys.path.insert(0, REF_LIB)
    try:
        from model.rpn.proposal_layer import _ProposalLayer
        from model.rpn.bbox_transform import bbox_overlaps
        from model.utils.config import cfg
        from model.Reinforcement.action import Action
    finally:
        sys.path.remove(REF_LIB)
    _LOADED.update(ProposalLayer=_ProposalLayer, bbox_overlaps=bbox_overlaps, cfg=cfg, Action=Action)
    return _LOADED


def step(orc, inputs, stride, scales, ratios, pre, post, nms_t, pool, act_delta):
    """One pass of the path over `inputs` (CPU tensors) -> (rois, reward, refined, pooled, pooled2)."""
    ref = load(orc)
    scores, deltas, im_info, feat, gt = inputs
    cfg = ref["cfg"]
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = pre, post, nms_t
    key = (stride, tuple(scales), tuple(ratios))
    if _LOADED.get("layer_key") != key:
        _LOADED["layer"] = ref["ProposalLayer"](stride, list(scales), list(ratios))
        _LOADED["layer_key"] = key
        _LOADED["action"] = ref["Action"](list(act_delta))
    layer, action = _LOADED["layer"], _LOADED["action"]
    with torch.no_grad():
        rois = layer((scores, deltas, im_info, "TEST"))                       # reference code, torch CPU
        rois_np = rois.numpy()
        featn = feat.numpy()
        pooled = orc.roi_align(featn, rois_np.reshape(-1, 5), pool, pool, 1.0 / stride, pool_mode=orc.POOL_AVG)
        # rewards: IoU of every moved box against the image's gt with the reference's bbox_overlaps
        act = torch.from_numpy(action.actDeltas)                              # (A, 4)
        B, N, _ = rois.shape
        b = rois[:, :, 1:5]
        w = b[..., 2] - b[..., 0] + 1
        h = b[..., 3] - b[..., 1] + 1
        whwh = torch.stack([w, h, w, h], -1)                                  # (B, N, 4)
        xywh = torch.stack([b[..., 0], b[..., 1], w, h], -1)
        moved = xywh[:, :, None, :] + act[None, None] * whwh[:, :, None, :]   # (B, N, A, 4) in x, y, w, h
        mbox = torch.stack([moved[..., 0], moved[..., 1], moved[..., 0] + moved[..., 2] - 1,
                            moved[..., 1] + moved[..., 3] - 1], -1)
        reward = torch.empty(B, N, act.size(0))
        for i in range(B):
            orig = ref["bbox_overlaps"](b[i].contiguous(), gt[i]).max(dim=1).values            # (N,)
            new = ref["bbox_overlaps"](mbox[i].reshape(-1, 4).contiguous(), gt[i]).max(dim=1).values
            reward[i] = new.view(N, -1) - orig[:, None]
        rr = reward.numpy()
        label = np.where(rr > action.iou_thres, 1.0, -1.0).astype(np.float32)
        refined, _ = orc.refine_best_action(rois_np, rr, label, action.actDeltas)
        pooled2 = orc.roi_align(featn, refined.reshape(-1, 5), pool, pool, 1.0 / stride, pool_mode=orc.POOL_AVG)
    return rois_np, rr, refined, pooled, pooled2
