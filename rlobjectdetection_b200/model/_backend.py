"""ctypes binding of csrc/librlod_sm100a.so (C ABI: include/rlod.h).

PyTorch is used only for device memory (tensors, the caching allocator for workspaces) and
streams; every computation is a hand-written sm_100a kernel behind the C ABI.  There is NO
fallback: if the shared library is missing, or a tensor is not on a CUDA device, the call
raises.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
LIB_PATH = os.path.join(CSRC, "librlod_sm100a.so")

POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2
KERNELS = ["align_fwd", "align_bwd", "align_fwd_generic", "align_bwd_generic", "roi_plan", "nms_mask",
           "nms_scan", "nms_small", "proposal_sort", "pool_fwd", "pool_bwd", "boxes", "reward", "move",
           "nms_lazy", "detect", "crop", "targets"]
IOU_COCO, IOU_RCNN = 0, 1
WTRANS_IDENTITY, WTRANS_EXP_ABS, WTRANS_RAW = 0, 1, 2
SORT_MAX = 16384  # rlod_proposal_forward: min(pre_nms_topN, H*W*A) limit

_c_void_p, _c_int, _c_float, _c_size_t, _c_ll = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                 ctypes.c_size_t, ctypes.c_longlong)
_P, _I, _F, _Z, _L = _c_void_p, _c_int, _c_float, _c_size_t, _c_ll

# name -> (restype, argtypes); mirrors include/rlod.h one to one
SIGNATURES = {
    "rlod_version": (_I, []),
    "rlod_error_string": (ctypes.c_char_p, [_I]),
    "rlod_launch_count": (_L, []),
    "rlod_profile_enable": (_I, [_I]),
    "rlod_profile_only": (_I, [_I]),
    "rlod_profile_collect": (_I, [_I, _P, _P]),
    "rlod_nms_workspace_bytes": (_Z, [_I, _I]),
    "rlod_nms": (_I, [_P, _I, _I, _F, _I, _P, _P, _P, _Z, _P]),
    "rlod_debug_nms_force_large": (_I, [_I]),
    "rlod_nms_batched": (_I, [_P, _I, _P, _I, _I, _F, _I, _P, _P, _P, _Z, _P]),
    "rlod_roi_align_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "rlod_roi_align_forward_route": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rlod_roi_align_forward": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P, _Z, _P]),
    "rlod_roi_align_plan": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _Z, _P]),
    "rlod_roi_align_forward_planned": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "rlod_roi_align_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P,
                                     _Z, _P]),
    "rlod_roi_pool_workspace_bytes": (_Z, [_I, _I]),
    "rlod_roi_pool_forward": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _Z, _P]),
    "rlod_roi_pool_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P]),
    "rlod_proposal_workspace_bytes": (_Z, [_I, _I, _I, _I, _I, _I]),
    "rlod_proposal_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P,
                                   _P, _Z, _P]),
    "rlod_bbox_transform_inv": (_I, [_P, _P, _L, _I, _P, _P]),
    "rlod_clip_boxes": (_I, [_P, _P, _I, _L, _I, _P]),
    "rlod_bbox_overlaps": (_I, [_P, _P, _I, _I, _P, _P]),
    "rlod_bbox_overlaps_batch": (_I, [_P, _L, _I, _P, _I, _I, _I, _I, _P, _P]),
    "rlod_action_reward": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P, _P, _P,
                                _P]),
    "rlod_move_from_act": (_I, [_P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "rlod_reward_refine": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P]),
    "rlod_proposal_target": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rlod_anchor_target_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "rlod_anchor_target": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _I, _F, _I, _F, _F, _P, _P, _P, _P, _P,
                                _Z, _P]),
    "rlod_affine_grid": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "rlod_roi_crop_forward": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rlod_roi_crop_backward": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rlod_rl_labels": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _P, _P]),
    "rlod_detect_postprocess": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _F, _I, _P, _P, _P]),
}

_LIB = None


def build(verbose=False):
    """Compile csrc/*.cu into csrc/librlod_sm100a.so with nvcc for sm_100a (in-tree)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building librlod_sm100a.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C rlobjectdetection_b200/csrc`). There is no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = l
    return _LIB


def profile_collect():
    """{kernel name: (total_ms, launches)} of everything recorded since rlod_profile_enable(1)."""
    out = {}
    for i, name in enumerate(KERNELS):
        ms, n = ctypes.c_double(0.0), ctypes.c_int(0)
        check(lib().rlod_profile_collect(i, ctypes.byref(ms), ctypes.byref(n)), "rlod_profile_collect")
        if n.value:
            out[name] = (ms.value, n.value)
    return out


def check(rc, what):
    if rc != 0:
        msg = lib().rlod_error_string(int(rc)).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            # the reference raises NotImplementedError for CPU features as well
            # (lib/model/roi_align/functions/roi_align.py:28-29)
            raise NotImplementedError(f"{name}: CUDA tensors required, there is no CPU path")


def f32c(t):
    """fp32 + contiguous view of t (copy only when needed)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def workspace(nbytes, device):
    """Caller-owned scratch from torch's caching allocator (stream-ordered with the launch)."""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------
# thin functional layer: one python function per C entry point
# ------------------------------------------------------------------------------------------
def nms_padded(dets, thresh, max_keep=0):
    """rlod_nms.  Returns (keep (n,) int32 padded with -1, num (1,) int32), no host sync."""
    require_cuda("nms", dets)
    dets = f32c(dets)
    n, stride = dets.shape
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=dets.device)
    num = torch.empty(1, dtype=torch.int32, device=dets.device)
    l = lib()
    with torch.cuda.device(dets.device):
        ws = workspace(l.rlod_nms_workspace_bytes(1, n), dets.device)
        check(l.rlod_nms(ptr(dets), n, stride, float(thresh), int(max_keep), ptr(keep), ptr(num),
                         ptr(ws), ws.numel(), stream_of(dets)), "rlod_nms")
    return keep[:n], num


def nms_batched(dets, seg_offsets, thresh, max_seg=None, max_keep=0):
    """rlod_nms_batched.  seg_offsets (nseg+1,) int32 on the device; max_seg: upper bound of
    the segment length (computed with one host read when not given).
    Returns (keep (n,) int32 segment-local indices, -1 padded; num (nseg,) int32)."""
    require_cuda("nms_batched", dets, seg_offsets)
    dets = f32c(dets)
    seg_offsets = seg_offsets.to(torch.int32).contiguous()
    n, stride = dets.shape
    nseg = seg_offsets.numel() - 1
    if max_seg is None:
        max_seg = int((seg_offsets[1:] - seg_offsets[:-1]).max().item()) if nseg > 0 else 0
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=dets.device)
    num = torch.empty(max(nseg, 1), dtype=torch.int32, device=dets.device)
    l = lib()
    with torch.cuda.device(dets.device):
        ws = workspace(l.rlod_nms_workspace_bytes(nseg, max_seg), dets.device)
        check(l.rlod_nms_batched(ptr(dets), stride, ptr(seg_offsets), nseg, int(max_seg),
                                 float(thresh), int(max_keep), ptr(keep), ptr(num), ptr(ws),
                                 ws.numel(), stream_of(dets)), "rlod_nms_batched")
    return keep[:n], num[:nseg]


def _check_rois(features, rois):
    if features.dim() != 4:
        raise ValueError("features must be (B, C, H, W)")
    if rois.dim() != 2 or rois.size(1) != 5:
        # the reference's C glue returns 0 here and the caller ignores it
        # (roi_align_cuda.c:20-24); we raise instead of producing garbage
        raise ValueError("rois must be (R, 5) = [batch_idx, x1, y1, x2, y2]")
    if rois.device != features.device:
        raise ValueError("features and rois must be on the same device")


def feature_layout(features, what):
    """(tensor to pass, channels_last flag) of a (B,C,H,W) feature map.  NCHW-dense and
    channels-last (NHWC-dense) maps are read in place; anything else is refused rather than
    silently copied (a C4 map is 369 MB)."""
    if features.dtype != torch.float32:
        raise TypeError(f"{what}: features must be float32")
    if features.is_contiguous():
        return features, 0
    if features.is_contiguous(memory_format=torch.channels_last):
        return features, 1
    raise ValueError(f"{what}: features must be dense NCHW or channels-last; call .contiguous() explicitly")


def roi_align_forward(features, rois, ah, aw, scale, pool_mode):
    require_cuda("roi_align", features, rois)
    _check_rois(features, rois)
    features, nhwc = feature_layout(features, "roi_align")
    rois = f32c(rois)
    B, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty(R, C, ah, aw, dtype=torch.float32, device=features.device)
    l = lib()
    with torch.cuda.device(features.device):
        ws = workspace(l.rlod_roi_align_workspace_bytes(B, R, ah, aw, pool_mode), features.device)
        check(l.rlod_roi_align_forward(ptr(features), ptr(rois), B, C, H, W, R, ah, aw,
                                       float(scale), pool_mode, nhwc, ptr(out), ptr(ws), ws.numel(),
                                       stream_of(features)), "rlod_roi_align_forward")
    return out


class RoiAlignPlan:
    """The planned rois of one RoIAlign forward call: the workspace rlod_roi_align_plan filled, the geometry it
    was filled for and the stream it was filled on (roi_align_plan -> roi_align_forward_planned)."""

    def __init__(self, ws, geometry, n_rois, stream):
        self.ws, self.geometry, self.n_rois, self.stream = ws, geometry, n_rois, stream


def roi_align_plan(rois, feature_size, ah, aw, scale, pool_mode):
    """rlod_roi_align_plan on the current stream.  feature_size = (B, C, H, W) of the map that will be pooled.
    The returned plan owns its workspace (a fresh allocation: the shared per-device workspace may be rewritten
    before the planned call runs)."""
    require_cuda("roi_align_plan", rois)
    rois = f32c(rois)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be (R, 5)")
    B, C, H, W = (int(v) for v in feature_size)
    R = rois.size(0)
    l = lib()
    with torch.cuda.device(rois.device):
        ws = torch.empty(max(int(l.rlod_roi_align_workspace_bytes(B, R, ah, aw, pool_mode)), 16), dtype=torch.uint8,
                         device=rois.device)
        check(l.rlod_roi_align_plan(ptr(rois), B, C, H, W, R, ah, aw, float(scale), pool_mode, ptr(ws), ws.numel(),
                                    stream_of(rois)), "rlod_roi_align_plan")
    return RoiAlignPlan(ws, (B, C, H, W, ah, aw, pool_mode), R, torch.cuda.current_stream(rois.device))


def roi_align_forward_planned(features, plan):
    """rlod_roi_align_forward_planned on the current stream; the caller orders it after the plan (same stream, or
    an event recorded behind roi_align_plan)."""
    require_cuda("roi_align", features)
    features, nhwc = feature_layout(features, "roi_align")
    B, C, H, W, ah, aw, pool_mode = plan.geometry
    if tuple(features.shape) != (B, C, H, W):
        raise ValueError(f"the plan was made for a {(B, C, H, W)} map, not {tuple(features.shape)}")
    out = torch.empty(plan.n_rois, C, ah, aw, dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib().rlod_roi_align_forward_planned(ptr(features), B, C, H, W, plan.n_rois, ah, aw, pool_mode, nhwc,
                                                   ptr(out), ptr(plan.ws), plan.ws.numel(), stream_of(features)),
              "rlod_roi_align_forward_planned")
    return out


def roi_align_backward(grad_out, rois, features, feature_size, ah, aw, scale, pool_mode,
                       grad_in=None):
    """grad wrt features.  grad_in given: accumulate into it; else a fresh tensor is written
    (no memset: the kernel overwrites every element)."""
    require_cuda("roi_align backward", grad_out, rois)
    grad_out, rois = f32c(grad_out), f32c(rois)
    B, C, H, W = feature_size
    R = rois.size(0)
    accumulate = grad_in is not None
    if grad_in is None:
        grad_in = torch.empty(B, C, H, W, dtype=torch.float32, device=grad_out.device)
    # RoIAlignMax recomputes its argmax from the features with the generic (NCHW) kernel: the one place a
    # channels-last map is copied
    feat = f32c(features) if (features is not None and pool_mode == POOL_MAX) else None
    l = lib()
    with torch.cuda.device(grad_out.device):
        ws = workspace(l.rlod_roi_align_workspace_bytes(B, R, ah, aw, pool_mode), grad_out.device)
        check(l.rlod_roi_align_backward(ptr(grad_out), ptr(rois), ptr(feat), B, C, H, W, R, ah, aw,
                                        float(scale), pool_mode, int(accumulate), ptr(grad_in),
                                        ptr(ws), ws.numel(), stream_of(grad_out)),
              "rlod_roi_align_backward")
    return grad_in


def roi_pool_forward(features, rois, ph, pw, scale, want_argmax=True):
    """-> (out, argmax).  want_argmax=False (inference: nothing will be back-propagated) returns (out, None) and
    lets the kernel track the maxima only."""
    require_cuda("roi_pool", features, rois)
    _check_rois(features, rois)
    features, nhwc = feature_layout(features, "roi_pool")
    rois = f32c(rois)
    B, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty(R, C, ph, pw, dtype=torch.float32, device=features.device)
    argmax = torch.empty(R, C, ph, pw, dtype=torch.int32, device=features.device) if want_argmax else None
    l = lib()
    with torch.cuda.device(features.device):
        ws = workspace(l.rlod_roi_pool_workspace_bytes(B, R), features.device)
        check(l.rlod_roi_pool_forward(ptr(features), ptr(rois), B, C, H, W, R, ph, pw,
                                      float(scale), nhwc, ptr(out), ptr(argmax), ptr(ws), ws.numel(),
                                      stream_of(features)), "rlod_roi_pool_forward")
    return out, argmax


def roi_pool_backward(grad_out, argmax, rois, feature_size, ph, pw, scale, grad_in=None):
    require_cuda("roi_pool backward", grad_out, argmax, rois)
    grad_out, rois = f32c(grad_out), f32c(rois)
    B, C, H, W = feature_size
    R = grad_out.size(0)
    accumulate = grad_in is not None
    if grad_in is None:
        grad_in = torch.empty(B, C, H, W, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(lib().rlod_roi_pool_backward(ptr(grad_out), ptr(argmax), ptr(rois), B, C, H, W, R, ph,
                                           pw, float(scale), int(accumulate), ptr(grad_in),
                                           stream_of(grad_out)),
              "rlod_roi_pool_backward")
    return grad_in


def proposal_forward(scores, deltas, im_info, anchors, feat_stride, pre_nms_topN, post_nms_topN,
                     nms_thresh, return_taps=False):
    """rlod_proposal_forward.  scores (B,2A,H,W), deltas (B,4A,H,W), im_info (B,3),
    anchors (A,4).  Returns rois (B, post, 5) [and order, props, nkeep taps]."""
    require_cuda("_ProposalLayer", scores, deltas, im_info, anchors)
    scores, deltas, im_info, anchors = f32c(scores), f32c(deltas), f32c(im_info), f32c(anchors)
    B, A2, H, W = scores.shape
    A = A2 // 2
    if anchors.shape != (A, 4) or deltas.shape != (B, 4 * A, H, W) or im_info.shape != (B, 3):
        raise ValueError("proposal layer: inconsistent input shapes")
    KA = A * H * W
    pre = pre_nms_topN if 0 < pre_nms_topN < KA else KA
    dev = scores.device
    rois = torch.empty(B, post_nms_topN, 5, dtype=torch.float32, device=dev)
    order = props = nkeep = None
    if return_taps:
        order = torch.empty(B, pre, dtype=torch.int32, device=dev)
        props = torch.empty(B, pre, 4, dtype=torch.float32, device=dev)
        nkeep = torch.empty(B, dtype=torch.int32, device=dev)
    l = lib()
    with torch.cuda.device(dev):
        ws = workspace(l.rlod_proposal_workspace_bytes(B, A, H, W, int(pre_nms_topN),
                                                       int(post_nms_topN)), dev)
        check(l.rlod_proposal_forward(ptr(scores), ptr(deltas), ptr(im_info), ptr(anchors), B, A,
                                      H, W, int(feat_stride), int(pre_nms_topN),
                                      int(post_nms_topN), float(nms_thresh), ptr(rois), ptr(order),
                                      ptr(props), ptr(nkeep), ptr(ws), ws.numel(),
                                      stream_of(scores)), "rlod_proposal_forward")
    if return_taps:
        return rois, order, props, nkeep
    return rois


def _boxes_pair(boxes, gt, mode):
    """fp64 rows stay fp64 in COCO mode (the reference's json boxes are float64); anything
    else is computed from fp32 rows."""
    f64 = mode == IOU_COCO and (boxes.dtype == torch.float64 or gt.dtype == torch.float64)
    if f64:
        return boxes.double().contiguous(), gt.double().contiguous(), 1
    return f32c(boxes), f32c(gt), 0


def action_reward(boxes, gt, act, crowd=None, ngt=None, mode=IOU_COCO, iou_thres=0.0,
                  pos_wratio=1.0, neg_wratio=1.0, want_labels=True, wtrans=WTRANS_EXP_ABS):
    """rlod_action_reward.  boxes (B,N,4), gt (B,G,4), act (A,4) -> reward[, label, weight]
    each (B,N,A).  float64 boxes / gt are kept in fp64 (COCO mode)."""
    require_cuda("action_reward", boxes, gt, act, crowd, ngt)
    boxes, gt, f64 = _boxes_pair(boxes, gt, mode)
    act = f32c(act)
    B, N, _ = boxes.shape
    G = gt.size(1)
    A = act.size(0)
    dev = boxes.device
    if crowd is not None:
        crowd = crowd.to(torch.uint8).contiguous()
    if ngt is not None:
        ngt = ngt.to(torch.int32).contiguous()
    reward = torch.empty(B, N, A, dtype=torch.float32, device=dev)
    label = torch.empty_like(reward) if want_labels else None
    weight = torch.empty_like(reward) if want_labels else None
    with torch.cuda.device(dev):
        check(lib().rlod_action_reward(ptr(boxes), ptr(gt), f64, ptr(crowd), ptr(ngt), ptr(act), B, N, A,
                                       G, int(mode), int(wtrans), float(iou_thres), float(pos_wratio),
                                       float(neg_wratio), ptr(reward), ptr(label), ptr(weight),
                                       stream_of(boxes)), "rlod_action_reward")
    if want_labels:
        return reward, label, weight
    return reward


def move_from_act(boxes, preds, targets, act, maxk, corners=False):
    """rlod_move_from_act.  boxes (B,N,4) [x,y,w,h] -- or, with corners=True, (B,N,4)
    [x1,y1,x2,y2] / (B,N,5) rois [b,x1,y1,x2,y2] -- is updated IN PLACE; returns the device
    int32 count of moved boxes (no host sync)."""
    require_cuda("move_from_act", boxes, preds, targets, act)
    if not (boxes.is_contiguous() and boxes.dtype == torch.float32):
        raise ValueError("move_from_act: boxes must be a contiguous fp32 tensor (updated in place)")
    preds, targets, act = f32c(preds), f32c(targets), f32c(act)
    B, N, stride = boxes.shape
    if stride not in (4, 5) or (stride == 5 and not corners):
        raise ValueError("move_from_act: boxes must be (B,N,4), or (B,N,5) rois with corners=True")
    A = act.size(0)
    correct = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    base = ctypes.c_void_p(boxes.data_ptr() + 4 * (stride - 4))
    with torch.cuda.device(boxes.device):
        check(lib().rlod_move_from_act(base, stride, int(bool(corners)), ptr(preds), ptr(targets),
                                       ptr(act), B, N, A, int(maxk), ptr(correct),
                                       stream_of(boxes)), "rlod_move_from_act")
    return correct


class FusedOutputs(dict):
    """{name: tensor} whose tensors are views into one allocation, `flat` (reward_refine)."""
    flat = None


def reward_refine(rois, gt, act, ngt=None, iou_thres=0.0, pos_wratio=1.0, neg_wratio=1.0, first_image=0,
                  wtrans=WTRANS_EXP_ABS, want=("reward", "label", "weight", "refined", "packed", "moved")):
    """rlod_reward_refine: rewards of every (box, action), the best positive action applied to every
    box, and the packed rows of the gather, in one launch.  rois (B,N,5), gt (B,G,4) x1y1x2y2.
    Returns a dict with the requested tensors (B,N,A) / (B,N,5) / (B,N,5+A) / (1,) int32."""
    require_cuda("reward_refine", rois, gt, act, ngt)
    rois, gt, act = f32c(rois), f32c(gt), f32c(act)
    B, N, five = rois.shape
    if five != 5:
        raise ValueError("reward_refine: rois must be (B, N, 5)")
    G, A = gt.size(1), act.size(0)
    dev = rois.device
    if ngt is not None:
        ngt = ngt.to(torch.int32).contiguous()
    # every requested tensor is a view into ONE allocation (the result's .flat): a caller on another stream takes its
    # copy of the whole result with a single device copy (hotpath.DetectRefineStep) instead of one per tensor
    shapes = [(name, shape) for name, shape in (("reward", (B, N, A)), ("label", (B, N, A)), ("weight", (B, N, A)),
                                                ("refined", (B, N, 5)), ("packed", (B, N, 5 + A))) if name in want]
    sizes = [(name, shape, (shape[0] * shape[1] * shape[2] + 3) // 4 * 4) for name, shape in shapes]  # 16-byte slices
    total = sum(sz for _, _, sz in sizes) + 4
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    out = {name: None for name in ("reward", "label", "weight", "refined", "packed", "moved")}
    pos = 0
    for name, shape, sz in sizes:
        out[name] = flat[pos:pos + shape[0] * shape[1] * shape[2]].view(shape)
        pos += sz
    if "moved" in want:
        out["moved"] = flat[pos:pos + 1].view(torch.int32)
        out["moved"].zero_()
    with torch.cuda.device(dev):
        check(lib().rlod_reward_refine(ptr(rois), ptr(gt), ptr(ngt), ptr(act), B, N, A, G, int(wtrans),
                                       float(iou_thres), float(pos_wratio), float(neg_wratio), int(first_image),
                                       ptr(out["reward"]), ptr(out["label"]), ptr(out["weight"]), ptr(out["refined"]),
                                       ptr(out["packed"]), ptr(out["moved"]), stream_of(rois)), "rlod_reward_refine")
    res = FusedOutputs((k, v) for k, v in out.items() if v is not None)
    res.flat = flat
    return res


def detect_postprocess(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=0.3, max_per_image=100,
                       stds=None, means=None, class_agnostic=False):
    """rlod_detect_postprocess: the per-class threshold / decode / sort / NMS / cap loop of
    test_net.py:244-307 for a whole batch.  Returns dets (B,K,N,5) and counts (B,K) int32 on the
    device; class j of image b = dets[b, j, :counts[b, j]] (rows [x1,y1,x2,y2,score])."""
    require_cuda("detect_postprocess", rois, cls_prob, bbox_pred, im_info)
    rois, cls_prob, im_info = f32c(rois), f32c(cls_prob), f32c(im_info)
    if rois.dim() != 3 or rois.size(2) != 5 or cls_prob.dim() != 3:
        raise ValueError("rois must be (B, N, 5) and cls_prob (B, N, K)")
    B, N, K = cls_prob.shape
    if bbox_pred is not None:
        bbox_pred = f32c(bbox_pred).view(B, N, -1)
        if bbox_pred.size(2) != (4 if class_agnostic else 4 * K):
            raise ValueError("bbox_pred must be (B, N, 4K), or (B, N, 4) when class_agnostic")
    dets = torch.zeros(B, K, N, 5, dtype=torch.float32, device=rois.device)
    counts = torch.zeros(B, K, dtype=torch.int32, device=rois.device)
    arr = (ctypes.c_float * 4)
    s = arr(*[float(v) for v in stds]) if stds is not None else None
    m = arr(*[float(v) for v in means]) if means is not None else None
    if (s is None) != (m is None):
        raise ValueError("stds and means go together")
    with torch.cuda.device(rois.device):
        check(lib().rlod_detect_postprocess(ptr(rois), ptr(cls_prob), ptr(bbox_pred), ptr(im_info), B, N, K,
                                            int(bool(class_agnostic)), s, m, float(thresh), float(nms_thresh),
                                            int(max_per_image), ptr(dets), ptr(counts), stream_of(rois)),
              "rlod_detect_postprocess")
    return dets, counts


def rl_labels(dets, gt, act, det_cat=None, ndet=None, gt_cat=None, crowd=None, ngt=None, iou_thres=0.0,
              pos_wratio=1.0, neg_wratio=1.0, wtrans=WTRANS_EXP_ABS):
    """rlod_rl_labels: labels (B,N,A,3) = (act_id, label, weight) of a collated RL batch.
    dets (B,N,>=4) xywh rows, gt (B,G,4) xywh (float64 rows are kept in fp64); categories int32;
    crowd uint8."""
    require_cuda("rl_labels", dets, gt, act)
    dets, gt, f64 = _boxes_pair(dets, gt, IOU_COCO)
    act = f32c(act)
    B, N, S = dets.shape
    G, A = gt.size(1), act.size(0)
    i32 = lambda t: None if t is None else t.to(device=dets.device, dtype=torch.int32).contiguous()  # noqa: E731
    det_cat, ndet, gt_cat, ngt = i32(det_cat), i32(ndet), i32(gt_cat), i32(ngt)
    crowd = None if crowd is None else crowd.to(device=dets.device, dtype=torch.uint8).contiguous()
    labels = torch.empty(B, N, A, 3, dtype=torch.float32, device=dets.device)
    with torch.cuda.device(dets.device):
        check(lib().rlod_rl_labels(ptr(dets), S, f64, ptr(det_cat), ptr(ndet), ptr(gt), ptr(gt_cat), ptr(crowd), ptr(ngt),
                                   ptr(act), B, N, A, G, int(wtrans), float(iou_thres), float(pos_wratio), float(neg_wratio),
                                   ptr(labels), stream_of(dets)), "rlod_rl_labels")
    return labels


def affine_grid(rois, input_size, grid_size, align_corners=True):
    """rlod_affine_grid: rois (R,5) -> grid_xy (R,g,g,2)."""
    require_cuda("_affine_grid_gen", rois)
    rois = f32c(rois)
    R = rois.size(0)
    grid = torch.empty(R, grid_size, grid_size, 2, dtype=torch.float32, device=rois.device)
    with torch.cuda.device(rois.device):
        check(lib().rlod_affine_grid(ptr(rois), R, int(input_size[0]), int(input_size[1]), int(grid_size),
                                     int(bool(align_corners)), ptr(grid), stream_of(rois)), "rlod_affine_grid")
    return grid


def roi_crop_forward(features, grid_yx):
    require_cuda("_RoICrop", features, grid_yx)
    features, grid_yx = f32c(features), f32c(grid_yx)
    B, C, H, W = features.shape
    R, gh, gw, two = grid_yx.shape
    if two != 2:
        raise ValueError("grid must be (R, gh, gw, 2) = (y, x)")
    out = torch.empty(R, C, gh, gw, dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib().rlod_roi_crop_forward(ptr(features), ptr(grid_yx), B, C, H, W, R, gh, gw, ptr(out),
                                          stream_of(features)), "rlod_roi_crop_forward")
    return out


def roi_crop_backward(grad_out, grid_yx, feature_size, grad_in=None):
    require_cuda("_RoICrop backward", grad_out, grid_yx)
    grad_out, grid_yx = f32c(grad_out), f32c(grid_yx)
    B, C, H, W = feature_size
    R, gh, gw, _ = grid_yx.shape
    accumulate = grad_in is not None
    if grad_in is None:
        grad_in = torch.empty(B, C, H, W, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(lib().rlod_roi_crop_backward(ptr(grad_out), ptr(grid_yx), B, C, H, W, R, gh, gw, int(accumulate),
                                           ptr(grad_in), stream_of(grad_out)), "rlod_roi_crop_backward")
    return grad_in


def _f4(v):
    return None if v is None else (ctypes.c_float * 4)(*[float(x) for x in v])


def proposal_target(rois, gt_boxes, fg_keys, bg_u, rois_per_image, fg_rois_per_image, fg_thresh, bg_hi, bg_lo,
                    means=None, stds=None, inside_weights=(1.0, 1.0, 1.0, 1.0)):
    """rlod_proposal_target -> rois (B,R,5), labels (B,R), targets / inside / outside (B,R,4), status (B)."""
    require_cuda("_ProposalTargetLayer", rois, gt_boxes, fg_keys, bg_u)
    rois, gt_boxes, fg_keys, bg_u = f32c(rois), f32c(gt_boxes), f32c(fg_keys), f32c(bg_u)
    B, N, _ = rois.shape
    G, R = gt_boxes.size(1), int(rois_per_image)
    if tuple(fg_keys.shape) != (B, N + G) or bg_u.size(0) != B or bg_u.size(1) < R:
        raise ValueError("fg_keys must be (B, N+G) and bg_u (B, >= rois_per_image)")
    bg_u = bg_u[:, :R].contiguous()
    dev = rois.device
    ro = torch.empty(B, R, 5, device=dev)
    lab = torch.empty(B, R, device=dev)
    tg, iw, ow = (torch.empty(B, R, 4, device=dev) for _ in range(3))
    status = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().rlod_proposal_target(ptr(rois), ptr(gt_boxes), ptr(fg_keys), ptr(bg_u), B, N, G, R,
                                         int(fg_rois_per_image), float(fg_thresh), float(bg_hi), float(bg_lo),
                                         _f4(means), _f4(stds), _f4(inside_weights), ptr(ro), ptr(lab), ptr(tg), ptr(iw),
                                         ptr(ow), ptr(status), stream_of(rois)), "rlod_proposal_target")
    return ro, lab, tg, iw, ow, status


def anchor_target(gt_boxes, im_info, anchors, keys, A, H, W, feat_stride, positive_overlap, negative_overlap,
                  clobber_positives, fg_fraction, batchsize, inside_weight, positive_weight):
    """rlod_anchor_target -> labels (B,1,A*H,W), bbox_targets / inside / outside (B,4A,H,W)."""
    require_cuda("_AnchorTargetLayer", gt_boxes, im_info, anchors, keys)
    gt_boxes, im_info, anchors, keys = f32c(gt_boxes), f32c(im_info), f32c(anchors), f32c(keys)
    B, G, _ = gt_boxes.shape
    if tuple(keys.shape) != (B, H * W * A):
        raise ValueError("keys must be (B, H*W*A)")
    dev = gt_boxes.device
    labels = torch.empty(B, 1, A * H, W, device=dev)
    tg, iw, ow = (torch.empty(B, 4 * A, H, W, device=dev) for _ in range(3))
    l = lib()
    with torch.cuda.device(dev):
        ws = workspace(l.rlod_anchor_target_workspace_bytes(B, A, H, W), dev)
        check(l.rlod_anchor_target(ptr(gt_boxes), ptr(im_info), ptr(anchors), ptr(keys), B, G, A, H, W, int(feat_stride),
                                   float(positive_overlap), float(negative_overlap), int(bool(clobber_positives)),
                                   float(fg_fraction), int(batchsize), float(inside_weight), float(positive_weight),
                                   ptr(labels), ptr(tg), ptr(iw), ptr(ow), ptr(ws), ws.numel(), stream_of(gt_boxes)),
              "rlod_anchor_target")
    return labels, tg, iw, ow
