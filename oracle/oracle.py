"""Python face of the parity oracle (TEST INFRASTRUCTURE ONLY -- never imported by the
product package `rlobjectdetection_b200`; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module).

Heavy loops live in rlod_oracle.c (liboracle.so, built by `make -C oracle`); the small
pieces (anchor table, action table, move_from_act) are restated in numpy here.  Every
function cites the reference file:line it follows (paths relative to the reference
checkout jbr97/RLObjectDetection).

Pinning status: PINNED -- see oracle/README.md.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = ctypes.POINTER(ctypes.c_float)
c_d = ctypes.POINTER(ctypes.c_double)
c_i = ctypes.POINTER(ctypes.c_int)
c_u8 = ctypes.POINTER(ctypes.c_ubyte)


def build(verbose=False):
    """make -C oracle (liboracle.so always; _ref/*.so only when /root/reference exists)."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_nms.restype = ctypes.c_int
        _LIB.orc_nms_iou.restype = ctypes.c_float
        _LIB.orc_max_threads.restype = ctypes.c_int
        _LIB.orc_roi_pool_fwd_cpu_nhwc.restype = ctypes.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    lib().orc_set_threads(int(n))


# ----------------------------------------------------------------------------------------
# anchors: lib/model/rpn/generate_anchors.py:45-105
# ----------------------------------------------------------------------------------------
def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)

    def whctrs(a):  # :58-67
        w = a[2] - a[0] + 1
        h = a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, xc, yc):  # :69-81
        ws = ws[:, None]
        hs = hs[:, None]
        return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))

    base = np.array([1, 1, base_size, base_size], dtype=np.float64) - 1
    w, h, xc, yc = whctrs(base)
    size_ratios = (w * h) / ratios  # :88-90
    ws = np.round(np.sqrt(size_ratios))  # :91 (np.round = half to even)
    hs = np.round(ws * ratios)  # :92
    ratio_anchors = mk(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):  # :53-54 ratio-major, scale-minor
        w, h, xc, yc = whctrs(ratio_anchors[i])
        out.append(mk(w * scales, h * scales, xc, yc))
    return np.vstack(out)


# ----------------------------------------------------------------------------------------
# NMS: lib/model/nms/src/nms_cuda_kernel.cu:31-39,77-81,123-144
# ----------------------------------------------------------------------------------------
def nms(dets, thresh, max_keep=0):
    dets = _f32(dets)
    n, stride = dets.shape if dets.ndim == 2 else (0, 5)
    keep = np.empty(max(n, 1), dtype=np.int32)
    num = lib().orc_nms(_p(dets, c_f), n, stride, ctypes.c_float(thresh), int(max_keep), _p(keep, c_i))
    return keep[:num].copy()


def nms_batched(dets, seg_offsets, thresh, max_keep=0):
    dets = _f32(dets)
    seg = np.ascontiguousarray(seg_offsets, dtype=np.int32)
    nseg = seg.shape[0] - 1
    keep = np.full(max(dets.shape[0], 1), -1, dtype=np.int32)
    num = np.zeros(max(nseg, 1), dtype=np.int32)
    lib().orc_nms_batched(_p(dets, c_f), dets.shape[1], _p(seg, c_i), nseg, ctypes.c_float(thresh),
                          int(max_keep), _p(keep, c_i), _p(num, c_i))
    return keep, num[:nseg]


def nms_iou(a, b):
    a = _f32(a)
    b = _f32(b)
    return float(lib().orc_nms_iou(_p(a, c_f), _p(b, c_f)))


# ----------------------------------------------------------------------------------------
# RoIAlign: lib/model/roi_align/src/roi_align_kernel.cu:15-70,94-143 + modules/roi_align.py
# ----------------------------------------------------------------------------------------
POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2


def roi_align_grid(feat, rois, ah, aw, scale):
    """RoIAlignFunction.forward: (R,C,ah,aw) bilinear samples."""
    feat = _f32(feat)
    rois = _f32(rois)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, C, ah, aw), dtype=np.float32)
    lib().orc_roi_align_fwd(_p(feat, c_f), _p(rois, c_f), B, C, H, W, R, ah, aw, ctypes.c_float(scale), _p(out, c_f))
    return out


def pool2x2(x, is_max):
    x = _f32(x)
    R, C, ah, aw = x.shape
    y = np.empty((R, C, ah - 1, aw - 1), dtype=np.float32)
    lib().orc_pool2x2(_p(x, c_f), ctypes.c_long(R * C), ah, aw, int(is_max), _p(y, c_f))
    return y


def roi_align(feat, rois, ph, pw, scale, pool_mode=POOL_AVG):
    """RoIAlign (pool_mode NONE: ph x pw grid), RoIAlignAvg / RoIAlignMax ((ph+1)x(pw+1) grid
    then 2x2 stride-1 pool) -- modules/roi_align.py:14-16, 26-29, 39-42."""
    if pool_mode == POOL_NONE:
        return roi_align_grid(feat, rois, ph, pw, scale)
    g = roi_align_grid(feat, rois, ph + 1, pw + 1, scale)
    return pool2x2(g, pool_mode == POOL_MAX)


def roi_align_bwd(grad_out, feat, rois, ph, pw, scale, pool_mode=POOL_AVG):
    """d(features) of roi_align(); returns fp64 (B,C,H,W).  `feat` is needed for MAX only."""
    grad_out = _f32(grad_out)
    rois = _f32(rois)
    feat = _f32(feat)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    if pool_mode == POOL_NONE:
        ah, aw, gx = ph, pw, grad_out
    else:
        ah, aw = ph + 1, pw + 1
        x = roi_align_grid(feat, rois, ah, aw, scale) if pool_mode == POOL_MAX else None
        gx = np.empty((R, C, ah, aw), dtype=np.float32)
        lib().orc_pool2x2_bwd(_p(grad_out, c_f), _p(x, c_f), ctypes.c_long(R * C), ah, aw,
                              int(pool_mode == POOL_MAX), _p(gx, c_f))
    gin = np.zeros((B, C, H, W), dtype=np.float64)
    lib().orc_roi_align_bwd(_p(gx, c_f), _p(rois, c_f), B, C, H, W, R, ah, aw, ctypes.c_float(scale), _p(gin, c_d))
    return gin


# ----------------------------------------------------------------------------------------
# RoIPool: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93,128-203; roi_pooling.c:4-104
# ----------------------------------------------------------------------------------------
def roi_pool(feat, rois, ph, pw, scale):
    feat = _f32(feat)
    rois = _f32(rois)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, C, ph, pw), dtype=np.float32)
    arg = np.zeros((R, C, ph, pw), dtype=np.int32)
    lib().orc_roi_pool_fwd(_p(feat, c_f), _p(rois, c_f), B, C, H, W, R, ph, pw, ctypes.c_float(scale),
                           _p(out, c_f), _p(arg, c_i))
    return out, arg


def roi_pool_bwd(grad_out, argmax, feat_shape, rois, scale):
    """roi_pooling_kernel.cu:128-203 (the gather's visiting rules included); fp64 (B,C,H,W)."""
    grad_out = _f32(grad_out)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    rois = _f32(rois)
    B, C, H, W = feat_shape
    R, _, ph, pw = grad_out.shape
    gin = np.zeros(feat_shape, dtype=np.float64)
    lib().orc_roi_pool_bwd(_p(grad_out, c_f), _p(argmax, c_i), _p(rois, c_f), B, C, H, W, R, ph, pw,
                           ctypes.c_float(scale), _p(gin, c_d))
    return gin


def roi_pool_cpu_nhwc(feat_nhwc, rois, ph, pw, scale):
    """roi_pooling.c semantics (the reference's only CPU pooling path): NHWC, batch 1."""
    feat_nhwc = _f32(feat_nhwc)
    rois = _f32(rois)
    B, H, W, C = feat_nhwc.shape
    R = rois.shape[0]
    out = np.zeros((R, C, ph, pw), dtype=np.float32)
    ok = lib().orc_roi_pool_fwd_cpu_nhwc(_p(feat_nhwc, c_f), _p(rois, c_f), B, H, W, C, R, ph, pw,
                                         ctypes.c_float(scale), _p(out, c_f))
    return out, int(ok)


# ----------------------------------------------------------------------------------------
# box algebra: lib/model/rpn/bbox_transform.py:77-103,125-133,136-166,168-257
# ----------------------------------------------------------------------------------------
def bbox_transform_inv(boxes, deltas):
    boxes = _f32(boxes)
    deltas = _f32(deltas)
    B, N, _ = boxes.shape
    k = deltas.shape[2] // 4
    out = np.empty_like(deltas)
    lib().orc_bbox_transform_inv(_p(boxes, c_f), _p(deltas, c_f), ctypes.c_long(B * N), k, _p(out, c_f))
    return out


def clip_boxes(boxes, im_info):
    boxes = _f32(boxes).copy()
    im_info = _f32(im_info)
    B, N, k4 = boxes.shape
    lib().orc_clip_boxes(_p(boxes, c_f), _p(im_info, c_f), B, ctypes.c_long(N), k4 // 4)
    return boxes


def bbox_overlaps(anchors, gt):
    anchors = _f32(anchors)
    gt = _f32(gt)
    N, K = anchors.shape[0], gt.shape[0]
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_bbox_overlaps(_p(anchors, c_f), _p(gt, c_f), N, K, _p(out, c_f))
    return out


def bbox_overlaps_batch(anchors, gt_boxes):
    """anchors (N,4) | (B,N,4) | (B,N,5: cols 1:5); gt (B,K,>=4)."""
    gt = _f32(np.asarray(gt_boxes)[:, :, :4])
    B, K = gt.shape[:2]
    a = np.asarray(anchors, dtype=np.float32)
    if a.ndim == 2:
        a = np.broadcast_to(a[None, :, :4], (B, a.shape[0], 4))
    elif a.shape[2] != 4:
        a = a[:, :, 1:5]
    a = _f32(a)
    N = a.shape[1]
    out = np.empty((B, N, K), dtype=np.float32)
    lib().orc_bbox_overlaps_batch(_p(a, c_f), _p(gt, c_f), B, N, K, _p(out, c_f))
    return out


# ----------------------------------------------------------------------------------------
# proposal layer: lib/model/rpn/proposal_layer.py:49-161
# ----------------------------------------------------------------------------------------
def proposal_layer(scores, deltas, im_info, anchors, feat_stride, pre_nms_topN, post_nms_topN,
                   nms_thresh, boxes_override=None, return_taps=False):
    scores = _f32(scores)
    deltas = _f32(deltas)
    im_info = _f32(im_info)
    anchors = _f32(anchors)
    B, A2, H, W = scores.shape
    A = A2 // 2
    KA = H * W * A
    pre = pre_nms_topN if 0 < pre_nms_topN < KA else KA
    out = np.zeros((B, post_nms_topN, 5), dtype=np.float32)
    order = np.zeros((B, pre), dtype=np.int32)
    props = np.zeros((B, pre, 4), dtype=np.float32)
    keep = np.zeros((B, post_nms_topN), dtype=np.int32)
    nkeep = np.zeros((B,), dtype=np.int32)
    bo = _f32(boxes_override) if boxes_override is not None else None
    lib().orc_proposal_layer(_p(scores, c_f), _p(deltas, c_f), _p(im_info, c_f), _p(anchors, c_f),
                             B, A, H, W, int(feat_stride), int(pre_nms_topN), int(post_nms_topN),
                             ctypes.c_float(nms_thresh), _p(out, c_f), _p(order, c_i), _p(props, c_f),
                             _p(keep, c_i), _p(nkeep, c_i), _p(bo, c_f))
    if return_taps:
        return out, order, props, keep, nkeep
    return out


# ----------------------------------------------------------------------------------------
# RL refinement: lib/model/Reinforcement/action.py:6-59, lib/datasets/RL_coco_dataset.py:119-137,
# lib/pycocotools/maskApi.c:98-109
# ----------------------------------------------------------------------------------------
def action_table(delta, alpha=1.0):
    """Action.__init__ (action.py:11-22): idx = (dim*len(delta)+j)*2 + sign."""
    num = 4 * len(delta) * 2
    t = np.zeros((num, 4), dtype=np.float32)
    idx = 0
    for i in range(4):
        for j in range(len(delta)):
            t[idx, i] = delta[j] * alpha
            idx += 1
            t[idx, i] = -delta[j] * alpha
            idx += 1
    return t


def bbiou(dt, gt, iscrowd=None):
    """maskApi.c:98-109; returns (m, n) like pycocotools' mask.iou (o is column-major n x m)."""
    dt = np.ascontiguousarray(dt, dtype=np.float64).reshape(-1, 4)
    gt = np.ascontiguousarray(gt, dtype=np.float64).reshape(-1, 4)
    m, n = dt.shape[0], gt.shape[0]
    cr = np.ascontiguousarray(iscrowd, dtype=np.uint8) if iscrowd is not None else None
    o = np.zeros((n, m), dtype=np.float64)
    lib().orc_bbiou(_p(dt, c_d), _p(gt, c_d), ctypes.c_long(m), ctypes.c_long(n), _p(cr, c_u8), _p(o, c_d))
    return o.T.copy()


MODE_COCO, MODE_RCNN = 0, 1


WTRANS_IDENTITY, WTRANS_EXP_ABS = 0, 1


def action_reward_f64(boxes, gt, act, crowd=None, iou_thres=0.0, pos_wratio=1.0, neg_wratio=1.0,
                      wtrans=WTRANS_EXP_ABS):
    """The reference loop (RL_coco_dataset.py:119-137) over float64 xywh boxes, as the json
    holds them: pure numpy over bbiou, small cases only."""
    boxes, gt = np.asarray(boxes, np.float64), np.asarray(gt, np.float64)
    B, N, _ = boxes.shape
    A = act.shape[0]
    reward = np.empty((B, N, A), np.float32)
    label = np.empty((B, N, A), np.float32)
    weight = np.empty((B, N, A), np.float32)
    for b in range(B):
        cr = np.asarray(crowd[b], np.uint8) if crowd is not None else np.zeros(gt.shape[1], np.uint8)
        for n in range(N):
            bbox = boxes[b, n]
            w, h = bbox[2], bbox[3]
            o0 = bbiou(bbox[None], gt[b], cr).max()
            for a in range(A):
                nb = bbox + act[a] * np.array([w, h, w, h])
                d = bbiou(nb[None], gt[b], cr).max() - o0
                pos = d > iou_thres
                reward[b, n, a] = d
                label[b, n, a] = 1.0 if pos else -1.0
                weight[b, n, a] = (np.exp(abs(d)) if wtrans == WTRANS_EXP_ABS else d) * (pos_wratio if pos else neg_wratio)
    return reward, label, weight


def action_reward(boxes, gt, act, crowd=None, ngt=None, mode=MODE_COCO, iou_thres=0.0,
                  pos_wratio=1.0, neg_wratio=1.0, wtrans=WTRANS_EXP_ABS):
    boxes = _f32(boxes)
    gt = _f32(gt)
    act = _f32(act)
    B, N, _ = boxes.shape
    G = gt.shape[1]
    A = act.shape[0]
    cr = np.ascontiguousarray(crowd, dtype=np.uint8) if crowd is not None else None
    ng = np.ascontiguousarray(ngt, dtype=np.int32) if ngt is not None else None
    reward = np.empty((B, N, A), dtype=np.float32)
    label = np.empty((B, N, A), dtype=np.float32)
    weight = np.empty((B, N, A), dtype=np.float32)
    lib().orc_action_reward(_p(boxes, c_f), _p(gt, c_f), _p(cr, c_u8), _p(ng, c_i), _p(act, c_f), B, N, A, G,
                            int(mode), int(wtrans), ctypes.c_float(iou_thres), ctypes.c_float(pos_wratio),
                            ctypes.c_float(neg_wratio), _p(reward, c_f), _p(label, c_f), _p(weight, c_f))
    return reward, label, weight


def move_from_act(bboxes, preds, targets, act, maxk):
    """Action.move_from_act (action.py:25-59).  Visit order = np.flip(np.argsort(pred)) as in
    the reference (:44) with the sort pinned to a stable one: among equal preds the HIGHER flat
    index comes first (numpy's default introsort is stable for short / presorted runs and
    unspecified beyond; tests/golden pins the tied case on the reference's own code)."""
    bboxes = np.array(bboxes, dtype=np.float32, copy=True)
    b, n, _ = bboxes.shape
    A = act.shape[0]
    correct = 0
    for bid in range(b):
        flat = preds[bid].reshape(-1)
        inds = np.flip(np.argsort(flat, kind="stable"), axis=0)
        cnt = 0
        vis = np.zeros(n, dtype=bool)
        for num in inds:
            idx, act_id = num // A, num % A
            if not vis[idx]:
                cnt += 1
                vis[idx] = True
                if targets[bid, idx, act_id] == 1:
                    correct += 1
                    x, y, w, h = bboxes[bid, idx]
                    bboxes[bid, idx] += act[act_id] * np.array([w, h, w, h], dtype=np.float32)
            if cnt >= maxk:
                break
    return bboxes, correct * 100.0 / (b * maxk)


def refine_best_action(rois, reward, label, act):
    """The hot path's refine: Action.move_from_act (action.py:25-59) with maxk = N, the rewards
    as predictions and the labels as targets, on x1y1x2y2 rois (B,N,5) with the +1 convention:
    every box takes its best action (ties -> the HIGHEST action id, the visit order of
    np.flip(np.argsort)) if that action's label is +1.  Returns (refined rois, boxes moved)."""
    rois = _f32(rois)
    B, N, _ = rois.shape
    A = reward.shape[2]
    best = A - 1 - np.argmax(reward[:, :, ::-1], axis=2)
    bi, ni = np.meshgrid(np.arange(B), np.arange(N), indexing="ij")
    take = label[bi, ni, best] == 1
    d = _f32(act)[best]
    b = rois[:, :, 1:5]
    f1 = np.float32(1)
    w = b[..., 2] - b[..., 0] + f1
    h = b[..., 3] - b[..., 1] + f1
    nx, ny = b[..., 0] + d[..., 0] * w, b[..., 1] + d[..., 1] * h
    nw, nh = w + d[..., 2] * w, h + d[..., 3] * h
    moved = np.stack([nx, ny, nx + nw - f1, ny + nh - f1], -1).astype(np.float32)
    refined = rois.copy()
    refined[:, :, 1:5] = np.where(take[..., None], moved, b)
    return refined, int(take.sum())


# ----------------------------------------------------------------------------------------
# the compiled reference itself (oracle/_ref, built from /root/reference in place)
# ----------------------------------------------------------------------------------------
def bbox_transform_batch(ex, gt):
    """bbox_transform.py:44-75 (3-d form), fp32 op by op."""
    ex, gt = _f32(ex), _f32(gt)
    f = np.float32
    ew, eh = ex[..., 2] - ex[..., 0] + f(1), ex[..., 3] - ex[..., 1] + f(1)
    ecx, ecy = ex[..., 0] + f(0.5) * ew, ex[..., 1] + f(0.5) * eh
    gw, gh = gt[..., 2] - gt[..., 0] + f(1), gt[..., 3] - gt[..., 1] + f(1)
    gcx, gcy = gt[..., 0] + f(0.5) * gw, gt[..., 1] + f(0.5) * gh
    return np.stack([(gcx - ecx) / ew, (gcy - ecy) / eh, np.log(gw / ew), np.log(gh / eh)], -1).astype(np.float32)


def proposal_target(all_rois, gt_boxes, fg_keys, bg_u, rois_per_image, fg_rois_per_image, fg_thresh=0.5, bg_hi=0.5,
                    bg_lo=0.1, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2), inside_w=(1, 1, 1, 1)):
    """_ProposalTargetLayer.forward, proposal_target_layer_cascade.py:33-213, with the RNG factored
    out: permutation(n) := stable argsort of the fg candidates' keys, rand(k) := bg_u[:k]."""
    all_rois, gt_boxes = _f32(all_rois), _f32(gt_boxes)
    B, N, _ = all_rois.shape
    app = np.zeros_like(gt_boxes)
    app[:, :, 1:5] = gt_boxes[:, :, :4]
    cand = np.concatenate([all_rois, app], 1)                                   # :41-44
    ov = bbox_overlaps_batch(cand, gt_boxes)                                    # :132
    R = rois_per_image
    rois_b = np.zeros((B, R, 5), np.float32)
    labels_b = np.zeros((B, R), np.float32)
    gt_b = np.zeros((B, R, 5), np.float32)
    status = np.zeros(B, np.int32)
    for i in range(B):
        mx, asg = ov[i].max(1), ov[i].argmax(1)
        labels = gt_boxes[i, asg, 4]
        fg = np.nonzero(mx >= np.float32(fg_thresh))[0]
        bg = np.nonzero((mx < np.float32(bg_hi)) & (mx >= np.float32(bg_lo)))[0]
        u = bg_u[i].astype(np.float64)
        if fg.size > 0 and bg.size > 0:
            k = min(fg_rois_per_image, fg.size)
            perm = np.argsort(fg_keys[i][fg], kind="stable")
            fg = fg[perm[:k]]
            bg = bg[np.floor(u[:R - k] * bg.size).astype(np.int64)]
            nfg = k
        elif fg.size > 0:
            fg = fg[np.floor(u[:R] * fg.size).astype(np.int64)]
            bg, nfg = bg[:0], R
        elif bg.size > 0:
            bg = bg[np.floor(u[:R] * bg.size).astype(np.int64)]
            fg, nfg = fg[:0], 0
        else:
            status[i] = 1
            rois_b[i, :, 0] = i
            continue
        keep = np.concatenate([fg, bg])
        labels_b[i] = labels[keep]
        labels_b[i, nfg:] = 0
        rois_b[i] = cand[i, keep]
        rois_b[i, :, 0] = i
        gt_b[i] = gt_boxes[i, asg[keep]]
    tg = bbox_transform_batch(rois_b[:, :, 1:5], gt_b[:, :, :4])
    if means is not None:
        tg = ((tg - _f32(means)) / _f32(stds)).astype(np.float32)
    pos = labels_b > 0
    targets = np.where(pos[..., None], tg, np.float32(0)).astype(np.float32)
    targets[status == 1] = 0
    inside = np.where(pos[..., None], _f32(inside_w), np.float32(0)).astype(np.float32)
    return rois_b, labels_b, targets, inside, (inside > 0).astype(np.float32), status


def anchor_target(gt_boxes, im_info, anchors, keys, H, W, feat_stride=16, pos_ov=0.7, neg_ov=0.3, clobber=False,
                  fg_fraction=0.5, batchsize=256, inside_w=1.0):
    """_AnchorTargetLayer.forward, anchor_target_layer.py:48-192 (RPN_POSITIVE_WEIGHT < 0), RNG
    factored out: permutation(n) := stable argsort of the members' keys."""
    gt_boxes, anchors = _f32(gt_boxes), _f32(anchors)
    B, A = gt_boxes.shape[0], anchors.shape[0]
    sx, sy = np.meshgrid(np.arange(W) * feat_stride, np.arange(H) * feat_stride)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose().astype(np.float32)
    allanc = (anchors[None] + shifts[:, None]).reshape(-1, 4)
    total = allanc.shape[0]
    keep = ((allanc[:, 0] >= 0) & (allanc[:, 1] >= 0) & (allanc[:, 2] < int(im_info[0][1])) &
            (allanc[:, 3] < int(im_info[0][0])))
    inds = np.nonzero(keep)[0]
    anc = allanc[inds]
    labels = np.full((B, inds.size), -1, np.float32)
    ov = bbox_overlaps_batch(anc, gt_boxes)
    mx, amx = ov.max(2), ov.argmax(2)
    gmx = ov.max(1)
    if not clobber:
        labels[mx < np.float32(neg_ov)] = 0
    gmx[gmx == 0] = np.float32(1e-5)
    k = (ov == gmx[:, None, :]).sum(2)
    labels[k > 0] = 1
    labels[mx >= np.float32(pos_ov)] = 1
    if clobber:
        labels[mx < np.float32(neg_ov)] = 0
    num_fg = int(fg_fraction * batchsize)
    sum_fg, sum_bg = (labels == 1).sum(1), (labels == 0).sum(1)
    for i in range(B):
        kk = keys[i][inds]
        if sum_fg[i] > num_fg:
            fg = np.nonzero(labels[i] == 1)[0]
            perm = np.argsort(kk[fg], kind="stable")
            labels[i][fg[perm[:fg.size - num_fg]]] = -1
        num_bg = batchsize - sum_fg[i]
        if sum_bg[i] > num_bg:
            bg = np.nonzero(labels[i] == 0)[0]
            perm = np.argsort(kk[bg], kind="stable")
            labels[i][bg[perm[:bg.size - num_bg]]] = -1
    tg = bbox_transform_batch(np.broadcast_to(anc[None], (B,) + anc.shape), gt_boxes[np.arange(B)[:, None], amx][:, :, :4])
    iw = np.zeros_like(labels)
    iw[labels == 1] = np.float32(inside_w)
    num_examples = (labels[B - 1] >= 0).sum()                                 # :158, i leaked from the loop
    w = np.float32(1.0) / np.float32(num_examples)
    ow = np.zeros_like(labels)
    ow[labels == 1] = w
    ow[labels == 0] = w

    def unmap(d, fill):
        shape = (B, total) + d.shape[2:]
        out = np.full(shape, fill, np.float32)
        out[:, inds] = d
        return out
    L = unmap(labels, -1).reshape(B, H, W, A).transpose(0, 3, 1, 2).reshape(B, 1, A * H, W)
    T = unmap(tg, 0).reshape(B, H, W, A * 4).transpose(0, 3, 1, 2)
    IW = np.repeat(unmap(iw, 0)[:, :, None], 4, 2).reshape(B, H, W, 4 * A).transpose(0, 3, 1, 2)
    OW = np.repeat(unmap(ow, 0)[:, :, None], 4, 2).reshape(B, H, W, 4 * A).transpose(0, 3, 1, 2)
    return np.ascontiguousarray(L), np.ascontiguousarray(T), np.ascontiguousarray(IW), np.ascontiguousarray(OW)


def affine_grid(rois, H, W, g, align_corners=True):
    """_affine_grid_gen, lib/model/utils/net_utils.py:143-165 (theta from roi / 16; base grid
    linspace(-1, 1, g), or its torch >= 1.3 default (2j+1)/g - 1) -> (R, g, g, 2) = (x, y)."""
    rois = _f32(rois)
    f = np.float32
    x1, y1, x2, y2 = (rois[:, i] / f(16.0) for i in (1, 2, 3, 4))
    t00 = (x2 - x1) / f(W - 1)
    t02 = (x1 + x2 - f(W) + f(1)) / f(W - 1)
    t11 = (y2 - y1) / f(H - 1)
    t12 = (y1 + y2 - f(H) + f(1)) / f(H - 1)
    if align_corners:
        base = np.linspace(-1.0, 1.0, g).astype(np.float32) if g > 1 else np.array([-1.0], np.float32)
    else:
        base = ((2 * np.arange(g) + 1).astype(np.float32) / f(g) - f(1)).astype(np.float32)
    x = (base[None, None, :].astype(np.float64) * t00[:, None, None] + t02[:, None, None]).astype(np.float32)
    y = (base[None, :, None].astype(np.float64) * t11[:, None, None] + t12[:, None, None]).astype(np.float32)
    return np.stack([np.broadcast_to(x, (len(rois), g, g)), np.broadcast_to(y, (len(rois), g, g))], 3).astype(np.float32)


def _crop_taps(grid_yx, H, W):
    f = np.float32
    yf, xf = grid_yx[..., 0].astype(np.float32), grid_yx[..., 1].astype(np.float32)
    xc = ((xf + f(1)) * f(W - 1)) / f(2)      # getTopLeft, roi_crop_cuda_kernel.cu:11-23
    yc = ((yf + f(1)) * f(H - 1)) / f(2)
    x0, y0 = np.floor(xc), np.floor(yc)
    xw, yw = f(1) - (xc - x0), f(1) - (yc - y0)
    return x0.astype(np.int64), y0.astype(np.int64), xw.astype(np.float32), yw.astype(np.float32)


def roi_crop(feat, grid_yx):
    """BilinearSamplerBHWD forward, roi_crop_cuda_kernel.cu:47-118; image of roi r = r // (R // B)."""
    feat, grid_yx = _f32(feat), _f32(grid_yx)
    B, C, H, W = feat.shape
    R, gh, gw, _ = grid_yx.shape
    per = R // B
    x0, y0, xw, yw = _crop_taps(grid_yx, H, W)
    out = np.zeros((R, C, gh, gw), np.float32)
    f = np.float32
    for r in range(R):
        b = r // per
        if b >= B:
            continue
        acc = np.zeros((C, gh, gw), np.float32)
        for dy, dx, wgt in ((0, 0, xw[r] * yw[r]), (0, 1, (f(1) - xw[r]) * yw[r]), (1, 0, xw[r] * (f(1) - yw[r])),
                            (1, 1, (f(1) - xw[r]) * (f(1) - yw[r]))):
            yy, xx = y0[r] + dy, x0[r] + dx
            ok = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            v = feat[b][:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[None]
            acc = (acc + wgt.astype(np.float32)[None] * v.astype(np.float32)).astype(np.float32)
        out[r] = acc
    return out


def roi_crop_bwd(grad_out, grid_yx, feat_shape):
    """:120-199, accumulated in fp64 (the reference's atomicAdd order is not fixed)."""
    grad_out, grid_yx = _f32(grad_out), _f32(grid_yx)
    B, C, H, W = feat_shape
    R, gh, gw, _ = grid_yx.shape
    per = R // B
    x0, y0, xw, yw = _crop_taps(grid_yx, H, W)
    gin = np.zeros((B, C, H, W), np.float64)
    f = np.float32
    for r in range(R):
        b = r // per
        if b >= B:
            continue
        for dy, dx, wgt in ((0, 0, xw[r] * yw[r]), (0, 1, (f(1) - xw[r]) * yw[r]), (1, 0, xw[r] * (f(1) - yw[r])),
                            (1, 1, (f(1) - xw[r]) * (f(1) - yw[r]))):
            yy, xx = y0[r] + dy, x0[r] + dx
            ok = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            contrib = (wgt.astype(np.float32)[None] * grad_out[r]).astype(np.float32)
            for c in range(C):
                np.add.at(gin[b, c], (np.clip(yy, 0, H - 1)[ok], np.clip(xx, 0, W - 1)[ok]), contrib[c][ok])
    return gin


def rl_labels(dets, det_cat, ndet, gt, gt_cat, crowd, ngt, act, iou_thres=0.0, pos_wratio=1.0, neg_wratio=1.0,
              wtrans=WTRANS_EXP_ABS):
    """Label tensor (B,N,A,3) = (act_id, label, weight) of a collated RL batch:
    lib/datasets/RL_coco_dataset.py:107-137 per box (category-specific gt, none -> [[0,0,0,0]],
    bbIou fp64) + the zero padding of lib/datasets/RL_coco_loader.py:66-72."""
    dets, gt = np.asarray(dets, np.float64), np.asarray(gt, np.float64)
    B, N = dets.shape[:2]
    A = act.shape[0]
    out = np.zeros((B, N, A, 3), np.float32)
    for b in range(B):
        for n in range(int(ndet[b])):
            bbox = dets[b, n, :4]
            w, h = bbox[2], bbox[3]
            sel = [g for g in range(int(ngt[b])) if int(gt_cat[b, g]) == int(det_cat[b, n])]
            gtb = gt[b, sel] if sel else np.zeros((1, 4))
            cr = np.asarray([crowd[b, g] for g in sel], np.uint8) if sel else np.zeros(1, np.uint8)
            o0 = bbiou(bbox[None], gtb, cr).max()
            for a in range(A):
                nb = bbox + act[a].astype(np.float64) * np.array([w, h, w, h])
                d = bbiou(nb[None], gtb, cr).max() - o0
                pos = d > iou_thres
                wt = np.exp(abs(d)) if wtrans == WTRANS_EXP_ABS else d
                out[b, n, a] = (a, 1.0 if pos else -1.0, wt * (pos_wratio if pos else neg_wratio))
    return out


def detect_postprocess(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=0.3, max_per_image=100,
                       stds=None, means=None, class_agnostic=False):
    """Test-time post-processing, RCNN_bases/test_net.py:244-307, for a batch (the reference
    runs batch 1).  Returns all_boxes[b][j] = (k, 5) float32 arrays [x1,y1,x2,y2,score]; class 0
    is always empty.  Sort ties: lower roi index first (torch.sort order is unpinned)."""
    rois, cls_prob, im_info = _f32(rois), _f32(cls_prob), _f32(im_info)
    B, N, K = cls_prob.shape
    out = []
    for b in range(B):
        boxes = rois[b, :, 1:5]
        if bbox_pred is not None:
            d = _f32(bbox_pred)[b].reshape(N, -1)
            if stds is not None:  # :251-260
                d = (d.reshape(-1, 4) * np.asarray(stds, np.float32) + np.asarray(means, np.float32)).reshape(N, -1)
            pred = bbox_transform_inv(boxes[None], d[None])           # :262
            pred = clip_boxes(pred, im_info[b:b + 1])[0]              # :263
        else:
            pred = np.tile(boxes, (1, K))                             # :266
        pred = (pred / im_info[b, 2]).astype(np.float32)              # :268
        per_class = [np.zeros((0, 5), np.float32)]
        for j in range(1, K):                                         # :277-297
            inds = np.nonzero(cls_prob[b, :, j] > np.float32(thresh))[0]
            if inds.size == 0:
                per_class.append(np.zeros((0, 5), np.float32))
                continue
            sc = cls_prob[b, inds, j]
            order = np.argsort(-sc.astype(np.float64), kind="stable")
            cb = pred[inds] if class_agnostic else pred[inds][:, 4 * j:4 * j + 4]
            dets = np.concatenate([cb, sc[:, None]], 1).astype(np.float32)[order]
            keep = nms(dets, nms_thresh)
            per_class.append(dets[keep])
        if max_per_image > 0:                                         # :299-307
            sc = np.concatenate([c[:, 4] for c in per_class[1:]])
            if sc.size > max_per_image:
                th = np.sort(sc)[-max_per_image]
                per_class = [c[c[:, 4] >= th] if i else c for i, c in enumerate(per_class)]
        out.append(per_class)
    return out


def ref_maskapi():
    path = os.path.join(_HERE, "_ref", "libmaskapi.so")
    return ctypes.CDLL(path) if os.path.exists(path) else None


def ref_bbiou(dt, gt, iscrowd=None):
    """The reference's own bbIou (maskApi.c compiled unchanged)."""
    l = ref_maskapi()
    dt = np.ascontiguousarray(dt, dtype=np.float64).reshape(-1, 4)
    gt = np.ascontiguousarray(gt, dtype=np.float64).reshape(-1, 4)
    m, n = dt.shape[0], gt.shape[0]
    cr = np.ascontiguousarray(iscrowd, dtype=np.uint8) if iscrowd is not None else None
    o = np.zeros((n, m), dtype=np.float64)
    l.bbIou(_p(dt, c_d), _p(gt, c_d), ctypes.c_size_t(m), ctypes.c_size_t(n), _p(cr, c_u8), _p(o, c_d))
    return o.T.copy()


def ref_legacy():
    """The reference's legacy CUDA kernels compiled unchanged for sm_100a (GPU box only)."""
    path = os.path.join(_HERE, "_ref", "libref_legacy.so")
    return ctypes.CDLL(path) if os.path.exists(path) else None
