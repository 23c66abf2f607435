"""Discrete box-delta actions (reference: lib/model/Reinforcement/action.py:6-59).

`Action(delta, alpha, iou_thres, wtrans)` builds the same (4*len(delta)*2, 4) fp32 table:
for every box dimension and every delta magnitude a +delta and a -delta row.
`move_from_act` keeps the reference's numpy signature (arrays in, (boxes, precision) out) but
runs on the GPU (rlod_move_from_act); `move_from_act_cuda` is the asynchronous tensor form.
Order of visits: pred descending, ties by HIGHER flat (box, action) index first -- what the
reference's np.flip(np.argsort(pred)) gives wherever numpy's sort is stable (action.py:44).

`Action.wtrans` (default Identify, action.py:7-10; Config.act_wtrans = exp(|x|), config.py:48-51)
is honoured by the label kernels: `wtrans_code(action)` recognises the two forms the reference
ships (by identity, or by probing an unknown callable at a few points) and maps them to the
kernel's selector; any other callable is applied to the raw delta_iou the kernel hands back."""
import math

import numpy as np
import torch

from .. import _backend as be


def Identify(x):
    return x


def exp_abs(x):
    """Config.act_wtrans of the reference (config.py:48-51)."""
    return math.exp(math.fabs(x))


_PROBE = (-0.75, -0.125, 0.0, 0.3, 0.9)


def wtrans_code(action):
    """RLOD_WTRANS_* selector of action.wtrans (be.WTRANS_RAW: a custom callable)."""
    f = action.wtrans
    if f is Identify:
        return be.WTRANS_IDENTITY
    if f is exp_abs:
        return be.WTRANS_EXP_ABS
    try:
        vals = [float(f(x)) for x in _PROBE]
    except Exception:  # noqa: BLE001 -- not a scalar callable: treat as custom
        return be.WTRANS_RAW
    if vals == [math.exp(math.fabs(x)) for x in _PROBE]:
        return be.WTRANS_EXP_ABS
    if vals == list(_PROBE):
        return be.WTRANS_IDENTITY
    return be.WTRANS_RAW


def apply_custom_wtrans(action, raw, label, pos_wratio, neg_wratio):
    """weight = wtrans(delta_iou) * (pos|neg)_wratio for a callable the kernels do not know
    (RL_coco_dataset.py:128-135): tried on the device tensor first; a scalar-only Python
    callable is mapped element by element on the host (it is the user's own Python)."""
    ratio = torch.where(label > 0, torch.full_like(raw, float(pos_wratio)), torch.full_like(raw, float(neg_wratio)))
    try:
        w = action.wtrans(raw)
        if not torch.is_tensor(w) or w.shape != raw.shape:
            raise TypeError
    except Exception:  # noqa: BLE001
        host = np.vectorize(action.wtrans, otypes=[np.float64])(raw.detach().cpu().double().numpy())
        w = torch.from_numpy(host).to(device=raw.device, dtype=raw.dtype)
    return w.to(raw.dtype) * ratio


class Action:
    def __init__(self, delta, alpha=1., iou_thres=0, wtrans=None):
        self.delta = delta
        self.alpha = alpha
        self.iou_thres = iou_thres
        self.wtrans = Identify if wtrans is None else wtrans
        mags = np.asarray(delta, dtype=np.float64) * alpha
        per_dim = np.stack([mags, -mags], axis=1).reshape(-1)  # +d0, -d0, +d1, -d1, ...
        self.num_acts = 4 * per_dim.size
        self.actDeltas = np.zeros((self.num_acts, 4), dtype=np.float32)
        for dim in range(4):
            self.actDeltas[dim * per_dim.size:(dim + 1) * per_dim.size, dim] = per_dim
        self._table = {}

    def table(self, device):
        """actDeltas as a device tensor (cached per device)."""
        key = str(device)
        if key not in self._table:
            self._table[key] = torch.from_numpy(self.actDeltas).to(device)
        return self._table[key]

    def move_from_act_cuda(self, bboxes, preds, targets, maxk):
        """bboxes (b,n,4) xywh fp32 CUDA tensor, updated in place; returns (bboxes, moved)
        where `moved` is a device int32 count (precision = moved*100/(b*maxk))."""
        return bboxes, be.move_from_act(bboxes, preds, targets, self.table(bboxes.device), maxk)

    def move_from_act(self, bboxes, preds, targets, maxk, device="cuda"):
        assert preds.shape == targets.shape
        assert bboxes.ndim == 3 and preds.ndim == 3
        assert preds.shape[0] == bboxes.shape[0] and preds.shape[1] == bboxes.shape[1]
        if torch.is_tensor(bboxes):
            out, moved = self.move_from_act_cuda(bboxes, preds, targets, maxk)
            return out, moved.item() * 100. / (bboxes.shape[0] * maxk)
        dev = torch.device(device)
        bb = torch.from_numpy(np.ascontiguousarray(bboxes, dtype=np.float32)).to(dev)
        pr = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).to(dev)
        tg = torch.from_numpy(np.ascontiguousarray(targets, dtype=np.float32)).to(dev)
        _, moved = self.move_from_act_cuda(bb, pr, tg, maxk)
        bboxes[...] = bb.cpu().numpy()  # the reference mutates its argument too (:55)
        return bboxes, moved.item() * 100. / (bboxes.shape[0] * maxk)
