"""Discrete box-delta actions (reference: lib/model/Reinforcement/action.py:6-59).

`Action(delta, alpha, iou_thres, wtrans)` builds the same (4*len(delta)*2, 4) fp32 table:
for every box dimension and every delta magnitude a +delta and a -delta row.
`move_from_act` keeps the reference's numpy signature (arrays in, (boxes, precision) out) but
runs on the GPU (rlod_move_from_act); `move_from_act_cuda` is the asynchronous tensor form.
Order of visits: pred descending, ties by lower flat (box, action) index -- numpy's quicksort
argsort leaves ties unspecified in the reference."""
import numpy as np
import torch

from .. import _backend as be


def Identify(x):
    return x


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
