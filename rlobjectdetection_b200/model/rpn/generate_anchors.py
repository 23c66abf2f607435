"""Base anchor table (reference: lib/model/rpn/generate_anchors.py:45-105): every aspect ratio
of the (0, 0, base-1, base-1) window, np.round-ed to whole pixels, then every scale of each;
ratio-major, scale-minor, float64.  Init-time host code; the per-location shifts are generated
arithmetically inside the proposal kernel."""
import numpy as np


def _centre_form(box):
    w = box[2] - box[0] + 1.0
    h = box[3] - box[1] + 1.0
    return w, h, box[0] + 0.5 * (w - 1.0), box[1] + 0.5 * (h - 1.0)


def _corner_form(ws, hs, cx, cy):
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    half_w, half_h = 0.5 * (ws - 1.0), 0.5 * (hs - 1.0)
    return np.concatenate([cx - half_w, cy - half_h, cx + half_w, cy + half_h], axis=1)


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w, h, cx, cy = _centre_form(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt(w * h / ratios))  # half-to-even, like the reference (:91)
    hs = np.round(ws * ratios)
    per_ratio = _corner_form(ws, hs, cx, cy)
    blocks = []
    for box in per_ratio:
        w, h, cx, cy = _centre_form(box)
        blocks.append(_corner_form(w * scales, h * scales, cx, cy))
    return np.concatenate(blocks, axis=0)
