"""nms_gpu(dets, thresh): device NMS with the reference's return convention
(lib/model/nms/nms_gpu.py:7-12): an int32 (num_kept, 1) tensor of row indices."""
from .. import _backend as be


def nms_gpu(dets, thresh):
    keep, num = be.nms_padded(dets, thresh)
    # slicing to the kept count needs the count on the host -- the one sync the reference
    # also has (`keep[:num_out[0]]`); use nms_padded / nms_batched to stay asynchronous
    return keep[: int(num.item())].view(-1, 1)
