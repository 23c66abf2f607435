"""Multi-GPU plumbing: the hot path is independent per image, so a global batch is cut into
contiguous per-rank shards and the only exchange is ONE all-gather of the packed per-image
results (rois || rewards) at the end -- NCCL over NVLink on GPUs, gloo on CPU in tests.
(reference: single-process nn.DataParallel scatter/gather, RCNN_bases/trainval_net.py:292-293)"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world_size):
    """Contiguous split; the first (global_batch % world_size) ranks take one extra image."""
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_results(rois, reward, first_image):
    """rois (b,N,5) with rank-local image indices + reward (b,N,A) -> (b,N,5+A) fp32 whose
    column 0 is the GLOBAL image index, so the gathered tensor equals a single-GPU run."""
    packed = torch.cat([rois, reward], dim=2)
    packed[:, :, 0] += float(first_image)
    return packed.contiguous()


def gather_results(packed, global_batch, group=None):
    """All ranks receive (global_batch, N, 5+A).  One collective; equal shards use
    all_gather_into_tensor, ragged ones pad to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed
    world = dist.get_world_size(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    cmax = max(counts)
    b, n, k = packed.shape
    if b < cmax:
        pad = packed.new_zeros(cmax - b, n, k)
        packed = torch.cat([packed, pad], 0)
    out = packed.new_empty(world * cmax, n, k)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return out
    return torch.cat([out[r * cmax:r * cmax + counts[r]] for r in range(world)], 0)
