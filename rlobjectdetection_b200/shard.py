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


class PipelinedGather:
    """The same all-gather, taken off the step's critical path: `submit(packed)` copies the step's packed rows
    into one of two staging buffers on the current stream and starts the collective asynchronously (NCCL runs
    it on its own stream), so the next step's kernels are enqueued without waiting for it; a staging pair is
    reused two steps later, after its collective has been waited for.  `result()` waits for the newest
    collective on the current stream and returns its (global_batch, N, 5+A) tensor.  Equal shards only
    (global_batch divisible by the world size); anything else falls back to gather_results."""

    def __init__(self, packed_like, global_batch, group=None):
        self.group, self.global_batch = group, global_batch
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.equal = self.world > 1 and global_batch % self.world == 0
        if self.equal:
            b, n, k = packed_like.shape
            self.stage = [torch.empty_like(packed_like) for _ in range(2)]
            self.out = [packed_like.new_empty(self.world * b, n, k) for _ in range(2)]
        self.work = [None, None]
        self.i = 0
        self.last = None

    def submit(self, packed):
        if not self.equal:
            self.last = (gather_results(packed, self.global_batch, self.group), None)
            return
        i = self.i & 1
        self.i += 1
        if self.work[i] is not None:
            self.work[i].wait()  # stream-ordered: the collective that read stage[i] / wrote out[i] two steps ago
        self.stage[i].copy_(packed)
        self.work[i] = dist.all_gather_into_tensor(self.out[i], self.stage[i], group=self.group, async_op=True)
        self.last = (self.out[i], self.work[i])

    def result(self):
        out, work = self.last
        if work is not None:
            work.wait()
        return out

    def drain(self):
        for w in self.work:
            if w is not None:
                w.wait()
