"""Multi-GPU parity on hardware: BASELINE config 4's global batch of 24 images sharded over N ranks
(12 / 6 / 3 images per GPU), one process per GPU, NCCL all-gather of the packed detections || rewards -- the gathered
tensor must equal the single-rank run of the whole batch BIT FOR BIT, and every rank's pooled features must equal the
single-rank rows of its images.  Skips with fewer than 2 GPUs (run under `gpurun --gpus 2 ...`; logs in profiles/)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp, graphed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    from rlobjectdetection_b200.hotpath import DetectRefineStep
    from rlobjectdetection_b200.model.utils.config import cfg
    from rlobjectdetection_b200.shard import gather_results, shard_bounds
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
        gb = bench.IMAGES
        full = bench.make_inputs(100, gb)                      # the bench's own global batch
        lo, hi = shard_bounds(gb, rank, world)
        mine = [t[lo:hi].contiguous().to(dev) for t in full]

        def make_step(first):
            return DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA,
                                    backward=False, outputs=("packed",), first_image=first)
        step = make_step(lo)
        if graphed:
            gs = step.capture(*mine, next_inputs=(mine[0], mine[1], mine[2], mine[4]))
            gs.prime()
            out = gs.replay()
        else:
            out = step(*mine)
        gathered = gather_results(out["packed"], gb)
        torch.cuda.synchronize()
        ok = tuple(gathered.shape) == (gb, bench.POST, 5 + 16)
        # every rank checks against its own single-rank run of the whole batch
        whole = [t.to(dev) for t in full]
        ref = make_step(0)(*whole)
        torch.cuda.synchronize()
        ok = ok and torch.equal(gathered, ref["packed"])
        n = bench.POST
        ok = ok and torch.equal(out["pooled"], ref["pooled"][lo * n:hi * n])
        ok = ok and torch.equal(out["pooled_refined"], ref["pooled_refined"][lo * n:hi * n])
        ok = ok and bool((gathered[:, :, 0] == torch.arange(gb, device=dev, dtype=torch.float32)[:, None]).all())
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graphed", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_gather_equals_single_rank(tmp_path, world, graphed):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() * 7 + world + 3 * graphed) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), graphed), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1"] * world
