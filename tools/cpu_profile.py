"""cProfile of the host side of one C4 step (where do the 0.7-1.1 ms of enqueue time go)."""
import cProfile, pstats, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlobjectdetection_b200.hotpath import DetectRefineStep
from rlobjectdetection_b200.model.utils.config import cfg
from rlobjectdetection_b200.shard import gather_results, pack_results
dev = torch.device("cuda", 0)
dev_in = [t.to(dev) for t in bench.make_inputs(100, bench.IMAGES_PER_GPU)]
step = DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA, backward=False)
cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
def one():
    out = step(*dev_in, inputs_ready=True)
    return gather_results(pack_results(out["refined"], out["reward"], 0), bench.IMAGES_PER_GPU)
for _ in range(5): one()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): one()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
