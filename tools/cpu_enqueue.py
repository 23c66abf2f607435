import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from rlobjectdetection_b200.hotpath import DetectRefineStep
from rlobjectdetection_b200.model.utils.config import cfg
from rlobjectdetection_b200.shard import gather_results, pack_results
dev = torch.device("cuda", 0)
host = bench.make_inputs(100, bench.IMAGES_PER_GPU)
dev_in = [t.to(dev) for t in host]
step = DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA, backward=False)
cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
def one():
    out = step(*dev_in, inputs_ready=True)
    packed = pack_results(out["refined"], out["reward"], 0)
    return gather_results(packed, bench.IMAGES_PER_GPU)
for _ in range(5): one()
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter()
    for _ in range(50): one()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0)/50:.3f} ms/step   total {1e3*(t2-t0)/50:.3f} ms/step")
