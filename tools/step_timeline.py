"""Where the caller's stream spends a C4 step: CUDA events between the stages of DetectRefineStep
(the same calls, in the same order, as hotpath.DetectRefineStep.__call__ with next_inputs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlobjectdetection_b200.hotpath import DetectRefineStep
from rlobjectdetection_b200.model.utils.config import cfg
from rlobjectdetection_b200.shard import gather_results, pack_results
dev = torch.device("cuda", 0)
s_, d_, i_, f_, g_ = [t.to(dev) for t in bench.make_inputs(100, bench.IMAGES_PER_GPU)]
step = DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA, backward=False)
cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
cur, light = torch.cuda.current_stream(), step._light_stream(dev)
names = ["wait rois", "plan+align 1", "wait refined", "plan+align 2", "copies+pack"]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(cur); return e
def one(rec, marks):
    have_rois, have_refined, lt = rec
    m = [ev()]
    cur.wait_event(have_rois); m.append(ev())
    pooled = step.align(f_, lt[0].view(-1, 5)); m.append(ev())
    nxt = step._light_work(cur, light, s_, d_, i_, g_, True)
    cur.wait_event(have_refined); m.append(ev())
    pooled2 = step.align(f_, lt[4].view(-1, 5)); m.append(ev())
    outs = [t.clone() for t in lt]
    gather_results(pack_results(outs[4], outs[1], 0), bench.IMAGES_PER_GPU)
    c = torch.cuda.Event(); c.record(cur); step._inflight.append((c, lt)); m.append(ev())
    marks.append(m)
    return nxt
rec = step._light_work(cur, light, s_, d_, i_, g_, True)
for _ in range(5): rec = one(rec, [])
torch.cuda.synchronize()
marks = []
for _ in range(50): rec = one(rec, marks)
torch.cuda.synchronize()
tot = marks[-1][-1]; first = marks[0][0]
print(f"step {first.elapsed_time(tot) / 50 * 1e3:.1f} us")
for k, n in enumerate(names):
    print(f"  {n:14s} {sum(m[k].elapsed_time(m[k + 1]) for m in marks) / 50 * 1e3:8.1f} us")
print(f"  between steps  {sum(marks[j][-1].elapsed_time(marks[j + 1][0]) for j in range(49)) / 49 * 1e3:8.1f} us")
