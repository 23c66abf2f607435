"""Which host call is slow in the first step after a device synchronize? (bench.py's timed region starts there)"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlobjectdetection_b200.hotpath import DetectRefineStep
from rlobjectdetection_b200.model import _backend as be
from rlobjectdetection_b200.model.utils.config import cfg
dev = torch.device("cuda", 0)
s_, d_, i_, f_, g_ = [t.to(dev) for t in bench.make_inputs(100, bench.IMAGES_PER_GPU)]
step = DetectRefineStep(bench.STRIDE, bench.SCALES, bench.RATIOS, "TEST", bench.POOL, bench.ACT_DELTA, backward=False)
cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = bench.PRE, bench.POST, bench.NMS_T
cur, light = torch.cuda.current_stream(), step._light_stream(dev)
segs = {}
def lap(name, t):
    now = time.perf_counter(); segs.setdefault(name, []).append(1e3 * (now - t)); return now
def one(rec):
    have_rois, have_refined, lt = rec
    t = time.perf_counter()
    cur.wait_event(have_rois); t = lap("wait_event", t)
    feat, rois = be.f32c(f_), be.f32c(lt[0].view(-1, 5)); t = lap("f32c", t)
    out = torch.empty(rois.size(0), 1024, 7, 7, dtype=torch.float32, device=dev); t = lap("empty 1.4GB", t)
    ws = be.workspace(be.lib().rlod_roi_align_workspace_bytes(24, rois.size(0), 7, 7, 1), dev); t = lap("empty ws", t)
    be.check(be.lib().rlod_roi_align_forward(be.ptr(feat), be.ptr(rois), 24, 1024, 50, 75, rois.size(0), 7, 7, 1 / 16.0, 1, 0,
                                            be.ptr(out), be.ptr(ws), ws.numel(), be.stream_of(feat)), "x"); t = lap("C call align", t)
    nxt = step._light_work(cur, light, s_, d_, i_, g_, True); t = lap("light work", t)
    cur.wait_event(have_refined)
    p2 = step.align(f_, lt[4].view(-1, 5)); t = lap("align 2", t)
    c = torch.cuda.Event(); c.record(cur); step._inflight.append((c, lt)); t = lap("event", t)
    return nxt
rec = step._light_work(cur, light, s_, d_, i_, g_, True)
for _ in range(3): rec = one(rec)
segs.clear()
for trial in range(30):
    torch.cuda.synchronize()
    rec = one(rec)          # first step after a sync
    rec = one(rec); rec = one(rec)
for k, v in segs.items():
    first = v[0::3]; rest = v[1::3] + v[2::3]
    print(f"{k:14s} first-after-sync: median {sorted(first)[len(first)//2]:.3f} max {max(first):.3f} ms | later: median {sorted(rest)[len(rest)//2]:.3f} max {max(rest):.3f} ms")
