#!/bin/bash
# A/B of the forward RoIAlign call: programmatic dependent launch of the list / pooling kernels and the split
# last wave against the plain launches (RLOD_NO_PDL, RLOD_NO_SPLIT), whole call, cold L2.
TAG=${1:-ab}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi_align or graphed or step or c4_bench" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for cfg in C2 C4; do
  python tools/time_op.py align_fwd $cfg 60
  RLOD_NO_PDL=1 python tools/time_op.py align_fwd $cfg 60
  RLOD_NO_SPLIT=1 python tools/time_op.py align_fwd $cfg 60
  RLOD_NO_PDL=1 RLOD_NO_SPLIT=1 python tools/time_op.py align_fwd $cfg 60
done 2>&1 | tee gpurun_out/ab_fwd_$TAG.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ops > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200
RLOD_NO_PDL=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ops > gpurun_out/bench_${TAG}_nopdl.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_${TAG}_nopdl.log | cut -c1-200
