#!/bin/bash
# One gpurun call: GPU parity tests, bench, ncu launch list + one full capture of the dominant kernel.
# usage: tools/gpu_check.sh TAG [steps]
TAG=${1:-x}; STEPS=${2:-20}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_$TAG.log
[ $rc -ne 0 ] && { grep -n "Error\|error\|assert" gpurun_out/pytest_$TAG.log | head -20; }
python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-2500
if [ "$3" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k_align8_fwd -s 6 -c 1 -o gpurun_out/prof_align_fwd_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
