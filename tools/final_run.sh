#!/bin/bash
# One gpurun call for the round's evidence: parity tests, bench lines, ncu launch list and kernel captures.
TAG=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align8_fwd_walk2 -s 6 -c 1 -o gpurun_out/prof_align_fwd_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ops > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu fwd rc=$?"
