#!/bin/bash
# One gpurun call for the round's evidence: parity tests, bench lines, ncu launch list and kernel captures.
TAG=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align8_bwd_own -s 2 -c 1 -o gpurun_out/prof_bwd_own_C2_$TAG -f python tools/prof_op.py align_bwd C2 > /dev/null 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_align8_bwd_own -s 2 -c 1 -o gpurun_out/prof_bwd_own_C4_$TAG -f python tools/prof_op.py align_bwd C4 > /dev/null 2>&1; echo rc=$?
ncu --set full --clock-control none -k regex:"k_nms_scan3|k_nms_mask_rm" -s 4 -c 2 -o gpurun_out/prof_nms_C1_$TAG -f python tools/prof_op.py proposal C1 > /dev/null 2>&1; echo rc=$?
