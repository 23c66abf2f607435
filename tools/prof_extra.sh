#!/bin/bash
# ncu captures of the RoIPool forward (C2) and of the tiled RoIAlign forward (large map), then the final bench line
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_roi_pool7_fwd_planes -s 2 -c 1 -o gpurun_out/prof_pool_fwd_r02 -f python tools/prof_op.py pool_fwd C2 > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_align8_fwd_walk2 -s 2 -c 1 -o gpurun_out/prof_align_fwd_tiled_r02 -f python tools/prof_op.py align_fwd BIGD > gpurun_out/ncu_tiled.log 2>&1; echo "ncu tiled rc=$?"
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02j.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r02j.log | cut -c1-300
