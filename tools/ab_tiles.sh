#!/bin/bash
# forward RoIAlign over maps beyond one CTA's shared memory: tiled plane kernel against the generic kernel
TAG=${1:-ab}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi_align" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$TAG.log
for cfg in P2 BIG BIGD; do
  RLOD_FORCE_TILES=1 python tools/time_op.py align_fwd $cfg 30
  RLOD_NO_TILES=1 python tools/time_op.py align_fwd $cfg 30
done 2>&1 | tee gpurun_out/ab_tiles_$TAG.log
