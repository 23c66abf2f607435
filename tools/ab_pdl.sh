#!/bin/bash
# A/B of programmatic dependent launch in the backward RoIAlign and RoIPool forward chains (whole call, cold L2)
TAG=${1:-ab}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for spec in "align_bwd C2" "align_bwd C4" "pool_fwd C2" "align_fwd C2"; do
  python tools/time_op.py $spec 60
  RLOD_NO_PDL=1 python tools/time_op.py $spec 60
done 2>&1 | tee gpurun_out/ab_pdl_$TAG.log
