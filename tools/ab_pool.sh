#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pool" > gpurun_out/pytest_pool.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pool.log
python tools/time_op.py pool_fwd C2 40
python tools/time_op.py pool_fwd C4 20
