#!/bin/bash
# 2-GPU check: NCCL parity test (gathered == single rank) and the sharded bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_s3.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_s3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 --no-ops > gpurun_out/bench_n2_s3.log 2>&1; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2_s3.log | cut -c1-400
RLOD_NO_PDL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 50 --warmup 5 --no-ops --no-cpu-baseline > gpurun_out/bench_n2_s3_nopdl.log 2>&1; echo "bench n2 nopdl rc=$?"; tail -1 gpurun_out/bench_n2_s3_nopdl.log | cut -c1-200
