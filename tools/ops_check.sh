#!/bin/bash
# the per-op table (bench.py --ops) with and without dependent launches, next to tools/time_op.py in the same call
mkdir -p gpurun_out
python bench.py --ops > gpurun_out/ops_a.json 2> gpurun_out/ops_a.err; echo "ops rc=$?"
RLOD_NO_PDL=1 python bench.py --ops > gpurun_out/ops_b.json 2> gpurun_out/ops_b.err; echo "ops nopdl rc=$?"
grep -h '"op"' gpurun_out/ops_a.err | cut -c1-110 | head -8
echo ---
grep -h '"op"' gpurun_out/ops_b.err | cut -c1-110 | head -8
for spec in "align_bwd C2" "pool_fwd C2"; do python tools/time_op.py $spec 10; python tools/time_op.py $spec 60; done
