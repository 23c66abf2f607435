#!/bin/bash
# Ablation timings of k_align8_fwd_walk on the GPU box: rebuild with -DRLOD_ABL=n (results are WRONG
# for n != 0; only the kernel time matters) and print the kernel's mean launch time.
for n in 0 1 2 3 4; do
  touch rlobjectdetection_b200/csrc/roi_align.cu
  make -C rlobjectdetection_b200/csrc EXTRA="-DRLOD_ABL=$n" > /dev/null 2>&1
  python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ABL=$n', 'align_fwd us/launch', d['roofline']['avg_launch_us'], 'step ms', d['ms_per_step'])"
done
touch rlobjectdetection_b200/csrc/roi_align.cu; make -C rlobjectdetection_b200/csrc > /dev/null 2>&1
