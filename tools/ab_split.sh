#!/bin/bash
# the split of the pooling launch's last wave at the shard sizes of 8 / 4 / 2 GPUs (3 / 6 / 12 images per rank):
# RLOD_SPLIT=S forces S parts per item of the last wave, RLOD_SPLIT_SLOTS the number of resident CTAs a wave is
for n in 3 6 12; do
  for slots in 148 296; do
    for S in 1 2 3; do RLOD_SPLIT_SLOTS=$slots RLOD_SPLIT=$S python tools/time_merge.py C4x$n | sed "s/^/slots=$slots S=$S /"; done
  done
done 2>&1 | cut -c1-175 | tee gpurun_out/ab_split2.log
