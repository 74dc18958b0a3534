#!/bin/bash
# grid bands: ncu --set full of one band kernel at step 1 (4096^2, one band): bash scripts/gpu_r2_m.sh <kernel-regex>
O=gpurun_out/r2m
mkdir -p $O
K=${1:-grid_shard_forward}
bash scripts/ncu_cap.sh $O/${K}_step1 $K 0 1 python scripts/profile_grid_bands.py 4096 2
cat $O/${K}_step1.summary.txt | head -40
head -40 $O/${K}_step1.hotspots.txt
