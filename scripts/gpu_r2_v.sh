#!/bin/bash
# ncu --set full of the shipped band kernels at step 1 (4096^2, one band)
O=gpurun_out/r2v
mkdir -p $O
for K in grid_shard_moveout grid_shard_forward grid_shard_apply; do
  bash scripts/ncu_cap.sh $O/${K}_step1 $K 0 1 python scripts/profile_grid_bands.py 4096 1 > /dev/null 2>&1
  head -24 $O/${K}_step1.summary.txt | grep "duration\|dram_bytes\|warp_instructions\|issue_slot\|top_stalls\|occupancy_pct\|registers"
done
