#!/bin/bash
# grid bands v2: ncu --set full of the moveout and apply kernels at step 1 (4096^2, one band)
O=gpurun_out/r2m
mkdir -p $O
bash scripts/ncu_cap.sh $O/moveout_step1 grid_shard_moveout 0 1 python scripts/profile_grid_bands.py 4096 2
bash scripts/ncu_cap.sh $O/apply_step1 grid_shard_apply 0 1 python scripts/profile_grid_bands.py 4096 2
cat $O/moveout_step1.summary.txt | head -40
head -30 $O/moveout_step1.hotspots.txt
