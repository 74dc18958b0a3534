#!/bin/bash
# grid bands (requests -> slot owner -> cell owner): parity on one GPU + per-kernel trace + bench lines
O=gpurun_out/r2p
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_grid_sharded.py -x -q -m gpu > $O/pytest_grid.log 2>&1; echo "rc=$?" >> $O/pytest_grid.log
tail -4 $O/pytest_grid.log
for g in 4096 16384; do
  JXB_GRID_BANDS=1 JXB_GS_TRACE=1 JXB_NO_GRAPH=1 timeout 600 python bench.py --workload schelling --grid $g --steps 20 --warmup 1 --no-cpu --no-e2e --no-also 2>&1 | grep "gs_trace" | tail -1 | cut -c1-700
done
bash scripts/gpu_r2_k.sh 1
