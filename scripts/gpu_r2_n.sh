#!/bin/bash
# grid bands v2 after the instruction diet: parity on one GPU, per-step launch list, bench lines
O=gpurun_out/r2n
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_grid_sharded.py -x -q -m gpu > $O/pytest_grid.log 2>&1; echo "rc=$?" >> $O/pytest_grid.log
tail -4 $O/pytest_grid.log
bash scripts/gpu_r2_l.sh 4096
cp gpurun_out/r2l/per_step_4096.txt $O/
bash scripts/gpu_r2_k.sh 1
