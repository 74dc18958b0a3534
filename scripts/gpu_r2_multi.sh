#!/bin/bash
# round 2, multi-GPU measurement: bash scripts/gpu_r2_multi.sh <N>   (under gpurun --gpus N)
set -x
N=$1
O=gpurun_out/r2m$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
# the sharded GPU tests that need >= 2 real GPUs (skipped on the 1-GPU boxes)
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_grid_sharded.py tests/test_gpu_net_sharded.py -q -m gpu -k "two_gpus" > $O/pytest_two_gpus.log 2>&1; echo "rc=$?" >> $O/pytest_two_gpus.log
  tail -3 $O/pytest_two_gpus.log
fi
run() { name=$1; shift; timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>$O/$name.err | tail -1 >> $O/bench.jsonl; tail -2 $O/$name.err; }
run default --steps 20 --warmup 5 --no-cpu
run market --workload market --shard --no-cpu --no-e2e
run economy --workload economy --shard --no-cpu --no-e2e
run sir --workload sir --shard --no-cpu --no-e2e
run sch8192 --workload schelling --shard --grid 8192 --steps 300 --no-cpu --no-e2e
run sch16384 --workload schelling --shard --grid 16384 --steps 100 --no-cpu --no-e2e
ls -la $O; wc -l $O/bench.jsonl
