#!/bin/bash
# grid bands v2: per-kernel device time per rank (JXB_GS_TRACE) on N GPUs and on one
N=$1
O=gpurun_out/r2o$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for g in 4096 16384; do
  JXB_GRID_BANDS=1 JXB_GS_TRACE=1 JXB_NO_GRAPH=1 timeout 600 python bench.py --workload schelling --grid $g --steps 20 --warmup 1 --no-cpu --no-e2e --no-also 2>&1 | grep "gs_trace\|^{" | cut -c1-600 > $O/trace_1gpu_$g.txt
  JXB_GRID_BANDS=1 JXB_GS_TRACE=1 JXB_NO_GRAPH=1 timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --workload schelling --shard --grid $g --steps 20 --warmup 1 --no-cpu --no-e2e --no-also 2>&1 | grep "gs_trace\|^{" | cut -c1-600 > $O/trace_${N}gpu_$g.txt
  echo "== $g, 1 GPU"; grep gs_trace $O/trace_1gpu_$g.txt | tail -1
  echo "== $g, $N GPUs"; grep gs_trace $O/trace_${N}gpu_$g.txt | tail -$N
done
