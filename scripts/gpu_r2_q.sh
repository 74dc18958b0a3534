#!/bin/bash
# grid bands on N GPUs: per-kernel trace (first rank lines) + bench lines
N=$1
O=gpurun_out/r2q$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_grid_sharded.py -q -m gpu -k "two_gpus" > $O/pytest_two_gpus.log 2>&1; echo "rc=$?" >> $O/pytest_two_gpus.log
  tail -3 $O/pytest_two_gpus.log
fi
for g in 4096 16384; do
  JXB_GRID_BANDS=1 JXB_GS_TRACE=1 JXB_NO_GRAPH=1 timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --workload schelling --shard --grid $g --steps 20 --warmup 1 --no-cpu --no-e2e --no-also 2>&1 | grep "gs_trace" | cut -c1-700 > $O/trace_$g.txt
  echo "== $g, $N GPUs"; tail -$N $O/trace_$g.txt | sort | head -2
done
bash scripts/gpu_r2_k.sh $N
cp gpurun_out/r2k$N/bench.jsonl $O/bench.jsonl
