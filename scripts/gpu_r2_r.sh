#!/bin/bash
# grid bands on N GPUs, bench lines only: bash scripts/gpu_r2_r.sh <N>
N=$1
O=gpurun_out/r2r$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; JXB_GRID_BANDS=1 timeout 600 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>$O/$name.err | tail -1 >> $O/bench.jsonl; grep -i "error\|Traceback" $O/$name.err | head -3; }
run sch4096 --workload schelling --shard --grid 4096 --steps 20 --no-cpu --no-e2e --no-also
run sch8192a --workload schelling --shard --grid 8192 --steps 20 --no-cpu --no-e2e --no-also
run sch8192 --workload schelling --shard --grid 8192 --steps 300 --no-cpu --no-e2e --no-also
run sch16384a --workload schelling --shard --grid 16384 --steps 20 --no-cpu --no-e2e --no-also
run sch16384 --workload schelling --shard --grid 16384 --steps 100 --no-cpu --no-e2e --no-also
python - <<PY
import json
for l in open('$O/bench.jsonl'):
    try:
        d = json.loads(l); print(d['n_gpus'], d['config'].get('workload'), 'K', d['steps'], 'us/step %.1f' % (d['ms_per_step']*1000), '%.3e' % d['value'])
    except Exception as e: print('bad line', l[:200])
PY
