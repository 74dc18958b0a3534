#!/bin/bash
# grid bands v2 (partitioned slots, records to the target rows' owner): parity on one GPU (1, 2, 3 ranks) + the
# one-GPU band-mode bench lines
O=gpurun_out/r2j
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_grid_sharded.py -x -q -m gpu > $O/pytest_grid.log 2>&1; echo "rc=$?" >> $O/pytest_grid.log
tail -30 $O/pytest_grid.log
for g in 4096 8192 16384; do
  JXB_GRID_BANDS=1 timeout 600 python bench.py --workload schelling --grid $g --steps 20 --warmup 5 --no-cpu --no-e2e --no-also 2>$O/b$g.err | tail -1 >> $O/bench.jsonl
  tail -2 $O/b$g.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2j/bench.jsonl'):
    try:
        d = json.loads(l); print(d['config'].get('workload'), d['ms_per_step'], d['value'], d.get('gpu_launches'))
    except Exception as e: print('bad line', l[:200])
PY
