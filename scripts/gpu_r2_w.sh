#!/bin/bash
O=gpurun_out/r2w
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_grid_sharded.py -x -q -m gpu -k "share_one_gpu or single_gpu_band" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for g in 4096 16384; do
JXB_GRID_BANDS=1 timeout 100 python bench.py --workload schelling --grid $g --steps 20 --warmup 5 --no-cpu --no-e2e --no-also 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$g us/step %.1f' % (d['ms_per_step']*1000))"
done
