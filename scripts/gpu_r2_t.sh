#!/bin/bash
# persistent Schelling kernel with lane-sequential Feistel walks: parity subset + bench lines
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "schelling or Schelling or c2 or C2" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
for a in "--steps 20 --warmup 5" "--steps 20 --warmup 5" ""; do
  python bench.py --workload schelling $a --no-cpu --no-e2e --no-also 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K', d['steps'], 'total ms %.4f' % (d['ms_per_step']*d['steps']), 'value %.4e' % d['value'])" | tee -a $O/schelling.txt
done
python bench.py --workload schelling --grid 8192 --steps 300 --no-cpu --no-e2e --no-also 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('8192 K', d['steps'], 'us/step %.2f' % (d['ms_per_step']*1000))" | tee -a $O/schelling.txt
