#!/bin/bash
# final check of the tree: the whole GPU suite, smoke(), the default bench line
O=gpurun_out/r2final
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 2>$O/bench.err | tail -1 > $O/bench_default.json
python -c "
import json; d=json.load(open('$O/bench_default.json'))
print('value %.3e' % d['value'], 'ms/step', d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'launches', d['gpu_launches'], 'also', [(a['config']['workload'], round(a['ms_per_step'],4)) for a in d.get('also', [])])
"
