#!/bin/bash
# final 2-GPU check: the tests that need two GPUs, the default line under torchrun, the band lines
N=2
O=gpurun_out/r2u
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_grid_sharded.py tests/test_gpu_net_sharded.py -q -m gpu -k "two_gpus" > $O/pytest_two_gpus.log 2>&1; echo "rc=$?" >> $O/pytest_two_gpus.log
tail -3 $O/pytest_two_gpus.log
timeout 600 $TR --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu 2>$O/default.err | tail -1 > $O/bench_default_2gpu.json
python -c "
import json; d=json.load(open('$O/bench_default_2gpu.json'))
print('value %.3e' % d['value'], 'sharded', [(a['config']['workload'], a.get('scaling'), round(a['ms_per_step'],4), '%.3e' % a['value']) for a in d.get('sharded', [])])
"
bash scripts/gpu_r2_r.sh 2
