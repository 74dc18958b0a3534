#!/bin/bash
set -x
O=gpurun_out/r2g
mkdir -p $O
JXB_SIR_PULL=lb timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_net_sharded.py tests/test_gpu_record.py -q -m gpu -k "sir or network" > $O/pytest_lb.log 2>&1; echo "rc=$?" >> $O/pytest_lb.log
tail -3 $O/pytest_lb.log
timeout 600 python bench.py --workload sir --no-cpu --no-e2e 2>>$O/bench.err | tail -1 >> $O/bench.jsonl
JXB_SIR_PULL=lb timeout 600 python bench.py --workload sir --no-cpu --no-e2e 2>>$O/bench.err | tail -1 >> $O/bench.jsonl
JXB_SIR_PULL=lb JXB_SIR_MODE=pull_s timeout 600 python bench.py --workload sir --no-cpu --no-e2e 2>>$O/bench.err | tail -1 >> $O/bench.jsonl
export JXB_NO_GRAPH=1
for mode in auto pull_s; do
  JXB_SIR_PULL=lb JXB_SIR_MODE=$mode timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:sir_ --csv --log-file $O/sir_launches_lb_$mode.csv python scripts/prof_target.py sir 100 > $O/ncu_sir_$mode.log 2>&1
done
JXB_SIR_MODE=push timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:sir_ --csv --log-file $O/sir_launches_push.csv python scripts/prof_target.py sir 100 > $O/ncu_sir_push.log 2>&1
JXB_SIR_PULL=lb JXB_SIR_MODE=pull_s bash scripts/ncu_cap.sh $O/sir_pull_lb_step20 "sir_pull_s" 20 1 python scripts/prof_target.py sir 22
du -sh $O
