#!/bin/bash
set -x
O=gpurun_out/r2e
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for w in sir economy; do
  timeout 600 python bench.py --workload $w --no-cpu --no-e2e 2>>$O/bench_full.err | tail -1 >> $O/bench_full.jsonl
done
export JXB_NO_GRAPH=1
for mode in auto push pull_s; do
  JXB_SIR_MODE=$mode timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:sir_ --csv --log-file $O/sir_launches_$mode.csv python scripts/prof_target.py sir 100 > $O/ncu_sir_$mode.log 2>&1
done
bash scripts/ncu_cap.sh $O/economy_step2 "economy_step|gini_accumulate" 5 3 python scripts/prof_target.py economy 3
du -sh $O; ls -la $O
