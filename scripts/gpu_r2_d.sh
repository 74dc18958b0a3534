#!/bin/bash
# round 2, GPU call D: suite + the default bench line (with `also`) + SIR per-step times per direction
set -x
O=gpurun_out/r2d
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) 2>$O/bench_default.err | tail -1 > $O/bench_default.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 1 2>>$O/bench_default.err | tail -1 > $O/bench_reference.json
for w in schelling sir economy; do
  timeout 600 python bench.py --workload $w --no-cpu --no-e2e 2>>$O/bench_full.err | tail -1 >> $O/bench_full.jsonl
done
export JXB_NO_GRAPH=1
for mode in auto push pull_s; do
  JXB_SIR_MODE=$mode timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum -k regex:sir_ --csv --log-file $O/sir_launches_$mode.csv python scripts/prof_target.py sir 100 > $O/ncu_sir_$mode.log 2>&1
done
du -sh $O; ls -la $O
