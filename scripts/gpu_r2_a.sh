#!/bin/bash
# round 2, GPU call A: full GPU test suite, every workload's bench line, launch list, sanitizer logs
set -x
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for w in schelling market economy walk sir ensemble; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 2>$O/bench_$w.err | tail -1 >> $O/bench20.jsonl
done
timeout 600 python bench.py --workload schelling --no-cpu 2>>$O/bench_schelling.err | tail -1 >> $O/bench_full.jsonl
timeout 600 python bench.py --workload sir --no-cpu --no-e2e 2>>$O/bench_sir.err | tail -1 >> $O/bench_full.jsonl
timeout 600 python bench.py --workload economy --no-cpu --no-e2e 2>>$O/bench_economy.err | tail -1 >> $O/bench_full.jsonl
timeout 600 python bench.py --workload ensemble --no-cpu 2>>$O/bench_ensemble.err | tail -1 >> $O/bench_full.jsonl
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1; echo "rc=$?" >> $O/sanitizer_$tool.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu > $O/ncu_bench.log 2>&1
ls -la $O
