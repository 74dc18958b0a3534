#!/bin/bash
# round 2, GPU call C: full GPU suite on the new Schelling mover walk + f4, bench, racecheck rerun, ncu summaries
set -x
O=gpurun_out/r2c
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --workload schelling --steps 20 --warmup 5 --no-cpu 2>$O/bench_s20.err | tail -1 >> $O/bench.jsonl
timeout 600 python bench.py --workload schelling --no-cpu --no-e2e 2>$O/bench_s1000.err | tail -1 >> $O/bench.jsonl
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py ensemble market schelling_bits > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py schelling_bits ensemble > $O/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck.log
export JXB_NO_GRAPH=1
bash scripts/ncu_cap.sh $O/schelling_bits_first20 schelling_bits 0 1 python scripts/prof_target.py schelling 20
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:sir_ --csv --log-file $O/sir_per_launch.csv python scripts/prof_target.py sir 100 > $O/ncu_sir1.log 2>&1
bash scripts/ncu_cap.sh $O/sir_pull_s_step13 sir_pull_s 12 1 python scripts/prof_target.py sir 14
bash scripts/ncu_cap.sh $O/sir_push_step3 "sir_push|sir_transition" 4 2 python scripts/prof_target.py sir 4
bash scripts/ncu_cap.sh $O/economy_step2 "economy_step|gini_" 7 7 python scripts/prof_target.py economy 3
bash scripts/ncu_cap.sh $O/walk_step_kernel step_kernel 2 1 python scripts/prof_target.py walk 4
bash scripts/ncu_cap.sh $O/market_step_kernel step_kernel 2 1 python scripts/prof_target.py market 4
bash scripts/ncu_cap.sh $O/ensemble_kernel ensemble_kernel 0 1 python scripts/prof_target.py ensemble 200
du -sh $O; ls -la $O
