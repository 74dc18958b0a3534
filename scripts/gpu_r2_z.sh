#!/bin/bash
# round 2, final evidence run on ONE B200: suite, driver-style bench lines, launch list, ncu summaries of the kernels
# the default paths launch, per-step SIR times, compute-sanitizer logs.  Summaries only (gpurun_out merges <= 64 MiB).
set -x
O=gpurun_out/r2z
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) 2>$O/bench_default.err | tail -1 > $O/bench_default.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 1 2>>$O/bench_default.err | tail -1 > $O/bench_reference.json
for w in schelling market economy walk sir ensemble; do
  timeout 600 python bench.py --workload $w --no-cpu 2>>$O/bench_full.err | tail -1 >> $O/bench_full.jsonl
done
timeout 600 python bench.py --workload schelling --grid 8192 --steps 300 --no-cpu --no-e2e 2>>$O/bench_full.err | tail -1 >> $O/bench_full.jsonl
timeout 600 python bench.py --workload schelling --grid 16384 --steps 100 --no-cpu --no-e2e 2>>$O/bench_full.err | tail -1 >> $O/bench_full.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-also > $O/ncu_bench.log 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > $O/sanitizer_$tool.log 2>&1; echo "rc=$?" >> $O/sanitizer_$tool.log
done
export JXB_NO_GRAPH=1
timeout 600 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:sir_ --csv --log-file $O/sir_per_launch.csv python scripts/prof_target.py sir 100 > $O/ncu_sir1.log 2>&1
bash scripts/ncu_cap.sh $O/schelling_bits_first20 schelling_bits 0 1 python scripts/prof_target.py schelling 20
bash scripts/ncu_cap.sh $O/sir_pull_s_step20 sir_pull_s 19 1 python scripts/prof_target.py sir 21
bash scripts/ncu_cap.sh $O/sir_push_step70 "sir_push|sir_transition" 138 2 python scripts/prof_target.py sir 72
bash scripts/ncu_cap.sh $O/economy_kernels "economy_step|gini_" 8 8 python scripts/prof_target.py economy 3
bash scripts/ncu_cap.sh $O/walk_step_kernel step_kernel 2 1 python scripts/prof_target.py walk 4
bash scripts/ncu_cap.sh $O/market_step_kernel step_kernel 2 1 python scripts/prof_target.py market 4
bash scripts/ncu_cap.sh $O/ensemble_kernel ensemble_kernel 0 1 python scripts/prof_target.py ensemble 200
du -sh $O; ls $O
