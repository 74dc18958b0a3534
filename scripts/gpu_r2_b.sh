#!/bin/bash
# round 2, GPU call B: f4 tests, racecheck rerun, ncu captures of the kernels the default paths launch
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_record.py tests/test_gpu_parity.py -q -m gpu --durations=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py ensemble market > $O/sanitizer_racecheck_ensemble.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck_ensemble.log
export JXB_NO_GRAPH=1
NCU="ncu --clock-control none --import-source on"
# Schelling: the driver's window (first 20 steps from the seeded layout)
timeout 600 $NCU --set full -k regex:schelling_bits -c 1 -o $O/schelling_bits_first20 -f python scripts/prof_target.py schelling 20 > $O/ncu_schelling.log 2>&1
# SIR: per-launch time + DRAM bytes over the 100-step epidemic, then full sets of one launch of each kernel
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:sir_ --csv --log-file $O/sir_per_launch.csv python scripts/prof_target.py sir 100 > $O/ncu_sir1.log 2>&1
timeout 900 $NCU --set full -k regex:sir_pull_s -s 12 -c 1 -o $O/sir_pull_s_step13 -f python scripts/prof_target.py sir 14 > $O/ncu_sir2.log 2>&1
timeout 900 $NCU --set full -k regex:"sir_push|sir_transition" -s 4 -c 2 -o $O/sir_push_step3 -f python scripts/prof_target.py sir 4 > $O/ncu_sir3.log 2>&1
# economy: the split kernels + Gini
timeout 900 $NCU --set full -k regex:"economy_step|gini_" -s 7 -c 7 -o $O/economy_step2 -f python scripts/prof_target.py economy 3 > $O/ncu_economy.log 2>&1
# walker / market step_kernel, ensemble kernel
timeout 600 $NCU --set full -k regex:step_kernel -s 2 -c 1 -o $O/walk_step_kernel -f python scripts/prof_target.py walk 4 > $O/ncu_walk.log 2>&1
timeout 600 $NCU --set full -k regex:step_kernel -s 2 -c 1 -o $O/market_step_kernel -f python scripts/prof_target.py market 4 > $O/ncu_market.log 2>&1
timeout 900 $NCU --set full -k regex:ensemble_kernel -c 1 -o $O/ensemble_kernel -f python scripts/prof_target.py ensemble 200 > $O/ncu_ensemble.log 2>&1
ls -la $O
