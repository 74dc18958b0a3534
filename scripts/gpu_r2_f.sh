#!/bin/bash
set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_economy.py tests/test_gpu_sharded.py -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --workload economy --no-cpu --no-e2e 2>>$O/bench.err | tail -1 >> $O/bench.jsonl
timeout 600 python bench.py --workload economy --steps 4 --no-cpu --no-e2e 2>>$O/bench.err | tail -1 >> $O/bench.jsonl
export JXB_NO_GRAPH=1
JXB_SIR_MODE=pull_s bash scripts/ncu_cap.sh $O/sir_pull_s_step20 "sir_pull_s" 20 1 python scripts/prof_target.py sir 22
JXB_SIR_MODE=pull_s bash scripts/ncu_cap.sh $O/sir_pull_s_step70 "sir_pull_s" 70 1 python scripts/prof_target.py sir 72
JXB_SIR_MODE=push bash scripts/ncu_cap.sh $O/sir_push_step70 "sir_push|sir_transition" 140 2 python scripts/prof_target.py sir 72
bash scripts/ncu_cap.sh $O/economy_step2 "economy_step|gini_accumulate" 5 3 python scripts/prof_target.py economy 3
du -sh $O; ls -la $O
