#!/bin/bash
# SIR pull kernel: grid size in CTAs per SM (JXB_SIR_PULL_CPS); the kernel holds 6 resident CTAs per SM
O=gpurun_out/r2i
mkdir -p $O
for cps in 8 6 5 4 3 6 8; do
  JXB_SIR_PULL_CPS=$cps python bench.py --workload sir --steps 100 --warmup 5 --no-cpu --no-e2e --no-also 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CPS=$cps us/step %.2f' % (d['ms_per_step']*1000), 'frac', d['roofline']['frac'])" | tee -a $O/sir_cps.txt
done
