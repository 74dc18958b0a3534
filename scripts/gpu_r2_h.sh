#!/bin/bash
# Schelling mover-loop experiments: entries per thread (JXB_SCH_MV) x resident CTAs per SM (JXB_SCH_MINB)
O=gpurun_out/r2h
mkdir -p $O
for cfg in "4 2" "8 2" "4 3" "8 3" "2 2"; do
  set -- $cfg
  JXB_NVCC_EXTRA="-DJXB_SCH_MV=$1 -DJXB_SCH_MINB=$2" python __graft_entry__.py --force > /dev/null 2>&1
  for a in "--steps 20 --warmup 5" ""; do
    python bench.py --workload schelling $a --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('MV=$1 MINB=$2 K', d['steps'], 'total ms %.4f' % (d['ms_per_step']*d['steps']), 'value %.4e' % d['value'])" | tee -a $O/schelling_mv.txt
  done
done
python __graft_entry__.py --force > /dev/null 2>&1
