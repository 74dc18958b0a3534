#!/bin/bash
# ncu_cap.sh <out-prefix> <kernel-regex> <skip> <count> <cmd...>: one `--set full` capture, summarised ON THE BOX
# (scripts/ncu_summary.py + the top stall sites); the .ncu-rep itself is deleted (gpurun_out/ merges <= 64 MiB).
out=$1; k=$2; skip=$3; cnt=$4; shift 4
timeout 900 ncu --clock-control none --import-source on --set full -k "regex:$k" -s $skip -c $cnt -o $out -f "$@" > $out.log 2>&1
if [ -f $out.ncu-rep ]; then
  python scripts/ncu_summary.py $out.ncu-rep > $out.summary.txt 2>&1
  python scripts/ncu_source_hotspots.py $out.ncu-rep "regex:$k" 24 > $out.hotspots.txt 2>&1
  rm -f $out.ncu-rep
fi
tail -3 $out.log
