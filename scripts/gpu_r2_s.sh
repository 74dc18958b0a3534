#!/bin/bash
# band kernels: record / upload tests in band mode + compute-sanitizer on a small band-mode run
O=gpurun_out/r2s
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_record.py -x -q -m gpu -k "schelling" > $O/pytest_record.log 2>&1; echo "rc=$?" >> $O/pytest_record.log
tail -15 $O/pytest_record.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py schelling_bands > $O/sanitizer_${tool}_bands.log 2>&1; echo "rc=$?" >> $O/sanitizer_${tool}_bands.log
  tail -6 $O/sanitizer_${tool}_bands.log
done
