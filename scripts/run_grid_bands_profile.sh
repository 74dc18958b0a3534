mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_grid_sharded.py -x -q -m gpu -k "single_gpu_band" 2>&1 | tail -6
run() { timeout 400 "$@" 2>gpurun_out/last_stderr.log | grep '^{' | tee -a gpurun_out/band_bench.jsonl | cut -c1-260; tail -3 gpurun_out/last_stderr.log | grep -i "error\|Traceback" ; }
run python bench.py --workload schelling --grid 16384 --steps 100 --no-cpu --no-e2e
JXB_GRID_BANDS=1 run python bench.py --workload schelling --grid 8192 --steps 300 --no-cpu --no-e2e
JXB_GRID_BANDS=1 run python bench.py --workload schelling --no-cpu --no-e2e
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file gpurun_out/r01_launches_grid_bands_8192.csv python scripts/profile_grid_bands.py 8192 40 > gpurun_out/prof1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grid_shard -s 8 -c 4 -o gpurun_out/grid_bands_8192_step3 -f python scripts/profile_grid_bands.py 8192 4 > gpurun_out/prof2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grid_shard_sweep -s 35 -c 1 -o gpurun_out/grid_bands_8192_sweep_step36 -f python scripts/profile_grid_bands.py 8192 40 > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof1.log gpurun_out/prof2.log gpurun_out/prof3.log
