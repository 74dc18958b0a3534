# 2-GPU check of the row-band sharded Schelling grid: parity tests, then strong-scaling bench lines
# (gpurun --gpus 2 -- bash scripts/run_grid_shard_2gpu.sh [bench-only])
mkdir -p gpurun_out
if [ "$1" != "bench-only" ]; then
  timeout 600 python -m pytest tests/test_gpu_grid_sharded.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -15
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600"
run() { timeout 400 "$@" 2>gpurun_out/last_stderr.log | grep '^{' | tee -a gpurun_out/grid_shard_bench.jsonl | cut -c1-420; tail -3 gpurun_out/last_stderr.log | grep -i "error\|Traceback" ; }
run $TR bench.py --gpus 2 --workload schelling --shard --no-cpu
run $TR bench.py --gpus 2 --workload schelling --shard --grid 8192 --steps 300 --no-cpu
run $TR bench.py --gpus 2 --workload schelling --shard --grid 16384 --steps 100 --no-cpu
run python bench.py --workload schelling --grid 16384 --steps 100 --no-cpu --no-e2e
