# 2-GPU check of the sharded SIR network: parity tests, then strong-scaling bench lines
# (gpurun --gpus 2 -- bash scripts/run_net_shard_2gpu.sh)
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_net_sharded.py -x -q -m gpu 2>&1 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600"
run() { timeout 400 "$@" 2>gpurun_out/last_stderr.log | grep '^{' | tee -a gpurun_out/net_shard_bench.jsonl | cut -c1-330; tail -3 gpurun_out/last_stderr.log | grep -i "error\|Traceback" ; }
run $TR bench.py --gpus 2 --workload sir --shard --no-cpu
JXB_NET_SPLIT=nodes run $TR bench.py --gpus 2 --workload sir --shard --no-cpu
