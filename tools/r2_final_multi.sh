#!/bin/bash
# N-GPU final pass of round 2 (run under gpurun --gpus N): bench, N-rank checks, batch verification at N = 8.
N=${1:-8}; tag=r2z
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
timeout 300 $TR --master-port 29502 tools/multi_gpu_check.py > gpurun_out/${tag}_multi_gpu_check_n$N.json 2> gpurun_out/${tag}_multi_gpu_check_n$N.err; echo "check rc=$?"
if [ "$N" = "8" ]; then
  timeout 400 $TR --master-port 29503 tools/batch_verify_bench.py 1024 16 4 > gpurun_out/${tag}_batch_verify_1024_n$N.json 2> gpurun_out/${tag}_batch_verify_1024_n$N.err; echo "batch rc=$?"
fi
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
fi
tail -c 300 gpurun_out/${tag}_bench_n$N.json
