#!/bin/bash
# N-GPU measurement pass of a round (run under gpurun --gpus N): bench, batch verification, N-rank checks.
tag=${1:-r}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
timeout 300 $TR --master-port 29502 tools/multi_gpu_check.py > gpurun_out/${tag}_multi_gpu_check_n$N.json 2> gpurun_out/${tag}_multi_gpu_check_n$N.err; echo "check rc=$?"
if [ "$N" = "8" ]; then
  timeout 400 $TR --master-port 29503 tools/batch_verify_bench.py 1024 16 4 > gpurun_out/${tag}_batch_verify_1024_n$N.json 2> gpurun_out/${tag}_batch_verify_1024_n$N.err; echo "batch rc=$?"
fi
timeout 400 $TR --master-port 29504 tools/sweep_multi.py 24 > gpurun_out/${tag}_sweep_strong_n$N.json 2> gpurun_out/${tag}_sweep_strong_n$N.err; echo "strong rc=$?"
timeout 200 $TR --master-port 29505 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-r1cs > gpurun_out/${tag}_bench_ref_n$N.json 2>/dev/null; echo "ref rc=$?"
tail -c 400 gpurun_out/${tag}_bench_n$N.json
