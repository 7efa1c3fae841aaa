#!/bin/bash
# One-GPU measurement pass of a round (run under gpurun): tests, bench, ncu launch list, one full
# capture of the top kernels, the config-2/3/4 sweep.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-r1cs --no-varbase > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_accum$|k_rs_scatter|k_rs_finish|k_rs_hist|k_reduce_leaf_thread|k_sum_encode" -s 20 -c 5 \
  -o gpurun_out/${tag}_top -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-r1cs --no-varbase > gpurun_out/${tag}_ncu_f.log 2>&1; echo "ncu full rc=$?"
timeout 900 python tools/sweep.py msm,varbase,ipp,r1cs,rand 24 > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"
timeout 120 python tools/prove_profile.py 16 2 > gpurun_out/${tag}_prove_prof.json 2> gpurun_out/${tag}_prove_prof.err
