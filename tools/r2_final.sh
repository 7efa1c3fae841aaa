#!/bin/bash
# One-GPU final pass of round 2: the whole GPU suite, the bench line, the reference arm, per-round shares of a proof,
# the config-2/4 sweep.  Outputs under gpurun_out/r2z_*.
tag=r2z
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
tools/r2_round_shares.sh
timeout 900 python tools/sweep.py ipp,r1cs,rand 24 > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
