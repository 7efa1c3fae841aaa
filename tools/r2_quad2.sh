#!/bin/bash
for q in 0 1; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__cycles_active.avg --clock-control none --csv --log-file gpurun_out/r2q_launches_q$q.csv \
    -k regex:"k_comb_round|k_comb_final|k_ipp_fold_cross" -c 60 env BPG_COMB_QUAD=$q python tools/prove_profile.py 12 0 > /dev/null 2> gpurun_out/r2q_ncu_$q.err
done
