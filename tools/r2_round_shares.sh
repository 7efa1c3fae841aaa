#!/bin/bash
# Per-round wall-clock (BPG_TRACE) and the ncu launch list of one R1CS proof at 2^16 and 2^10 multipliers.
for lg in 16 10; do
  BPG_TRACE=1 python tools/prove_profile.py $lg 2 > gpurun_out/r2p_prove_$lg.json 2> gpurun_out/r2p_trace_$lg.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_$lg.csv \
    python tools/prove_profile.py $lg 0 > /dev/null 2> gpurun_out/r2p_ncu_$lg.err
done
for lg in 12 14 18; do python tools/prove_profile.py $lg 0 > gpurun_out/r2p_prove_$lg.json 2>/dev/null; done
