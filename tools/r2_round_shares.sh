#!/bin/bash
# Per-round wall-clock (BPG_TRACE) and the ncu launch list of one R1CS proof at 2^16 and 2^10 multipliers.
python -m pytest tests/test_gpu_protocol.py -m gpu -x -q -k "verification_scalars" 2>&1 | tail -2
for lg in 16 10; do
  BPG_TRACE=1 python tools/prove_profile.py $lg 0 > gpurun_out/r2p_prove_$lg.json 2> gpurun_out/r2p_trace_$lg.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_$lg.csv \
    python tools/prove_profile.py $lg 0 > /dev/null 2> gpurun_out/r2p_ncu_$lg.err
done
tail -n 60 gpurun_out/r2p_trace_16.txt | head -80
