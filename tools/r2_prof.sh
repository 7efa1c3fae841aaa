python -m pytest tests/test_gpu_ipp_modes.py -m gpu -x -q 2>&1 | tail -2
for lg in 16 18; do
  python tools/prove_profile.py $lg 1 > gpurun_out/r2s_prove_$lg.json 2>gpurun_out/r2s_prove_$lg.err
done
