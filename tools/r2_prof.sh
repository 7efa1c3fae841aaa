for lg in 10 16; do
  python tools/prove_profile.py $lg 1 > gpurun_out/r2n_prove_$lg.json 2>gpurun_out/r2n_prove_$lg.err
done
