for leaf in "" t4 t8 t16 q4; do for c in 0 16 15; do
  BPG_LEAF=$leaf timeout 120 python tools/prove_profile.py 16 2 $c >> gpurun_out/r1i_tune.jsonl 2>> gpurun_out/r1i_tune.err
done; done
