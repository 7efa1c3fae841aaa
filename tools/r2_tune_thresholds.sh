#!/bin/bash
# The measured thresholds of round 2 (profiles/r2_ipp_strategy_tuning.json): commitments through the generator combs
# against the bucket method (BPG_COMMIT_COMB_MAX), the round strategy of the inner-product argument
# (BPG_IPP_DIRECT_MAX, BPG_IPP_M0), the verifier's table terms through the combs (BPG_MIXED_COMB_MAX), the window
# width of the generator table.  Run under gpurun on one B200; prints one line per setting.
pp() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 lg',d['lg'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]],'verify',[round(x,3) for x in d['verify_ms_unprofiled'][:3]])"; }
for lg in 8 10 12 13 14; do
  BPG_COMMIT_COMB_MAX=0 python tools/prove_profile.py $lg 0 2>/dev/null | pp commit-bucket
  BPG_COMMIT_COMB_MAX=1000000 python tools/prove_profile.py $lg 0 2>/dev/null | pp commit-comb
done
for lg in 12 13 14 16; do
  for dm in 2048 4096 8192; do
    for m0 in 1024 2048; do
      BPG_IPP_DIRECT_MAX=$dm BPG_IPP_M0=$m0 python tools/prove_profile.py $lg 0 2>/dev/null | pp "direct$dm-m0_$m0"
    done
  done
done
for lg in 8 10 12 13 14 15; do
  BPG_MIXED_COMB_MAX=0 python tools/prove_profile.py $lg 0 2>/dev/null | pp verify-bucket
  BPG_MIXED_COMB_MAX=10000000 python tools/prove_profile.py $lg 0 2>/dev/null | pp verify-comb
done
for lg in 16 12; do
  for c in 13 14 15 16 17; do python tools/prove_profile.py $lg 0 $c 2>/dev/null | pp "table_c$c"; done
done
