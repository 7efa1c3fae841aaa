#!/bin/bash
pp() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 lg',d['lg'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]])"; }
for lg in 12 13 14 16; do
  for dm in 2048 4096 8192; do
    for m0 in 1024 2048; do
      BPG_IPP_DIRECT_MAX=$dm BPG_IPP_M0=$m0 python tools/prove_profile.py $lg 0 2>/dev/null | pp "direct$dm-m0_$m0"
    done
  done
done
