#!/bin/bash
# window width of the generator table against prove/verify time (R1CS, 2^16 and 2^12 multipliers)
for lg in 16 12; do
  for c in 11 12 13 14 15 16 17; do
    python tools/prove_profile.py $lg 0 $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lg',d['lg'],'c',d['table_c'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]],'verify',[round(x,2) for x in d['verify_ms_unprofiled'][:3]])"
  done
done | tee gpurun_out/r2q_tablec.txt
