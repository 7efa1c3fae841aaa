#!/bin/bash
python -m pytest tests/test_gpu_protocol.py tests/test_gpu_mpc.py -m gpu -x -q 2>&1 | tail -3
pp() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 lg',d['lg'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]],'verify',[round(x,2) for x in d['verify_ms_unprofiled'][:3]])"; }
for lg in 8 10 12 13 14; do
  BPG_COMMIT_COMB_MAX=0 python tools/prove_profile.py $lg 0 2>/dev/null | pp bucket
  BPG_COMMIT_COMB_MAX=1000000 python tools/prove_profile.py $lg 0 2>/dev/null | pp comb
done
python tools/prove_profile.py 16 0 2>/dev/null | pp default
