#!/bin/bash
python -m pytest tests/test_gpu_protocol.py tests/test_gpu_msm.py tests/test_gpu_mpc.py -m gpu -x -q 2>&1 | tail -3
pp() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 lg',d['lg'],'verify',[round(x,3) for x in d['verify_ms_unprofiled'][:3]])"; }
for lg in 8 10 12 13 14 15; do
  BPG_MIXED_COMB_MAX=0 python tools/prove_profile.py $lg 0 2>/dev/null | pp bucket
  BPG_MIXED_COMB_MAX=10000000 python tools/prove_profile.py $lg 0 2>/dev/null | pp comb
done
