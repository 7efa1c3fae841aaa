#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_protocol.py tests/test_gpu_ipp_modes.py -m gpu -x -q 2>&1 | tail -3
for lg in 16 12 10; do
  python tools/prove_profile.py $lg 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lg',d['lg'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]],'verify',[round(x,2) for x in d['verify_ms_unprofiled'][:5]])"
done
BPG_TRACE=1 python tools/prove_profile.py 16 0 2>&1 >/dev/null | grep "verify " | tail -12
