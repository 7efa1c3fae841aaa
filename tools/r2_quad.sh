#!/bin/bash
python -m pytest tests/test_gpu_ipp_modes.py tests/test_gpu_protocol.py tests/test_gpu_mpc.py -m gpu -x -q 2>&1 | tail -3
pp() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 lg',d['lg'],'prove',[round(x,2) for x in d['prove_ms_unprofiled'][:3]],'verify',[round(x,2) for x in d['verify_ms_unprofiled'][:3]])"; }
for lg in 16 12 10; do
  BPG_COMB_QUAD=0 python tools/prove_profile.py $lg 0 2>/dev/null | pp old
  for q in 32 64 96 192 384; do
    BPG_COMB_QUADS_PER_SM=$q python tools/prove_profile.py $lg 0 2>/dev/null | pp quad$q
  done
done | tee gpurun_out/r2q_quad.txt
BPG_TRACE=1 python tools/prove_profile.py 16 0 2>&1 >/dev/null | grep "ipp" | tail -34 | head -20
