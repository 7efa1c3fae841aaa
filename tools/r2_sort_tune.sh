run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-r1cs --no-varbase --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), d['result_ok'], {k:round(v,4) for k,v in d['phases_ms'].items() if k in ('hist','scan','scatter')})"; }
run BPG_SORT=atomic
run BPG_RS_FIN_THREADS=256
run BPG_RS_FIN_THREADS=512
run BPG_RS_PARTS=2048 BPG_RS_FIN_CAP=8192 BPG_RS_FIN_THREADS=256
run BPG_RS_PARTS=2048 BPG_RS_FIN_CAP=8192 BPG_RS_FIN_THREADS=512
run BPG_RS_PARTS=4096 BPG_RS_FIN_CAP=4096 BPG_RS_FIN_THREADS=256
run BPG_RS_PARTS=512 BPG_RS_FIN_CAP=32768 BPG_RS_FIN_THREADS=512
