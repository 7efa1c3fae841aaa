"""Scratch timing of the MSM pipeline (development aid, not the bench)."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc_bulletproof_b200 import Context, Table
from mpc_bulletproof_b200.api import dev_sum_encode
from tests.util import *

ctx = Context(0)
r = rng(3)
m = 4096
pb = points_bytes([rand_point(r) for _ in range(m)])
for lg in (12, 16, 18, 20, 22):
    n = 1 << lg
    tab = Table(ctx, pb * (n // m))
    sc = torch.randint(-2**31, 2**31, (n, 8), dtype=torch.int64, device="cuda").to(torch.int32)
    sc[:, 7] &= 0x0FFFFFFF
    out = torch.zeros(32, dtype=torch.int32, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for c in [0, -1]:
        if c == -1:
            tab.set_windows(0)
        for _ in range(2):
            tab.dev_msm(sc.data_ptr(), 1, out.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tab.dev_msm(sc.data_ptr(), 1, out.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ctx.profile(True); ctx.profile_reset()
        for _ in range(3):
            tab.dev_msm(sc.data_ptr(), 1, out.data_ptr())
        pr = ctx.profile_read(); ctx.profile(False)
        br = " ".join(f"{k}={v[0]/max(v[1],1):.3f}" for k, v in pr.items() if v[1])
        print(f"n=2^{lg} c={c} {ms:.3f} ms  {n/ms/1e3:.1f} Mpoints/s | {br}", flush=True)
    tab.close()
