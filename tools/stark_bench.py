"""Stark-curve MSM timing on one B200: host-buffer call `bpg_stark_msm_table`
over a resident table, per-phase device times.  Points are i*G (sequential additions, a few
thousand distinct ones tiled to size); scalars uniform below the group order.  Prints JSON."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ctypes  # noqa: E402

from mpc_bulletproof_b200 import Context  # noqa: E402
from mpc_bulletproof_b200._lib import check, lib  # noqa: E402
from mpc_bulletproof_b200.stark import StarkTable  # noqa: E402

P = 2**251 + 17 * 2**192 + 1
N = 0x0800000000000010FFFFFFFFFFFFFFFFB781126DCAE7B2321E66A241ADC64D2F
GX = 0x01EF15C18599971B7BECED415A40F0C7DEACFD9B0D1819E03D723D8BC943CFCA
GY = 0x005668060AA49730B7BE4801DF46EC62DE53ECD11ABE43A32873000C36E8DC1F


def add(p, q):
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        lam = (3 * x1 * x1 + 1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def main():
    lgs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12,16,18,20").split(",")]
    windowed = not (len(sys.argv) > 2 and sys.argv[2] == "plain")
    ctx = Context(0)
    base, pts, cur = 4096, [], (GX, GY)
    for _ in range(base):
        pts.append(cur)
        cur = add(cur, (GX, GY))
    tile = b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts)
    out = {"rows": []}
    for lg in lgs:
        n = 1 << lg
        pb = tile * (n // base) if n >= base else tile[: 64 * n]
        t = StarkTable(ctx, pb)
        if windowed:
            t.set_windows(0)
        g = np.random.Generator(np.random.PCG64(lg))
        a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 31] &= 0x07  # < 2^251 < group order
        sc_t = torch.from_numpy(a).pin_memory()  # scalars start in page-locked host memory
        res = ctypes.create_string_buffer(64)

        def run():
            check(lib().bpg_stark_msm_table(ctx._h, t._h, 0, n, ctypes.c_void_p(sc_t.data_ptr()), 1, res))
            return res.raw

        for _ in range(2):
            r0 = run()
        ctx.profile(True)
        ctx.profile_reset()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            r = run()
        dt = (time.perf_counter() - t0) / reps * 1e3
        prof = ctx.profile_read()
        ctx.profile(False)
        assert r == r0
        out["rows"].append({"lg_n": lg, "window": t.window, "e2e_ms": round(dt, 3), "mpoints_s": round(n / dt / 1e3, 1),
                            "phases_ms": {k: round(v[0] / reps, 3) for k, v in prof.items() if v[1]}, "result_x": r[:32].hex()})
        print(out["rows"][-1], file=sys.stderr, flush=True)
        t.close()
    # InnerProductProof::create rounds over the Stark curve (transcript stand-in: SHA-256 of L || R)
    import hashlib

    from mpc_bulletproof_b200.stark import StarkIpp

    out["ipp"] = []
    for lg in (10, 14, 16):
        n = 1 << lg
        pb = tile * (2 * n // base) if 2 * n >= base else tile[: 64 * 2 * n]
        tG = StarkTable(ctx, pb[: 64 * n])
        tH = StarkTable(ctx, pb[64 * n : 128 * n])
        g = np.random.Generator(np.random.PCG64(100 + lg))
        vecs = []
        for _ in range(4):
            a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
            a[:, 31] &= 0x07
            vecs.append(a.tobytes())
        Q = tile[64 * 7 : 64 * 8]
        best = 1e9
        for it in range(3):
            t0 = time.perf_counter()
            st = StarkIpp(ctx, tG, tH, Q, vecs[2], vecs[3], vecs[0], vecs[1])
            j = 0
            while st.rounds_left():
                L, R = st.round_lr()
                u = int.from_bytes(hashlib.sha256(bytes([j]) + L + R).digest(), "little") % N
                st.round_fold(u.to_bytes(32, "little"), pow(u, -1, N).to_bytes(32, "little"))
                j += 1
            fa, fb = st.finish()
            best = min(best, (time.perf_counter() - t0) * 1e3)
            st.close()
        out["ipp"].append({"lg_n": lg, "create_ms": round(best, 3), "rounds": lg})
        print(out["ipp"][-1], file=sys.stderr, flush=True)
        tG.close()
        tH.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
