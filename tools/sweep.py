"""BASELINE.json configs 2-4 on one B200, through the C ABI with inputs serialised beforehand:
  msm   2^12..2^24 points: device-resident ms (CUDA events) and host-buffer ms (bpg_msm_table)
  ipp   InnerProductProof create / verify, n = 2^10..2^18 (reference benches/inner_product.rs)
  r1cs  prove / verify on the reference's bench circuit (benches/r1cs.rs), 2^10..2^18 multipliers
Usage: python tools/sweep.py [msm,ipp,r1cs] [max_lg_msm]   -> one JSON object."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context, Table  # noqa: E402
from mpc_bulletproof_b200 import protocol as P  # noqa: E402
from mpc_bulletproof_b200._lib import check, lib  # noqa: E402

BASE = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
DEV = torch.device("cuda", 0)


def dev_scalars(n, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    s = torch.randint(-(2**31), 2**31, (n, 8), dtype=torch.int64, device=DEV, generator=g).to(torch.int32)
    s[:, 7] &= 0x0FFFFFFF
    return s.contiguous()


def host_scalars(n, seed) -> bytes:
    g = np.random.Generator(np.random.PCG64(seed))
    a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F
    return a.tobytes()


def dev_points(ctx, comb, n, seed):
    out = torch.empty(n * 32, dtype=torch.uint8, device=DEV)
    k = dev_scalars(n, seed)
    torch.cuda.synchronize()
    comb.dev_mul(k.data_ptr(), n, out.data_ptr())
    ctx.sync()
    return out


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def sweep_msm(ctx, comb, max_lg):
    rows = []
    for lg in range(12, max_lg + 1, 2):
        n = 1 << lg
        pts = dev_points(ctx, comb, n, 100 + lg)
        t0 = time.perf_counter()
        table = Table(ctx, dev_ptr=pts.data_ptr(), n=n).set_windows(0)
        ctx.sync()
        t_table = (time.perf_counter() - t0) * 1e3
        sc = [dev_scalars(n, 200 + lg + i) for i in range(2)]
        out = torch.zeros(32, dtype=torch.int32, device=DEV)
        host = [s.cpu().pin_memory() for s in sc]
        res = ctypes.create_string_buffer(32)
        reps = 20 if lg <= 20 else 5
        for i in range(3):
            table.dev_msm(sc[i % 2].data_ptr(), 1, out.data_ptr())
        ctx.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        # the context has its own stream: bracket with host syncs
        t0 = time.perf_counter()
        for i in range(reps):
            table.dev_msm(sc[i % 2].data_ptr(), 1, out.data_ptr())
        ctx.sync()
        dev_ms = (time.perf_counter() - t0) * 1e3 / reps
        for i in range(2):
            check(lib().bpg_msm_table(ctx._h, table._h, 0, n, ctypes.c_void_p(host[i % 2].data_ptr()), 1, res))
        t0 = time.perf_counter()
        for i in range(reps):
            check(lib().bpg_msm_table(ctx._h, table._h, 0, n, ctypes.c_void_p(host[i % 2].data_ptr()), 1, res))
        e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
        rows.append({"lg_n": lg, "window": table.window, "dev_ms": round(dev_ms, 4), "dev_mpoints_s": round(n / dev_ms / 1e3, 1),
                     "e2e_ms": round(e2e_ms, 4), "e2e_mpoints_s": round(n / e2e_ms / 1e3, 1),
                     "gpoint_ops_s_eq": round(16 * n / dev_ms / 1e6, 2), "table_build_ms": round(t_table, 2)})
        print(rows[-1], file=sys.stderr, flush=True)
        table.close()
        del pts, sc, host
        torch.cuda.empty_cache()
    return rows


def sweep_ipp(ctx, comb, lgs):
    rows = []
    maxn = 1 << max(lgs)
    Gb = bytes(dev_points(ctx, comb, maxn, 1).cpu().numpy().tobytes())
    Hb = bytes(dev_points(ctx, comb, maxn, 2).cpu().numpy().tobytes())
    Q = bytes(dev_points(ctx, comb, 1, 3).cpu().numpy().tobytes())
    for lg in lgs:
        n = 1 << lg
        # generators resident as ONE windowed table [G | H] (what a BulletproofGens holds); Q is per call
        tG = tH = Table(ctx, Gb[: 32 * n] + Hb[: 32 * n]).set_windows(0)
        if 2 * n * 49152 <= 32e9:  # combs (48 KB per generator), as bpg_gens_new builds them
            tG.build_comb()
        h_off = n
        a, b, Gf, Hf = (host_scalars(n, 10 * lg + k) for k in range(4))
        cr, proof = [], None
        for it in range(5):
            t0 = time.perf_counter()
            proof = P.InnerProductProof.create(ctx, P.Transcript(b"bench"), Q, Gf, Hf, tG, tH, a, b, h_off=h_off)
            cr.append((time.perf_counter() - t0) * 1e3)
        # P = <a o Gf, G> + <b o Hf, H> + <a,b> Q with the scalar products done on the host (numpy object ints)
        L = P.L
        ai = [int.from_bytes(a[32 * i : 32 * i + 32], "little") for i in range(n)]
        bi = [int.from_bytes(b[32 * i : 32 * i + 32], "little") for i in range(n)]
        gi = [int.from_bytes(Gf[32 * i : 32 * i + 32], "little") for i in range(n)]
        hi = [int.from_bytes(Hf[32 * i : 32 * i + 32], "little") for i in range(n)]
        c = sum(x * y for x, y in zip(ai, bi)) % L
        sc = b"".join(P.sc_bytes(x) for x in [c] + [ai[i] * gi[i] % L for i in range(n)] + [bi[i] * hi[i] % L for i in range(n)])
        out = ctypes.create_string_buffer(32)
        tabs = (ctypes.c_void_p * 2)(tG._h, tH._h)
        offs = (ctypes.c_size_t * 2)(0, h_off)
        lens = (ctypes.c_size_t * 2)(n, n)
        check(lib().bpg_msm_mixed(ctx._h, Q, 1, tabs, offs, lens, 2, sc, out))
        vr = []
        for it in range(5):
            t0 = time.perf_counter()
            proof.verify(ctx, n, P.Transcript(b"bench"), Gf, Hf, out.raw, Q, tG, tH, h_off=h_off)
            vr.append((time.perf_counter() - t0) * 1e3)
        rows.append({"lg_n": lg, "create_ms": round(med(cr[1:]), 3), "verify_ms": round(med(vr[1:]), 3), "proof_bytes": len(proof.to_bytes())})
        print(rows[-1], file=sys.stderr, flush=True)
        tG.close()
    return rows


def sweep_r1cs(ctx, comb, lgs):
    rows = []
    maxn = 1 << max(lgs)
    Gb = bytes(dev_points(ctx, comb, maxn, 1).cpu().numpy().tobytes())
    Hb = bytes(dev_points(ctx, comb, maxn, 2).cpu().numpy().tobytes())
    Bb = bytes(dev_points(ctx, comb, 1, 4).cpu().numpy().tobytes())
    for lg in lgs:
        n = 1 << lg
        gens = P.Gens(ctx, Gb[: 32 * n], Hb[: 32 * n], BASE, Bb)
        pm, vm, proof = [], [], None
        for it in range(6):
            p = P.Prover(gens, P.Transcript(b"bench r1cs"))
            p.square_chain(p.commit_public(77), n)
            t0 = time.perf_counter()
            proof = p.prove(1234 + it)
            pm.append((time.perf_counter() - t0) * 1e3)
            v = P.Verifier(gens, P.Transcript(b"bench r1cs"))
            v.square_chain(v.commit_public(77), n)
            t0 = time.perf_counter()
            v.verify(proof)
            vm.append((time.perf_counter() - t0) * 1e3)
        rows.append({"lg_multipliers": lg, "prove_ms": round(med(pm[1:]), 3), "verify_ms": round(med(vm[1:]), 3), "proof_bytes": len(proof)})
        print(rows[-1], file=sys.stderr, flush=True)
        gens.close()
    return rows


def random_circuit(ctx, comb, lg=16, m=16):
    """BASELINE.json config 4: 2^lg multipliers, 2^(lg+1) random linear constraints, m committed values."""
    n = 1 << lg
    Gb = bytes(dev_points(ctx, comb, n, 1).cpu().numpy().tobytes())
    Hb = bytes(dev_points(ctx, comb, n, 2).cpu().numpy().tobytes())
    Bb = bytes(dev_points(ctx, comb, 1, 4).cpu().numpy().tobytes())
    gens = P.Gens(ctx, Gb, Hb, BASE, Bb)
    vals = [1000 + 7 * j for j in range(m)]
    pm, vm, proof, c0 = [], [], None, None
    for it in range(5):
        p = P.Prover(gens, P.Transcript(b"bench rand"))
        Vs = [p.commit(v, 5 + j)[0] for j, v in enumerate(vals)]
        c0 = p.random_circuit(99, n, 2 * n)
        t0 = time.perf_counter()
        proof = p.prove(4321 + it)
        pm.append((time.perf_counter() - t0) * 1e3)
        v = P.Verifier(gens, P.Transcript(b"bench rand"))
        for V in Vs:
            v.commit(V)
        v.random_circuit(99, n, 2 * n, c0)
        t0 = time.perf_counter()
        v.verify(proof)
        vm.append((time.perf_counter() - t0) * 1e3)
    gens.close()
    return {"lg_multipliers": lg, "constraints": 2 * n, "terms": 10 * n, "committed": m, "prove_ms": round(med(pm[1:]), 3),
            "verify_ms": round(med(vm[1:]), 3), "proof_bytes": len(proof)}


def sweep_varbase(ctx, comb, max_lg):
    """config 3 as the reference calls it on points it has never seen: bpg_msm on COMPRESSED points + scalars in
    host memory (upload, decode, plain table, Pippenger with Horner), nothing precomputed; wall clock per call."""
    rows = []
    for lg in range(12, max_lg + 1, 2):
        n = 1 << lg
        pts = dev_points(ctx, comb, n, 300 + lg).cpu().pin_memory()
        sc = dev_scalars(n, 400 + lg).cpu().pin_memory()
        res = ctypes.create_string_buffer(32)
        ts = []
        for it in range(5 if lg <= 20 else 3):
            t0 = time.perf_counter()
            check(lib().bpg_msm(ctx._h, ctypes.c_void_p(sc.data_ptr()), ctypes.c_void_p(pts.data_ptr()), n, res))
            ts.append((time.perf_counter() - t0) * 1e3)
        rows.append({"lg_n": lg, "ms": round(min(ts), 4), "mpoints_s": round(n / min(ts) / 1e3, 1)})
        print(rows[-1], file=sys.stderr, flush=True)
        del pts, sc
        torch.cuda.empty_cache()
    return rows


def main():
    what = (sys.argv[1] if len(sys.argv) > 1 else "msm,ipp,r1cs").split(",")
    max_lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    ctx = Context(0)
    comb = Comb(ctx, BASE)
    out = {"gpu": torch.cuda.get_device_name(0)}
    if "msm" in what:
        out["msm"] = sweep_msm(ctx, comb, max_lg)
    if "varbase" in what:
        out["variable_base_msm"] = sweep_varbase(ctx, comb, min(max_lg, 22))
    if "ipp" in what:
        out["ipp"] = sweep_ipp(ctx, comb, [10, 12, 14, 16, 18])
    if "r1cs" in what:
        out["r1cs"] = sweep_r1cs(ctx, comb, [10, 12, 14, 16, 18])
    if "rand" in what:
        out["random_circuit"] = random_circuit(ctx, comb)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
