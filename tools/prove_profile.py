"""Phase breakdown (CUDA events on the launch stream) of one R1CS prove + verify at 2^lg multipliers
on the reference's DummyCircuit (benches/r1cs.rs:24-32).  Development aid; prints JSON."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_protocol import BASE, rand_scalars, synth_points  # noqa: E402

from mpc_bulletproof_b200 import Comb, Context  # noqa: E402
from mpc_bulletproof_b200 import protocol as P  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    table_c = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # window width of the generator table (0: the library picks)
    n = 1 << lg
    ctx = Context(0)
    ctx.set_window(table_c)
    comb = Comb(ctx, BASE)
    gens = P.Gens(ctx, synth_points(ctx, comb, n, 1), synth_points(ctx, comb, n, 2), BASE, synth_points(ctx, comb, 1, 4))

    ctx.set_window(0)

    def build(cs, val):
        cs.square_chain(cs.commit_public(val), n)

    val = rand_scalars(1, 99)[0]
    out = {"lg": lg, "table_c": gens.table.window, "leaf": os.environ.get("BPG_LEAF", ""), "runs": []}
    # unprofiled wall-clock first (the profile's events serialise the auxiliary stream)
    walls = []
    for it in range(6):
        p = P.Prover(gens, P.Transcript(b"bench r1cs"))
        build(p, val)
        t0 = time.perf_counter()
        proof = p.prove(99 + it)
        walls.append((time.perf_counter() - t0) * 1e3)
    out["prove_ms_unprofiled"] = sorted(walls[1:])
    for it in range(reps):
        p = P.Prover(gens, P.Transcript(b"bench r1cs"))
        build(p, val)
        ctx.profile(True)
        ctx.profile_reset()
        l0 = ctx.launches
        t0 = time.perf_counter()
        proof = p.prove(1234 + it)
        t_prove = (time.perf_counter() - t0) * 1e3
        prof = ctx.profile_read()
        launches = ctx.launches - l0
        ctx.profile(False)
        out["runs"].append({"prove_ms": t_prove, "launches": launches,
                            "phases_ms_total": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
                            "phase_counts": {k: v[1] for k, v in prof.items() if v[1]}})
    # verification of the last proof: wall-clock unprofiled, then the phases
    walls = []
    for it in range(6):
        v = P.Verifier(gens, P.Transcript(b"bench r1cs"))
        build(v, val)
        t0 = time.perf_counter()
        v.verify(proof)
        walls.append((time.perf_counter() - t0) * 1e3)
    out["verify_ms_unprofiled"] = sorted(walls[1:])
    v = P.Verifier(gens, P.Transcript(b"bench r1cs"))
    build(v, val)
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launches
    t0 = time.perf_counter()
    v.verify(proof)
    t_verify = (time.perf_counter() - t0) * 1e3
    prof = ctx.profile_read()
    out["verify"] = {"verify_ms": t_verify, "launches": ctx.launches - l0,
                     "phases_ms_total": {k: round(v_[0], 3) for k, v_ in prof.items() if v_[1]}}
    ctx.profile(False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
