import os, sys, time
sys.path.insert(0, "/root/repo")
from tools.sweep import *
ctx = Context(0); comb = Comb(ctx, BASE)
lg = int(sys.argv[1]); n = 1 << lg
Gb = bytes(dev_points(ctx, comb, n, 1).cpu().numpy().tobytes())
Hb = bytes(dev_points(ctx, comb, n, 2).cpu().numpy().tobytes())
Bb = bytes(dev_points(ctx, comb, 1, 4).cpu().numpy().tobytes())
gens = P.Gens(ctx, Gb, Hb, BASE, Bb)
p = P.Prover(gens, P.Transcript(b"bench r1cs")); p.square_chain(p.commit_public(77), n)
proof = p.prove(1)
for it in range(8):
    v = P.Verifier(gens, P.Transcript(b"bench r1cs")); v.square_chain(v.commit_public(77), n)
    ctx.profile(True); ctx.profile_reset()
    t0 = time.perf_counter(); v.verify(proof); dt = (time.perf_counter() - t0) * 1e3
    pr = ctx.profile_read(); ctx.profile(False)
    print(round(dt, 3), {k: round(x[0], 3) for k, x in pr.items() if x[1]}, file=sys.stderr)
