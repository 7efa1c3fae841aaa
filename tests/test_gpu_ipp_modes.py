"""The three round strategies of the inner-product argument must give the oracle's proof bytes
(`InnerProductProof::create`, reference src/inner_product_proof.rs:49-193, which folds the generators every round):
  * comb rounds on the generators' own combs (short vectors),
  * bucket-method rounds over the original generators, then the folded generators MATERIALISED once from the
    combs and comb rounds over their own combs (long vectors; forced here at small n through the tuning knobs),
  * bucket-method rounds throughout (tables without combs).
Both entry shapes: a caller's windowed table with an arbitrary Q (`bpg_ipp_create`), and the R1CS prover's shared
generator table with Q = w*B (whole-proof bytes), plus the two-lane shares path."""
import os
import random

import pytest

from oracle import gadgets
from oracle import group as G
from oracle import protocol as O
from tests.util import points_bytes

pytestmark = pytest.mark.gpu
L = G.L

MODES = {
    "comb_direct": {"BPG_IPP_DIRECT_MAX": "8192", "BPG_IPP_M0": "2048"},
    "materialize_at_4": {"BPG_IPP_DIRECT_MAX": "0", "BPG_IPP_M0": "4"},
    "materialize_at_16": {"BPG_IPP_DIRECT_MAX": "0", "BPG_IPP_M0": "16"},
    "buckets_only": {"BPG_IPP_DIRECT_MAX": "0", "BPG_IPP_M0": "0"},
}


@pytest.fixture(params=list(MODES))
def mode(request):
    old = {k: os.environ.get(k) for k in MODES[request.param]}
    os.environ.update(MODES[request.param])
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("n", [2, 8, 64, 256])
def test_ipp_create_strategies(ctx, mode, n):
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200 import protocol as P

    r = random.Random(7000 + n)
    bp = O.BulletproofGens(n, 1)
    Gs, Hs = bp.G(n), bp.H(n)
    Q = G.hash_to_group_sha512(b"strategies q")
    a = [r.randrange(L) for _ in range(n)]
    b = [r.randrange(L) for _ in range(n)]
    Gf = [r.randrange(1, L) for _ in range(n)]
    Hf = [r.randrange(1, L) for _ in range(n)]
    want = O.InnerProductProof.create(O.Transcript(b"modes"), Q, Gf, Hf, Gs, Hs, a, b)
    pad = [G.BASEPOINT] * 2
    t = Table(ctx, points_bytes(pad + Gs + pad + Hs)).set_windows(0).build_comb()
    assert t.has_comb
    got = P.InnerProductProof.create(ctx, P.Transcript(b"modes"), Q.encode(), Gf, Hf, t, t, a, b, g_off=2, h_off=4 + n)
    assert got.to_bytes() == want.to_bytes(), mode
    t.close()


@pytest.mark.parametrize("k", [3, 24])
def test_r1cs_proof_strategies(ctx, mode, k):
    """whole R1CS proofs (shuffle gadget: two phases, u-factors, padding) under every strategy"""
    from mpc_bulletproof_b200 import protocol as P

    pc = O.PedersenGens()
    bp = O.BulletproofGens(64, 1)
    gens = P.Gens(ctx, points_bytes(bp.G(64)), points_bytes(bp.H(64)), pc.B.encode(), pc.B_blinding.encode())

    def build(p, r, enc):
        inp = [r.randrange(2**64) for _ in range(k)]
        outp = inp[:]
        r.shuffle(outp)
        ic = [p.commit(v, r.randrange(L)) for v in inp]
        oc = [p.commit(v, r.randrange(L)) for v in outp]
        gadgets.shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
        return [enc(c) for c, _ in ic + oc]

    op = O.Prover(pc, O.Transcript(b"modes"))
    build(op, random.Random(k), lambda c: c.encode())
    want = op.prove(bp, O.Blindings(40 + k)).to_bytes()
    pp = P.Prover(gens, P.Transcript(b"modes"))
    coms = build(pp, random.Random(k), lambda c: c)
    got = pp.prove(40 + k)
    assert got == want, mode
    vf = P.Verifier(gens, P.Transcript(b"modes"))
    vs = [vf.commit(c) for c in coms]
    gadgets.shuffle_gadget(vf, vs[:k], vs[k:])
    vf.verify(got)
    gens.close()
