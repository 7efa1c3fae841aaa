"""Generator derivation on the device: the chains of BulletproofGens::new and
PedersenGens::default (reference src/generators.rs:61-71, 80-125, 182-235) against the oracle's
restatement, and a proof over derived generators against the oracle's proof bytes."""
import hashlib
import random

import pytest

from oracle import group as G
from oracle import protocol as O

pytestmark = pytest.mark.gpu


def test_points_from_uniform(ctx):
    from mpc_bulletproof_b200 import protocol as P

    r = random.Random(5)
    blocks = [bytes(64), b"\xff" * 64] + [r.randbytes(64) for _ in range(300)]
    got = P.points_from_uniform(ctx, b"".join(blocks))
    for i, b in enumerate(blocks):
        assert got[32 * i : 32 * i + 32] == G.from_uniform_bytes(b).encode(), i
    assert P.points_from_uniform(ctx, b"") == b""
    # RFC 9496 A.3 (hash-to-group with SHA-512), first vector
    v = P.points_from_uniform(ctx, hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest())
    assert v.hex() == "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"


def test_chain_and_fast_forward(ctx):
    from mpc_bulletproof_b200 import protocol as P

    label = b"G" + (0).to_bytes(4, "little")
    want = [p.encode() for p in O.GeneratorsChain(label).take(0, 70)]
    got = P.gens_chain(ctx, label, 0, 70)
    assert [got[32 * i : 32 * i + 32] for i in range(70)] == want
    tail = P.gens_chain(ctx, label, 64, 6)  # fast_forward(64)
    assert [tail[32 * i : 32 * i + 32] for i in range(6)] == want[64:]


@pytest.mark.parametrize("party", [0, 3])
def test_derived_gens_equal_the_oracles(ctx, party):
    from mpc_bulletproof_b200 import protocol as P

    cap = 64
    gens = P.Gens.derive(ctx, cap, party)
    bp = O.BulletproofGens(cap, party + 1)
    pc = O.PedersenGens()
    assert gens.G == b"".join(p.encode() for p in bp.G(cap, party))
    assert gens.H == b"".join(p.encode() for p in bp.H(cap, party))
    assert gens.B == pc.B.encode() and gens.B_blinding == pc.B_blinding.encode()
    assert gens.commit(5, 7) == pc.commit(5, 7).encode()
    gens.close()


def test_proof_over_derived_gens(ctx):
    """A proof made over generators derived on the device is the oracle's proof over its own chains,
    byte for byte, and the oracle's verifier accepts it."""
    from mpc_bulletproof_b200 import protocol as P

    gens = P.Gens.derive(ctx, 32)
    pc, bp = O.PedersenGens(), O.BulletproofGens(32, 1)
    n, val, seed = 19, 0xC0FFEE, 77
    p = P.Prover(gens, P.Transcript(b"derived"))
    p.square_chain(p.commit_public(val), n)
    got = p.prove(seed)
    op = O.Prover(pc, O.Transcript(b"derived"))
    var = op.commit_public(val)
    for _ in range(n):
        _, _, var = op.multiply(var, var)
    assert got == op.prove(bp, O.Blindings(seed)).to_bytes()
    v = P.Verifier(gens, P.Transcript(b"derived"))
    v.square_chain(v.commit_public(val), n)
    v.verify(got)
    gens.close()
