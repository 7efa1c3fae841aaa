"""The C restatement (oracle/c/bp_oracle.c, also the timed CPU baseline) against the
big-integer oracle, across the Straus / Pippenger w=6,7,8 size classes."""
import pytest

from oracle import cbind
from oracle import group as G
from tests.util import points_bytes, rand_point, rand_scalar, rng, scalars_bytes


@pytest.mark.parametrize("n", [0, 1, 2, 50, 189, 190, 499, 500, 800, 2000])
def test_c_msm_matches_bigint(n):
    r = rng(n)
    ks = [rand_scalar(r) for _ in range(n)]
    ps = [rand_point(r) for _ in range(n)]
    if n > 3:
        ks[0], ks[1], ks[2] = 0, G.L - 1, 2**252
        ps[3] = G.IDENTITY
    want = G.msm(ks, ps).encode()
    assert cbind.msm(scalars_bytes(ks), points_bytes(ps), 1) == want
    assert cbind.msm(scalars_bytes(ks), points_bytes(ps), 4) == want


def test_c_basepoint_mul_and_rfc_vectors():
    from tests.test_oracle_group import MULTIPLES

    out = cbind.basepoint_mul(scalars_bytes(range(16)))
    assert [out[32 * i : 32 * i + 32].hex() for i in range(16)] == MULTIPLES
    r = rng(3)
    ks = [rand_scalar(r) for _ in range(50)]
    out = cbind.basepoint_mul(scalars_bytes(ks))
    assert all(out[32 * i : 32 * i + 32] == (k * G.BASEPOINT).encode() for i, k in enumerate(ks))


def test_c_rejects_bad_point():
    with pytest.raises(ValueError):
        cbind.msm(scalars_bytes([1]), bytes.fromhex("01" + "00" * 31))


def test_cpu_ipp_create_matches_protocol_oracle():
    """oracle_ipp_create (the timed CPU restatement of InnerProductProof::create,
    src/inner_product_proof.rs:49-193) against oracle/protocol.py on the same inputs, the
    challenges being those the Merlin transcript produced: same L_j, R_j, a, b."""
    import random

    from oracle import cbind
    from oracle import group as G
    from oracle import protocol as O

    class Recorder(O.Transcript):
        def __init__(self, label):
            super().__init__(label)
            self.us = []

        def challenge_scalar(self, label):
            u = super().challenge_scalar(label)
            if label == b"u":
                self.us.append(u)
            return u

    r = random.Random(21)
    for n in (1, 2, 8, 32):
        bp = O.BulletproofGens(n, 1)
        Gs, Hs = bp.G(n), bp.H(n)
        Q = G.hash_to_group_sha512(b"cpu ipp")
        a = [r.randrange(G.L) for _ in range(n)]
        b = [r.randrange(G.L) for _ in range(n)]
        Gf = [1 if i < n // 2 else 5 for i in range(n)]
        Hf = [pow(7, i, G.L) for i in range(n)]
        tr = Recorder(b"cpu ipp test")
        want = O.InnerProductProof.create(tr, Q, Gf, Hf, Gs, Hs, a, b)
        sb = lambda xs: b"".join(G.sc_to_bytes(x) for x in xs)
        pb = lambda ps: b"".join(p.encode() for p in ps)
        for threads in (1, 4):
            lr, fa, fb = cbind.ipp_create(Q.encode(), sb(Gf), sb(Hf), pb(Gs), pb(Hs), sb(a), sb(b), sb(tr.us) or b"\0" * 32, threads)
            assert [x for pair in lr for x in pair] == [p.encode() for pair in zip(want.L_vec, want.R_vec) for p in pair]
            assert fa == G.sc_to_bytes(want.a) and fb == G.sc_to_bytes(want.b)
