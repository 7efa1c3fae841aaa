"""Stark-curve policy (csrc/stark_fp.cuh, stark_pt.cuh) compiled for the host with the PTX carry
flag modelled, against the big-integer oracle (oracle/stark.py): the sparse Montgomery reduction
mod p = 2^251 + 17*2^192 + 1 and the XYZZ group formulas including their exceptional cases."""
import ctypes
import os
import random

import pytest

from oracle import stark as S

HS = os.path.join(os.path.dirname(__file__), "hostsim", "libhostsim.so")
pytestmark = pytest.mark.skipif(not os.path.exists(HS), reason="tests/hostsim/libhostsim.so not built (run __graft_entry__.build())")

U8 = ctypes.c_uint32 * 8
U16 = ctypes.c_uint32 * 16
U32 = ctypes.c_uint32 * 32
R = 2**256
P = S.P


@pytest.fixture(scope="module")
def hs():
    return ctypes.CDLL(HS)


def w8(x):
    return U8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def val(a, n=8, off=0):
    return sum(int(a[off + i]) << (32 * i) for i in range(n))


EDGE = [0, 1, 2, P - 1, P - 2, 17 << 192, (1 << 251), (1 << 192) - 1, (1 << 192), 0xFFFFFFFF, P >> 1]


def rnd(r):
    return r.choice(EDGE) if r.random() < 0.2 else r.randrange(P)


def test_field(hs):
    r = random.Random(11)
    o = U8()
    rinv = pow(R, -1, P)
    for _ in range(5000):
        a, b = rnd(r), rnd(r)
        hs.hs_fp_mul(w8(a), w8(b), o)
        assert val(o) == a * b * rinv % P
        hs.hs_fp_add(w8(a), w8(b), o)
        assert val(o) == (a + b) % P
        hs.hs_fp_sub(w8(a), w8(b), o)
        assert val(o) == (a - b) % P
        hs.hs_fp_to_mont(w8(a), o)
        assert val(o) == a * R % P
        hs.hs_fp_from_mont(w8(a), o)
        assert val(o) == a * rinv % P
    for _ in range(10):
        a = r.randrange(1, P)
        hs.hs_fp_invert(w8(a * R % P), o)  # Montgomery in, Montgomery out
        assert val(o) == pow(a, -1, P) * R % P


def _aff_words(pt):
    b = pt.encode()
    return U16(*[int.from_bytes(b[4 * i : 4 * i + 4], "little") for i in range(16)])


def _to_point(hs, x):
    out = U16()
    hs.hs_sp_to_affine(x, out)
    raw = b"".join(int(out[i]).to_bytes(4, "little") for i in range(16))
    pt = S.Point.decode(raw)
    assert pt is not None
    return pt


def _load(hs, pt):
    aff = U16()
    assert hs.hs_sp_from_affine(_aff_words(pt), aff) == 1
    return aff


def test_group_laws(hs):
    r = random.Random(12)
    pts = [S.IDENTITY, S.GENERATOR] + [r.randrange(1, S.N) * S.GENERATOR for _ in range(6)]
    x = U32()
    for p in pts:
        for q in pts + [-p]:
            ap, aq = _load(hs, p), _load(hs, q)
            xp, xq = U32(), U32()
            hs.hs_sp_from_aff(ap, xp)
            hs.hs_sp_from_aff(aq, xq)
            for neg in (0, 1):
                hs.hs_sp_madd(xp, aq, neg, x)
                assert _to_point(hs, x) == (p - q if neg else p + q)
            hs.hs_sp_add(xp, xq, x)
            assert _to_point(hs, x) == p + q
            # non-trivial ZZ on both sides: (p + g) + (q + g) and ((p + g) + g) mixed
            g = _load(hs, S.GENERATOR)
            x1, x2 = U32(), U32()
            hs.hs_sp_madd(xp, g, 0, x1)
            hs.hs_sp_madd(xq, g, 0, x2)
            hs.hs_sp_add(x1, x2, x)
            assert _to_point(hs, x) == p + q + 2 * S.GENERATOR
            hs.hs_sp_madd(x1, aq, 1, x)
            assert _to_point(hs, x) == p + S.GENERATOR - q
        hs.hs_sp_dbl(xp, x)
        assert _to_point(hs, x) == p + p


def test_decode_rejects(hs):
    aff = U16()
    g = S.GENERATOR
    off = S.Point(g.x, g.y + 1)
    b = off.x.to_bytes(32, "little") + off.y.to_bytes(32, "little")
    assert hs.hs_sp_from_affine(U16(*[int.from_bytes(b[4 * i : 4 * i + 4], "little") for i in range(16)]), aff) == 0
    big = (S.P).to_bytes(32, "little") + g.y.to_bytes(32, "little")  # x = p: not canonical
    assert hs.hs_sp_from_affine(U16(*[int.from_bytes(big[4 * i : 4 * i + 4], "little") for i in range(16)]), aff) == 0
    assert hs.hs_sp_from_affine(_aff_words(S.IDENTITY), aff) == 1


def test_scalar_field(hs):
    """stark_sc.cuh: Montgomery arithmetic modulo the Stark group order."""
    r = random.Random(13)
    o = U8()
    n = S.N
    rinv = pow(R, -1, n)
    edge = [0, 1, n - 1, n - 2, (1 << 251), 0xFFFFFFFF, (1 << 192) - 1]
    for _ in range(4000):
        a = r.choice(edge) if r.random() < 0.2 else r.randrange(n)
        b = r.choice(edge) if r.random() < 0.2 else r.randrange(n)
        hs.hs_scs_montmul(w8(a), w8(b), o)
        assert val(o) == a * b * rinv % n
        hs.hs_scs_add(w8(a), w8(b), o)
        assert val(o) == (a + b) % n
        hs.hs_scs_sub(w8(a), w8(b), o)
        assert val(o) == (a - b) % n
    # any 256-bit left operand is reduced (conversion of canonical bytes and of products)
    for _ in range(200):
        a, b = r.getrandbits(256), r.randrange(n)
        hs.hs_scs_montmul(w8(a), w8(b), o)
        assert val(o) == a * b * rinv % n
