"""The device arithmetic (fe.cuh / sc.cuh / ge.cuh limb schedules, digit recoding and
the Pippenger bookkeeping) compiled for the host with the PTX carry flag modelled,
against the big-integer oracle.  This is what can be checked without a GPU; the
kernels themselves are checked by the -m gpu tests."""
import ctypes
import os
import random

import pytest

from oracle import group as G

HS = os.path.join(os.path.dirname(__file__), "hostsim", "libhostsim.so")
pytestmark = pytest.mark.skipif(not os.path.exists(HS), reason="tests/hostsim/libhostsim.so not built (run __graft_entry__.build())")

U32x8 = ctypes.c_uint32 * 8
E32 = ctypes.c_uint32 * 32
N24 = ctypes.c_uint32 * 24
B32 = ctypes.c_uint8 * 32


@pytest.fixture(scope="module")
def hs():
    return ctypes.CDLL(HS)


def tol(x):
    return U32x8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def frl(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


EDGE_FE = [0, 1, G.P - 1, G.P, G.P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 39, 19, 38, 2**255 + 18]


def rnd_fe(r):
    return r.choice(EDGE_FE) if r.random() < 0.15 else r.getrandbits(256)


def test_field(hs):
    r = random.Random(1)
    o = U32x8()
    for _ in range(4000):
        a, b = rnd_fe(r), rnd_fe(r)
        hs.hs_fe_mul(tol(a), tol(b), o)
        assert frl(o) < 2**255 and frl(o) % G.P == a * b % G.P
        hs.hs_fe_sq(tol(a), o)
        assert frl(o) < 2**255 and frl(o) % G.P == a * a % G.P
        hs.hs_fe_add(tol(a), tol(b), o)
        assert frl(o) % G.P == (a + b) % G.P
        hs.hs_fe_sub(tol(a), tol(b), o)
        assert frl(o) % G.P == (a - b) % G.P
        hs.hs_fe_canon(tol(a), o)
        assert frl(o) == a % G.P
    for _ in range(20):
        a = rnd_fe(r)
        hs.hs_fe_invert(tol(a), o)
        assert frl(o) % G.P == pow(a, G.P - 2, G.P)


def test_scalars(hs):
    r = random.Random(2)
    o = U32x8()
    R = 2**256
    edge = [0, 1, G.L - 1, G.L - 2, 2**252, 2**252 - 1]
    for _ in range(3000):
        a = r.choice(edge) if r.random() < 0.1 else r.randrange(G.L)
        b = r.choice(edge) if r.random() < 0.1 else r.randrange(G.L)
        hs.hs_sc_montmul(tol(a), tol(b), o)
        assert frl(o) == a * b * pow(R, -1, G.L) % G.L
        hs.hs_sc_mul(tol(a), tol(b), o)
        assert frl(o) == a * b % G.L
        hs.hs_sc_add(tol(a), tol(b), o)
        assert frl(o) == (a + b) % G.L
        hs.hs_sc_sub(tol(a), tol(b), o)
        assert frl(o) == (a - b) % G.L


def test_signed_digits(hs):
    r = random.Random(3)
    dg = (ctypes.c_int * 200)()
    for c in range(2, 21):
        for k in [0, 1, G.L - 1, 2**252, 2**253 - 1] + [r.randrange(G.L) for _ in range(100)]:
            W = hs.hs_digits(tol(k), c, dg)
            assert W == (255 + c - 1) // c
            assert sum(dg[w] << (c * w) for w in range(W)) == k
            assert all(-(1 << (c - 1)) <= dg[w] < (1 << (c - 1)) for w in range(W))


def _ext(r, pt):
    z = r.randrange(1, G.P)
    x, y = pt.affine()
    vals = [x * z % G.P, y * z % G.P, z, x * y % G.P * z % G.P]
    e = E32()
    for j, v in enumerate(vals):
        for i in range(8):
            e[8 * j + i] = (v >> (32 * i)) & 0xFFFFFFFF
    return e


def _pt(e):
    v = [sum(int(e[8 * j + i]) << (32 * i) for i in range(8)) for j in range(4)]
    assert (v[0] * v[1] - v[2] * v[3]) % G.P == 0
    assert all(x < 2**255 for x in v), "extended coordinates must stay tight"
    return G.Point(*v)


def test_group(hs):
    from tests.test_oracle_group import BAD, MULTIPLES

    r = random.Random(4)
    pts = [r.randrange(G.L) * G.BASEPOINT for _ in range(8)] + [G.IDENTITY]
    for p in pts:
        enc = p.encode()
        e = E32()
        assert hs.hs_decode(B32(*enc), e) == 1 and _pt(e).encode() == enc
        o = B32()
        hs.hs_encode(_ext(r, p), o)
        assert bytes(o) == enc
        d = E32()
        hs.hs_dbl(_ext(r, p), d)
        assert _pt(d).encode() == (p + p).encode()
        for q in pts[:3] + [G.IDENTITY, p, -p]:
            s = E32()
            hs.hs_add(_ext(r, p), _ext(r, q), s)
            assert _pt(s).encode() == (p + q).encode()
            n = N24()
            hs.hs_to_niels(_ext(r, q), n)
            for neg in (0, 1):
                hs.hs_madd(_ext(r, p), n, neg, s)
                assert _pt(s).encode() == ((p - q) if neg else (p + q)).encode()
    for i, h in enumerate(MULTIPLES):
        e = E32()
        assert hs.hs_decode(B32(*bytes.fromhex(h)), e) == 1
        assert _pt(e).encode() == (i * G.BASEPOINT).encode()
    for h in BAD:
        assert hs.hs_decode(B32(*bytes.fromhex(h)), E32()) == 0, h


@pytest.mark.parametrize("n,c,chunk", [(1, 4, 4), (5, 3, 2), (33, 5, 4), (64, 8, 32), (100, 6, 8)])
def test_pipeline_walkthrough(hs, n, c, chunk):
    r = random.Random(10 * n + c)
    ks = [r.randrange(G.L) for _ in range(n)]
    ps = [r.randrange(G.L) * G.BASEPOINT for _ in range(n)]
    o = B32()
    sb = b"".join(G.sc_to_bytes(k) for k in ks)
    pb = b"".join(p.encode() for p in ps)
    assert hs.hs_msm(sb, pb, n, c, chunk, o) == 0
    assert bytes(o) == G.msm_naive(ks, ps).encode()


def test_field_on_sixteen_lanes(hs):
    """fe16.cuh (the latency form of the field layer: radix 2^16, one limb per lane of a half-warp,
    lazy carries) with one host thread per lane: products, carried sums and differences, the
    steady-state limb bounds the header states, and the 252-squaring chain."""
    U16 = ctypes.c_uint32 * 16
    r = random.Random(16)
    o = U32x8()
    for _ in range(150):
        a, b = rnd_fe(r) % 2**255, rnd_fe(r) % 2**255  # operands are tight, as everywhere on the device
        limbs = U16()
        hs.hs_fe16_op(tol(a), tol(b), 0, 0, o, limbs)
        assert frl(o) == a * b % G.P
        assert limbs[0] < 2**22 and all(limbs[k] < 2**18 for k in range(1, 16))
        hs.hs_fe16_op(tol(a), tol(b), 1, 0, o, limbs)
        assert frl(o) == (a + b) % G.P and all(limbs[k] < 2**17 for k in range(16))
        hs.hs_fe16_op(tol(a), tol(b), 2, 0, o, limbs)
        assert frl(o) == (a - b) % G.P and all(limbs[k] < 2**17 for k in range(16))
        hs.hs_fe16_mix(tol(a), tol(b), o)
        assert frl(o) == pow(((a - b) * (a + b) - a), 2, G.P)
    for a, b in [(2**255 - 1, 2**255 - 1), (G.P - 1, G.P - 1), (0, 5), (1, 1), (2**255 - 1, 1)]:
        limbs = U16()
        hs.hs_fe16_op(tol(a), tol(b), 3, 40, o, limbs)  # 40 squarings of the product: bounds are a fixed point
        assert frl(o) == pow(a * b, 2**40, G.P)
        assert limbs[0] < 2**22 and all(limbs[k] < 2**18 for k in range(1, 16))
    for a, b in [(2**255 - 1, 2**255 - 1), (r.getrandbits(255), r.getrandbits(255)), (1, 0)]:
        limbs = U16()
        hs.hs_fe16_wide_chain(tol(a), tol(b), 25, o, limbs)  # whole-warp products: same values, same bounds
        assert frl(o) == pow(a * b, 2**25, G.P)
        assert limbs[0] < 2**22 and all(limbs[k] < 2**18 for k in range(1, 16))
    for a in [3, G.P - 2, r.getrandbits(255)]:
        hs.hs_fe16_pow22523(tol(a), o)
        assert frl(o) == pow(a, 2**252 - 3, G.P)


def test_encode_on_sixteen_lanes(hs):
    """ge_encode16 = ge_encode byte for byte: RFC 9496 multiples of the generator, random points in
    random projective representations (both rotations, negative x / y), the identity and the other
    points of its coset."""
    from tests.test_oracle_group import MULTIPLES

    r = random.Random(17)
    pts = [k * G.BASEPOINT for k in range(16)] + [r.randrange(G.L) * G.BASEPOINT for _ in range(24)]
    for i, p in enumerate(pts):
        want = p.encode()
        if i < 16:
            assert want.hex() == MULTIPLES[i]
        for _ in range(2):
            e = _ext(r, p)
            o, o1 = B32(), B32()
            hs.hs_encode16(e, o)
            hs.hs_encode(e, o1)
            assert bytes(o) == want == bytes(o1)
            if i in (0, 1, 7, 16, 17, 18):  # 32 host threads per run: keep the CPU suite short
                o2 = B32()
                hs.hs_encode32(e, o2)  # whole-warp form: the half-warps split every product
                assert bytes(o2) == want


def test_decode_on_sixteen_lanes(hs):
    """ge_decode16 (fe16.cuh; the verifier's proof points are decoded one per warp before their doubling chains
    start): the same point and the same verdict as ge_decode on RFC 9496 A.1 multiples, its 28 invalid encodings,
    random points and random strings; whole-warp form on a few of them."""
    from tests.test_oracle_group import BAD, MULTIPLES

    r = random.Random(16)
    cases = [bytes.fromhex(h) for h in MULTIPLES] + [bytes.fromhex(h) for h in BAD]
    cases += [(r.randrange(G.L) * G.BASEPOINT).encode() for _ in range(12)]
    cases += [bytes(r.randrange(256) for _ in range(31)) + bytes([r.randrange(128)]) for _ in range(24)]
    cases += [b"\xff" * 32, b"\x00" * 32, (G.P - 1).to_bytes(32, "little"), (G.P + 2).to_bytes(32, "little")]
    for i, enc in enumerate(cases):
        e0, e1 = E32(), E32()
        ok0 = hs.hs_decode(B32(*enc), e0)
        ok1 = hs.hs_decode16(B32(*enc), 0, e1)
        assert ok0 == ok1, enc.hex()
        def coords(e):  # the same field elements (ge_decode leaves loose representatives, ge_decode16 canonical ones)
            return [sum(int(e[8 * j + k]) << (32 * k) for k in range(8)) % G.P for j in range(4)]

        if ok0:
            assert coords(e0) == coords(e1) and _pt(e1).encode() == enc, enc.hex()
        if i % 9 == 0:  # 32 host threads per run: keep the CPU suite short
            e2 = E32()
            assert hs.hs_decode16(B32(*enc), 1, e2) == ok0
            if ok0:
                assert coords(e0) == coords(e2)


def test_element_derivation(hs):
    """ge_elligator_map / from_uniform_bytes (generator chains, reference src/generators.rs:107-125)
    against the oracle, which RFC 9496 A.3's hash-to-group vectors pin (tests/test_oracle_group.py)."""
    import hashlib

    r = random.Random(18)
    blocks = [bytes(64), b"\xff" * 64, hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()]
    blocks += [r.randbytes(64) for _ in range(40)]
    for b in blocks:
        o = B32()
        hs.hs_from_uniform((ctypes.c_uint8 * 64)(*b), o)
        assert bytes(o) == G.from_uniform_bytes(b).encode()
    assert bytes(o) != bytes(32)


def test_chacha_blinding_scalar(hs):
    """The device side of the keyed blinding vectors (sc.cuh: sc_from_chacha_block) against the oracle's
    ChaCha20 block (RFC 8439) reduced mod l: element e of a phase is block e under the drawn key."""
    import struct

    from oracle import protocol as O

    r = random.Random(77)
    for _ in range(20):
        key = bytes(r.getrandbits(8) for _ in range(32))
        block = r.choice([0, 1, 2, 2**31, 2**32 - 1, r.getrandbits(32)])
        out = U32x8()
        hs.hs_chacha_scalar(U32x8(*struct.unpack("<8I", key)), ctypes.c_uint32(block), out)
        want = int.from_bytes(O.chacha20_block(key, block, O.BLIND_NONCE), "little") % G.L
        assert frl(out) == want
