"""Pins for oracle/stark.py (the Stark curve of the mounted reference's mpc-stark dependency, which is
not vendored: parity unpinned at that boundary).  Constants as published."""
import random

from oracle import stark as S


def _is_prime(n, rounds=24):
    if n < 2:
        return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    r = random.Random(5)
    for _ in range(rounds):
        a = r.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def test_constants():
    assert S.P == 0x0800000000000011000000000000000000000000000000000000000000000001
    assert _is_prime(S.P) and _is_prime(S.N)
    assert S.GENERATOR.on_curve()
    acc, base, k = S.IDENTITY, S.GENERATOR, S.N
    while k:
        if k & 1:
            acc = acc + base
        base = base + base
        k >>= 1
    assert acc.inf  # n * G = O
    # Hasse: |n - (p + 1)| <= 2 sqrt(p)
    assert (S.N - S.P - 1) ** 2 <= 4 * S.P


def test_group_laws_and_encoding():
    r = random.Random(6)
    G = S.GENERATOR
    a, b, c = (r.randrange(1, S.N) for _ in range(3))
    A, B, C = a * G, b * G, c * G
    assert (A + B) + C == A + (B + C)
    assert A + B == B + A == ((a + b) % S.N) * G
    assert A - A == S.IDENTITY and (A + S.IDENTITY) == A
    assert (a * b % S.N) * G == a * B
    assert S.msm([a, b, c], [G, A, B]) == ((a + b * a + c * b) % S.N) * G
    for pt in (A, B, S.IDENTITY):
        assert S.Point.decode(pt.encode()) == pt
    assert S.IDENTITY.encode() == bytes(64)
    assert S.Point.decode((S.P).to_bytes(32, "little") + bytes(32)) is None


def test_ipp_oracle_roundtrip():
    """The Stark IPP restatement proves and verifies (the reference's own test idiom,
    src/inner_product_proof.rs:474-505) for n = 1, 2, 8 with non-trivial factors."""
    import hashlib

    from oracle import stark_ipp as I

    def challenge(j, L, R):
        return int.from_bytes(hashlib.sha256(b"stark ipp test" + bytes([j]) + L.encode() + R.encode()).digest(), "little")

    r = random.Random(8)
    Gen = S.GENERATOR
    for n in (1, 2, 8):
        Gs = [r.randrange(1, S.N) * Gen for _ in range(n)]
        Hs = [r.randrange(1, S.N) * Gen for _ in range(n)]
        Q = r.randrange(1, S.N) * Gen
        a = [r.randrange(S.N) for _ in range(n)]
        b = [r.randrange(S.N) for _ in range(n)]
        Gf = [r.randrange(1, S.N) for _ in range(n)]
        Hf = [r.randrange(1, S.N) for _ in range(n)]
        c = I.inner_product(a, b)
        P = S.msm([a[i] * Gf[i] for i in range(n)] + [b[i] * Hf[i] for i in range(n)] + [c], Gs + Hs + [Q])
        L, Rv, af, bf = I.create(challenge, Q, Gf, Hf, Gs, Hs, a, b)
        assert len(L) == n.bit_length() - 1
        assert I.verify(challenge, n, L, Rv, af, bf, Gf, Hf, P, Q, Gs, Hs)
        assert not I.verify(challenge, n, L, Rv, (af + 1) % S.N, bf, Gf, Hf, P, Q, Gs, Hs)
