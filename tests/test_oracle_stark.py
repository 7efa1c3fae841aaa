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
