"""Shared helpers for the test-suite: seeded scalars and points."""
from __future__ import annotations

import random

from oracle import group as G


def rng(seed):
    return random.Random(seed)


def rand_scalar(r) -> int:
    return r.randrange(G.L)


def rand_point(r) -> G.Point:
    """uniform scalar times the basepoint (like mpc-stark's random_point used by
    reference benches/inner_product.rs:28-30); libsodium does the scalar
    multiplication when PyNaCl is present, the oracle otherwise."""
    k = r.randrange(1, G.L)
    try:
        import nacl.bindings as nb

        enc = nb.crypto_scalarmult_ed25519_base_noclamp(k.to_bytes(32, "little"))
        y = int.from_bytes(enc, "little") & (2**255 - 1)
        sign = enc[31] >> 7
        u, v = (y * y - 1) % G.P, (G.D * y * y + 1) % G.P
        ok, x = G.sqrt_ratio_m1(u, v)
        assert ok
        if (x & 1) != sign:
            x = G.P - x
        return G.Point(x, y, 1, x * y)
    except ImportError:
        return k * G.BASEPOINT


def scalars_bytes(ks) -> bytes:
    return b"".join(G.sc_to_bytes(k) for k in ks)


def points_bytes(ps) -> bytes:
    return b"".join(p.encode() for p in ps)
