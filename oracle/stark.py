"""ORACLE (test infrastructure; never imported by the product).

The Stark curve — the group the mounted reference actually computes over
(`mpc_stark::algebra::stark_curve::StarkPoint`, reference Cargo.toml:13,21; src/generators.rs:11-16):
    y^2 = x^3 + x + beta   over  F_p,  p = 2^251 + 17*2^192 + 1,  prime order n.
mpc-stark itself is not in /root/reference (un-vendored crate, semver ^0.2, no Cargo.lock), so this
is a restatement of the published curve (StarkWare's "STARK curve", constants as in SURVEY.md 8c-2)
with plain affine big-integer arithmetic: PARITY UNPINNED at the dependency boundary.  What pins it:
G on the curve, n*G = O, p and n prime, group laws (tests/test_oracle_stark.py).  Affine coordinates
are unique, so any correct implementation yields the same bytes.

Encoding used at the C ABI (the fork's transcript encoding, reference src/util.rs:274-289): affine
x || y, 32 bytes little-endian each; the identity is 64 zero bytes.  Scalars: 32 bytes little-endian < n.
"""
from __future__ import annotations

P = 2**251 + 17 * 2**192 + 1
N = 0x0800000000000010FFFFFFFFFFFFFFFFB781126DCAE7B2321E66A241ADC64D2F
A = 1
BETA = 0x06F21413EFBE40DE150E596D72F7A8C5609AD26C15C915C1F4CDFCB99CEE9E89
GX = 0x01EF15C18599971B7BECED415A40F0C7DEACFD9B0D1819E03D723D8BC943CFCA
GY = 0x005668060AA49730B7BE4801DF46EC62DE53ECD11ABE43A32873000C36E8DC1F


class Point:
    """Affine point or the identity (inf = True)."""

    __slots__ = ("x", "y", "inf")

    def __init__(self, x=0, y=0, inf=False):
        self.x, self.y, self.inf = x % P, y % P, inf

    def on_curve(self) -> bool:
        return self.inf or (self.y * self.y - (self.x**3 + A * self.x + BETA)) % P == 0

    def __eq__(self, o) -> bool:
        return (self.inf and o.inf) or (not self.inf and not o.inf and self.x == o.x and self.y == o.y)

    def __hash__(self):
        return hash((self.x, self.y, self.inf))

    def __neg__(self) -> "Point":
        return Point(self.x, -self.y, self.inf)

    def __add__(self, o: "Point") -> "Point":
        if self.inf:
            return o
        if o.inf:
            return self
        if self.x == o.x:
            if (self.y + o.y) % P == 0:
                return IDENTITY
            lam = (3 * self.x * self.x + A) * pow(2 * self.y, -1, P) % P
        else:
            lam = (o.y - self.y) * pow(o.x - self.x, -1, P) % P
        x3 = (lam * lam - self.x - o.x) % P
        return Point(x3, lam * (self.x - x3) - self.y)

    def __sub__(self, o: "Point") -> "Point":
        return self + (-o)

    def __rmul__(self, k: int) -> "Point":
        k %= N
        acc, base = IDENTITY, self
        while k:
            if k & 1:
                acc = acc + base
            base = base + base
            k >>= 1
        return acc

    def encode(self) -> bytes:
        if self.inf:
            return bytes(64)
        return self.x.to_bytes(32, "little") + self.y.to_bytes(32, "little")

    @staticmethod
    def decode(b: bytes) -> "Point":
        """None if not a valid encoding (coordinate >= p or point off the curve)."""
        assert len(b) == 64
        if b == bytes(64):
            return IDENTITY
        x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")
        if x >= P or y >= P:
            return None
        pt = Point(x, y)
        return pt if pt.on_curve() else None


IDENTITY = Point(0, 0, True)
GENERATOR = Point(GX, GY)


def sc_to_bytes(k: int) -> bytes:
    return (k % N).to_bytes(32, "little")


def msm(scalars, points) -> Point:
    """sum k_i P_i (Straus-free, plain): the value `StarkPoint::msm_iter` returns."""
    acc = IDENTITY
    for k, p in zip(scalars, points):
        acc = acc + (k % N) * p
    return acc


# ---- the fork's hashing conventions that live in /root/reference (src/util.rs:252-267, src/generators.rs:80-125)
def hash_to_scalar(low: bytes) -> int:
    """(low || keccak256(low)) read as a 512-bit little-endian integer, mod the group order (src/util.rs:252-267:
    the bytes are reversed and handed to from_be_bytes_mod_order)."""
    from .merlin import keccak256

    assert len(low) == 32
    return int.from_bytes(low + keccak256(low), "little") % N


def gens_chain(state0: bytes, skip: int, n: int):
    """GeneratorsChain (src/generators.rs:80-125) from its initial state keccak256(pad_label("GeneratorsChain" || label));
    `pad_label` is the un-vendored merlin fork's, so the state is an input here."""
    from .merlin import keccak256

    st = state0
    for _ in range(skip):  # fast_forward (:92-100)
        st = keccak256(st)
    out = []
    for _ in range(n):
        st = keccak256(st)
        out.append(hash_to_scalar(st) * GENERATOR)
    return out
