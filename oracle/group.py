"""ORACLE (test infrastructure, never on the product path).

Big-integer restatement of the group layer the hot path computes in: the
ristretto255 prime-order group over edwards25519 (RFC 9496) and scalars mod
the group order.  In the reference this layer is the third-party dependency the
protocol code calls into (`StarkPoint::msm_iter`, `StarkPoint::msm`,
`Scalar * StarkPoint`: reference src/inner_product_proof.rs:90-114,226-227,
src/r1cs/verifier.rs:516-547, src/generators.rs:41-43); it is not vendored in
/root/reference (Cargo.toml:21 `mpc-stark = "0.2"`), so the published
algorithm (RFC 9496 §4, RFC 8032 §5.1) is restated here and pinned by the
RFC's own known-answer vectors (tests/test_oracle_group.py) and by libsodium
through PyNaCl at the Edwards level.  SURVEY.md §0-D1 records why the
ristretto255 instantiation is the primary one.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import hashlib

# --- field GF(2^255 - 19) ----------------------------------------------------
P = 2**255 - 19
D = (-121665 * pow(121666, P - 2, P)) % P
D2 = (2 * D) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
# group order
L = 2**252 + 27742317777372353535851937790883648493


def _is_neg(x: int) -> bool:
    """RFC 9496 §4.1 IS_NEGATIVE: lsb of the canonical encoding."""
    return (x % P) & 1 == 1


def _abs(x: int) -> int:
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u: int, v: int):
    """RFC 9496 §4.2 SQRT_RATIO_M1 -> (was_square, r)."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (P - u) % P
    flipped_i = check == (P - u) * SQRT_M1 % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    return (correct or flipped), _abs(r)


INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, (-1 - D) % P)[1]
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P
# RFC 9496 §4.1: sqrt(a*d - 1) with a = -1, the root listed in the RFC.
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P


# --- extended twisted-Edwards points (X:Y:Z:T), a = -1 -----------------------
class Point:
    """A ristretto255 element held as an edwards25519 extended point."""

    __slots__ = ("X", "Y", "Z", "T")

    def __init__(self, X, Y, Z, T):
        self.X, self.Y, self.Z, self.T = X % P, Y % P, Z % P, T % P

    # RFC 8032 §5.1.4 unified addition (complete for a = -1, d non-square)
    def __add__(self, o: "Point") -> "Point":
        A = (self.Y - self.X) * (o.Y - o.X) % P
        B = (self.Y + self.X) * (o.Y + o.X) % P
        C = self.T * D2 % P * o.T % P
        Dd = 2 * self.Z * o.Z % P
        E, F, G, H = B - A, Dd - C, Dd + C, B + A
        return Point(E * F, G * H, F * G, E * H)

    def __neg__(self) -> "Point":
        return Point(-self.X, self.Y, self.Z, -self.T)

    def __sub__(self, o: "Point") -> "Point":
        return self + (-o)

    def double(self) -> "Point":
        return self + self

    def __rmul__(self, k: int) -> "Point":
        k %= L
        acc, base = IDENTITY, self
        while k:
            if k & 1:
                acc = acc + base
            base = base.double()
            k >>= 1
        return acc

    def __eq__(self, o) -> bool:
        # RFC 9496 §4.3.3 equality on ristretto cosets
        return (self.X * o.Y - self.Y * o.X) % P == 0 or (self.Y * o.Y - self.X * o.X) % P == 0

    def __hash__(self):
        return hash(self.encode())

    def is_identity(self) -> bool:
        return self == IDENTITY

    def affine(self):
        zi = pow(self.Z, P - 2, P)
        return self.X * zi % P, self.Y * zi % P

    def encode(self) -> bytes:
        """RFC 9496 §4.3.2."""
        x0, y0, z0, t0 = self.X, self.Y, self.Z, self.T
        u1 = (z0 + y0) * (z0 - y0) % P
        u2 = x0 * y0 % P
        _, invsqrt = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
        den1 = invsqrt * u1 % P
        den2 = invsqrt * u2 % P
        z_inv = den1 * den2 % P * t0 % P
        ix0 = x0 * SQRT_M1 % P
        iy0 = y0 * SQRT_M1 % P
        enchanted = den1 * INVSQRT_A_MINUS_D % P
        rotate = _is_neg(t0 * z_inv)
        if rotate:
            x, y, den_inv = iy0, ix0, enchanted
        else:
            x, y, den_inv = x0, y0, den2
        if _is_neg(x * z_inv):
            y = (-y) % P
        s = _abs(den_inv * (z0 - y) % P)
        return s.to_bytes(32, "little")


IDENTITY = Point(0, 1, 1, 0)
_BY = 4 * pow(5, P - 2, P) % P
_BX = 0x216936D3CD6E53FEC0A4E231FDD6DC5C692CC7609525A7B2C9562D608F25D51A
BASEPOINT = Point(_BX, _BY, 1, _BX * _BY)


def decode(b: bytes):
    """RFC 9496 §4.3.1; returns a Point or None for an invalid encoding."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= P or s & 1:
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2_sqr = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2_sqr) % P
    was_square, invsqrt = sqrt_ratio_m1(1, v * u2_sqr % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = _abs(2 * s * den_x % P)
    y = u1 * den_y % P
    t = x * y % P
    if (not was_square) or _is_neg(t) or y == 0:
        return None
    return Point(x, y, 1, t)


def _map(t: int) -> Point:
    """RFC 9496 §4.3.4 MAP (Elligator 2)."""
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    was_square, s = sqrt_ratio_m1(u, v)
    s_prime = (-_abs(s * t)) % P
    if not was_square:
        s = s_prime
        c = r
    else:
        c = P - 1
    N = (c * ((r - 1) % P) % P * D_MINUS_ONE_SQ - v) % P
    ss = s * s % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - ss) % P
    w3 = (1 + ss) % P
    return Point(w0 * w3, w2 * w1, w1 * w3, w0 * w2)


def from_uniform_bytes(b: bytes) -> Point:
    """RFC 9496 §4.3.4 element derivation from 64 uniform bytes."""
    assert len(b) == 64
    t1 = int.from_bytes(b[:32], "little") & (2**255 - 1)
    t2 = int.from_bytes(b[32:], "little") & (2**255 - 1)
    return _map(t1 % P) + _map(t2 % P)


def hash_to_group_sha512(msg: bytes) -> Point:
    return from_uniform_bytes(hashlib.sha512(msg).digest())


# --- scalars -------------------------------------------------------------------
def sc_from_bytes_le(b: bytes) -> int:
    return int.from_bytes(b, "little")


def sc_from_wide(b: bytes) -> int:
    assert len(b) == 64
    return int.from_bytes(b, "little") % L


def sc_to_bytes(s: int) -> bytes:
    return (s % L).to_bytes(32, "little")


def sc_inv(s: int) -> int:
    return pow(s % L, L - 2, L)


# --- multiscalar multiplication: the definition, term by term -----------------
def msm_naive(scalars, points) -> Point:
    """sum_i k_i * P_i — what `StarkPoint::msm_iter` returns at every call site."""
    acc = IDENTITY
    for k, pt in zip(scalars, points, strict=True):
        acc = acc + (k % L) * pt
    return acc


def msm(scalars, points, c: int = 0) -> Point:
    """Same sum through unsigned c-bit Pippenger buckets, so that sizes up to
    ~2^12 stay within seconds of pure Python.  Equal to msm_naive by
    construction; tests check that on small inputs."""
    scalars = [k % L for k in scalars]
    n = len(scalars)
    assert n == len(points)
    if n == 0:
        return IDENTITY
    if c == 0:
        c = max(2, min(12, n.bit_length() - 2))
    nwin = (253 + c - 1) // c
    total = IDENTITY
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = total.double()
        buckets = [None] * (1 << c)
        for k, pt in zip(scalars, points):
            d = (k >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = pt if buckets[d] is None else buckets[d] + pt
        run, acc = IDENTITY, IDENTITY
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = run + buckets[d]
            acc = acc + run
        total = total + acc
    return total
