"""ORACLE (test infrastructure).  `InnerProductProof::create` / `verify` of the mounted reference
(src/inner_product_proof.rs:49-193, 202-248, 254-372) restated over the Stark curve (oracle/stark.py)
exactly as the fork runs it: the generators ARE folded every round (G' = u^-1 G_lo + u G_hi,
H' = u H_lo + u^-1 H_hi, :226-227) and multiplied by their factors in round 0 (:125-134) -- the
product never folds them, so equal L/R bytes every round are a real check.

The fork's transcript (`merlin::HashChainTranscript`, an un-vendored git dependency) is not in
/root/reference; the challenge source is therefore a parameter: `challenge(round, L, R) -> scalar`.
"""
from __future__ import annotations

from . import stark as S

N = S.N


def inner_product(a, b):  # :463-472
    assert len(a) == len(b)
    return sum(x * y for x, y in zip(a, b)) % N


def create(challenge, Q, G_factors, H_factors, G_vec, H_vec, a_vec, b_vec):
    """-> (L_vec, R_vec, a, b); points are oracle.stark.Point, scalars ints mod N."""
    n = len(G_vec)
    assert n == len(H_vec) == len(a_vec) == len(b_vec) == len(G_factors) == len(H_factors)  # :59-66
    assert n & (n - 1) == 0 and n > 0  # :69
    G, H, a, b = list(G_vec), list(H_vec), [x % N for x in a_vec], [x % N for x in b_vec]
    L_vec, R_vec = [], []
    rnd = 0
    if n != 1:  # round 0 with factors, :80-146
        n //= 2
        a_lo, a_hi, b_lo, b_hi = a[:n], a[n:], b[:n], b[n:]
        G_lo, G_hi, H_lo, H_hi = G[:n], G[n:], H[:n], H[n:]
        c_L, c_R = inner_product(a_lo, b_hi), inner_product(a_hi, b_lo)  # :87-88
        L = S.msm(
            [a_lo[i] * G_factors[n + i] for i in range(n)] + [b_hi[i] * H_factors[i] for i in range(n)] + [c_L],
            G_hi + H_lo + [Q],
        )  # :90-101
        R = S.msm(
            [a_hi[i] * G_factors[i] for i in range(n)] + [b_lo[i] * H_factors[n + i] for i in range(n)] + [c_R],
            G_lo + H_hi + [Q],
        )  # :103-114
        L_vec.append(L)
        R_vec.append(R)
        u = challenge(rnd, L, R) % N  # :119-122
        u_inv = pow(u, -1, N)
        rnd += 1
        G = [(G_factors[i] % N) * G[i] for i in range(2 * n)]  # :125-134
        H = [(H_factors[i] % N) * H[i] for i in range(2 * n)]
        G_lo, G_hi, H_lo, H_hi = G[:n], G[n:], H[:n], H[n:]
        a, b, G, H = _fold(u, u_inv, a_lo, a_hi, b_lo, b_hi, G_lo, G_hi, H_lo, H_hi)
    while n != 1:  # :149-185
        n //= 2
        a_lo, a_hi, b_lo, b_hi = a[:n], a[n:], b[:n], b[n:]
        G_lo, G_hi, H_lo, H_hi = G[:n], G[n:], H[:n], H[n:]
        c_L, c_R = inner_product(a_lo, b_hi), inner_product(a_hi, b_lo)
        L = S.msm(a_lo + b_hi + [c_L], G_hi + H_lo + [Q])  # :159
        R = S.msm(a_hi + b_lo + [c_R], G_lo + H_hi + [Q])  # :166
        L_vec.append(L)
        R_vec.append(R)
        u = challenge(rnd, L, R) % N
        u_inv = pow(u, -1, N)
        rnd += 1
        a, b, G, H = _fold(u, u_inv, a_lo, a_hi, b_lo, b_hi, G_lo, G_hi, H_lo, H_hi)
    return L_vec, R_vec, a[0], b[0]


def _fold(u, u_inv, a_lo, a_hi, b_lo, b_hi, G_lo, G_hi, H_lo, H_hi):  # fold_witness :202-248
    n = len(a_lo)
    a = [(a_lo[i] * u + u_inv * a_hi[i]) % N for i in range(n)]
    b = [(b_lo[i] * u_inv + u * b_hi[i]) % N for i in range(n)]
    G = [S.msm([u_inv, u], [G_lo[i], G_hi[i]]) for i in range(n)]
    H = [S.msm([u, u_inv], [H_lo[i], H_hi[i]]) for i in range(n)]
    return a, b, G, H


def verify(challenge, n, L_vec, R_vec, a, b, G_factors, H_factors, P, Q, G, H) -> bool:
    """:254-372: verification_scalars (challenges replayed, batch inverse, s vector) and the check
    P == a*b*Q + <a s o g, G> + <b s^-1 o h, H> - sum u_j^2 L_j - sum u_j^-2 R_j."""
    lg_n = len(L_vec)
    if lg_n >= 32 or n != (1 << lg_n):
        return False
    us = [challenge(j, L_vec[j], R_vec[j]) % N for j in range(lg_n)]
    u_inv = [pow(u, -1, N) for u in us]
    allinv = 1
    for x in u_inv:
        allinv = allinv * x % N
    u_sq = [u * u % N for u in us]
    u_inv_sq = [x * x % N for x in u_inv]
    s = [allinv]
    for i in range(1, n):  # :300-307
        lg_i = i.bit_length() - 1
        k = 1 << lg_i
        s.append(s[i - k] * u_sq[(lg_n - 1) - lg_i] % N)
    g_s = [a * s[i] % N * G_factors[i] % N for i in range(n)]
    h_s = [b * s[n - 1 - i] % N * H_factors[i] % N for i in range(n)]
    expect = S.msm([a * b % N] + g_s + h_s + [(-x) % N for x in u_sq] + [(-x) % N for x in u_inv_sq], [Q] + G + H + L_vec + R_vec)
    return expect == P
