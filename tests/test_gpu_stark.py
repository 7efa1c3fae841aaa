"""GPU parity of the Stark-curve MSM (the group the mounted fork computes over)
against oracle/stark.py: identical affine coordinates, byte for byte.  Same edge vectors as the
ristretto255 suite: zero / one / n-1 scalars, the identity as an input, duplicate points (the
short-Weierstrass formulas are not unified: P + P and P + (-P) take their own paths), repeated
scalars (over-long buckets), several window widths, several sets."""
import pytest

from oracle import stark as S
from tests.util import rng

pytestmark = pytest.mark.gpu
N = S.N


def _pts(r, n):
    return [r.randrange(1, N) * S.GENERATOR for _ in range(n)]


def _sb(ks):
    return b"".join(S.sc_to_bytes(k) for k in ks)


def _pb(ps):
    return b"".join(p.encode() for p in ps)


def _check(ctx, ks, ps):
    from mpc_bulletproof_b200 import stark

    got = stark.msm(ctx, _sb(ks), _pb(ps))
    want = S.msm(ks, ps).encode()
    assert got == want, (got.hex(), want.hex())


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 300])
def test_msm_random(ctx, n):
    r = rng(700 + n)
    _check(ctx, [r.randrange(N) for _ in range(n)], _pts(r, n))


def test_msm_edge_scalars(ctx):
    r = rng(711)
    ks = [0, 1, N - 1, N - 2, 2, 2**128, 2**251, (1 << 251) - 1, N >> 1]
    ps = _pts(r, len(ks))
    _check(ctx, ks, ps)
    for k, p in zip(ks, ps):
        _check(ctx, [k], [p])


def test_msm_edge_points(ctx):
    r = rng(712)
    p, q = _pts(r, 2)
    k1, k2, k3 = (r.randrange(N) for _ in range(3))
    _check(ctx, [k1, k2], [p, p])  # the same point twice: doubling inside a bucket when digits agree
    _check(ctx, [k1, k1], [p, p])
    _check(ctx, [k1, k1], [p, -p])  # cancels to the identity
    _check(ctx, [k1, k2, k3], [S.IDENTITY, p, S.IDENTITY])
    _check(ctx, [k1, N - k1], [p, p])
    _check(ctx, [5, 5, 5], [p, q, p])
    _check(ctx, [0, 0], [p, q])


def test_msm_one_bucket_and_segments(ctx):
    """Repeated scalars: every term of a window in one bucket (multi-segment and block paths)."""
    from mpc_bulletproof_b200 import stark

    r = rng(713)
    m, reps = 256, 12
    ps = _pts(r, m)
    total = S.msm([1] * m, ps)
    k = r.randrange(N)
    for kk in (1, k):
        want = (kk * reps % N) * total
        assert stark.msm(ctx, _sb([kk]) * (m * reps), _pb(ps) * reps) == want.encode()
    bits = [i & 1 for i in range(m)]
    _check(ctx, bits, ps)


@pytest.mark.parametrize("c", [3, 5, 8, 11, 13])
def test_msm_all_windows(ctx, c):
    r = rng(720 + c)
    n = 200
    ks = [r.randrange(N) for _ in range(n)]
    ps = _pts(r, n)
    ctx.set_window(c)
    try:
        _check(ctx, ks, ps)
    finally:
        ctx.set_window(0)


def test_table_sets_and_offset(ctx):
    from mpc_bulletproof_b200.stark import StarkTable

    r = rng(731)
    n_tab, off, n, sets = 400, 60, 256, 3
    ps = _pts(r, n_tab)
    ps[70] = S.IDENTITY
    t = StarkTable(ctx, _pb(ps))
    assert len(t) == n_tab
    ks = [[r.randrange(N) for _ in range(n)] for _ in range(sets)]
    got = t.msm(b"".join(_sb(k) for k in ks), n_sets=sets, offset=off, n=n)
    for s in range(sets):
        assert got[s] == S.msm(ks[s], ps[off : off + n]).encode()
    t.close()


def test_invalid_point_rejected(ctx):
    from mpc_bulletproof_b200 import BpgError, stark
    from mpc_bulletproof_b200._lib import BPG_ERR_DECODE

    g = S.GENERATOR
    off_curve = g.x.to_bytes(32, "little") + ((g.y + 1) % S.P).to_bytes(32, "little")
    not_canonical = S.P.to_bytes(32, "little") + g.y.to_bytes(32, "little")
    for bad in (off_curve, not_canonical):
        with pytest.raises(BpgError) as e:
            stark.msm(ctx, _sb([1]), bad)
        assert e.value.code == BPG_ERR_DECODE


def test_linearity_large(ctx):
    """2^16 terms: tiling a 1024-point table 64 times equals the MSM of the column sums."""
    from mpc_bulletproof_b200 import stark

    r = rng(741)
    m, reps = 1024, 64
    ps = _pts(r, 64)
    ps = [ps[i % 64] + (i // 64) * S.GENERATOR for i in range(m)]  # cheap distinct points
    pb = _pb(ps)
    ks = [[r.randrange(N) for _ in range(m)] for _ in range(reps)]
    big = stark.msm(ctx, b"".join(_sb(k) for k in ks), pb * reps)
    summed = [sum(ks[j][i] for j in range(reps)) % N for i in range(m)]
    small = stark.msm(ctx, _sb(summed), pb)
    assert big == small


@pytest.mark.parametrize("c,gsub", [(0, 0), (5, 0), (8, 3), (13, 1), (13, 20), (16, 0)])
def test_windowed_table(ctx, c, gsub):
    """Tables holding 2^(c w) P_i (affine, batch-inverted): one bucket array per set, no Horner.
    Whole table, a sub-range, two sets, forced bucket groups, identity and edge scalars."""
    from mpc_bulletproof_b200.stark import StarkTable

    r = rng(750 + c)
    n_tab, off, n, sets = 300, 21, 256, 2
    ps = _pts(r, n_tab)
    ps[5] = S.IDENTITY
    t = StarkTable(ctx, _pb(ps)).set_windows(c)
    assert t.window == (c or t.window) and t.window >= 2
    ks = [r.randrange(N) for _ in range(n_tab)]
    ks[0], ks[1], ks[2] = 0, N - 1, 1
    ctx.set_groups(gsub)
    try:
        assert t.msm(_sb(ks))[0] == S.msm(ks, ps).encode()
        kk = [[r.randrange(N) for _ in range(n)] for _ in range(sets)]
        got = t.msm(b"".join(_sb(k) for k in kk), n_sets=sets, offset=off, n=n)
    finally:
        ctx.set_groups(0)
    for s in range(sets):
        assert got[s] == S.msm(kk[s], ps[off : off + n]).encode()
    # repeated scalars over a windowed table: long buckets
    assert t.msm(_sb([7] * n_tab))[0] == (7 * S.msm([1] * n_tab, ps)).encode()
    t.close()


def _challenge(j, Lb, Rb):
    import hashlib

    return int.from_bytes(hashlib.sha256(b"stark ipp test" + bytes([j]) + Lb + Rb).digest(), "little") % N


@pytest.mark.parametrize("n", [1, 2, 4, 32, 64])
def test_ipp_rounds(ctx, n):
    """`InnerProductProof::create` over the Stark curve, round by round (the transcript is the
    caller's): L_j, R_j and the final a, b equal the oracle's, which FOLDS the generators as the
    reference does (src/inner_product_proof.rs:125-134, 226-227) while the product never does; the
    oracle's verifier accepts the result.  Sizes as the reference's make_ipp_* tests (:507-583)."""
    from mpc_bulletproof_b200.stark import StarkIpp, StarkTable
    from oracle import stark_ipp as I

    r = rng(800 + n)
    Gs, Hs = _pts(r, n), _pts(r, n)
    Q = r.randrange(1, N) * S.GENERATOR
    a = [r.randrange(N) for _ in range(n)]
    b = [r.randrange(N) for _ in range(n)]
    y_inv = r.randrange(1, N)
    Gf = [1 if i < n // 2 else 7 for i in range(n)]  # as the R1CS prover: one, then a challenge
    Hf = [pow(y_inv, i, N) for i in range(n)]
    want_L, want_R, want_a, want_b = I.create(lambda j, L, R: _challenge(j, L.encode(), R.encode()), Q, Gf, Hf, Gs, Hs, a, b)
    tG, tH = StarkTable(ctx, _pb(Gs)), StarkTable(ctx, _pb(Hs)).set_windows(0)
    st = StarkIpp(ctx, tG, tH, Q.encode(), _sb(Gf), _sb(Hf), _sb(a), _sb(b))
    assert st.rounds_left() == n.bit_length() - 1
    Ls, Rs = [], []
    j = 0
    while st.rounds_left():
        L, R = st.round_lr()
        assert L == want_L[j].encode() and R == want_R[j].encode(), f"round {j}"
        u = _challenge(j, L, R)
        st.round_fold(S.sc_to_bytes(u), S.sc_to_bytes(pow(u, -1, N)))
        Ls.append(S.Point.decode(L))
        Rs.append(S.Point.decode(R))
        j += 1
    fa, fb = st.finish()
    assert fa == S.sc_to_bytes(want_a) and fb == S.sc_to_bytes(want_b)
    c = I.inner_product(a, b)
    P = S.msm([a[i] * Gf[i] for i in range(n)] + [b[i] * Hf[i] for i in range(n)] + [c], Gs + Hs + [Q])
    assert I.verify(lambda k, L, R: _challenge(k, L.encode(), R.encode()), n, Ls, Rs, want_a, want_b, Gf, Hf, P, Q, Gs, Hs)
    st.close()
    tG.close()
    tH.close()


def test_fork_generator_chain(ctx):
    """the fork's GeneratorsChain (reference src/generators.rs:80-125): keccak hash chain -> hash_to_scalar -> k * G,
    from a given initial state (the `pad_label`ed seed is the un-vendored merlin fork's); with fast_forward."""
    from mpc_bulletproof_b200 import stark as PS
    from oracle.merlin import keccak256

    state0 = keccak256(b"GeneratorsChain" + b"G" + (0).to_bytes(4, "little"))  # stand-in for pad_label(...)
    want = S.gens_chain(state0, 0, 24)
    got = PS.gens_chain(ctx, state0, 0, 24)
    assert [got[64 * i : 64 * i + 64] for i in range(24)] == [p.encode() for p in want]
    assert PS.gens_chain(ctx, state0, 7, 9) == b"".join(p.encode() for p in want[7:16])
    assert PS.gens_chain(ctx, state0, 3, 0) == b""
