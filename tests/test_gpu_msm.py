"""GPU parity: libbpgpu's MSM (through the C ABI) against the oracle.

Bit-exact bar: the 32-byte ristretto255 encodings must be identical.
Mirrors the reference's test idiom (random inputs, reference
src/inner_product_proof.rs:474-505) and adds the edge vectors SURVEY.md §8c lists.
"""
import pytest

from oracle import group as G
from tests.util import points_bytes, rand_point, rand_scalar, rng, scalars_bytes

pytestmark = pytest.mark.gpu


def _check(ctx, ks, ps):
    from mpc_bulletproof_b200 import msm

    got = msm(ctx, scalars_bytes(ks), points_bytes(ps))
    want = G.msm(ks, ps).encode()
    assert got == want, (got.hex(), want.hex())


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 255, 1000])
def test_msm_random(ctx, n):
    r = rng(100 + n)
    _check(ctx, [rand_scalar(r) for _ in range(n)], [rand_point(r) for _ in range(n)])


def test_msm_4096(ctx):
    r = rng(7)
    n = 4096
    _check(ctx, [rand_scalar(r) for _ in range(n)], [rand_point(r) for _ in range(n)])


def test_msm_edge_scalars(ctx):
    r = rng(11)
    ks = [0, 1, G.L - 1, 2**252, 2**252 + 1, G.L - 2, 2, 2**128, 2**251, (1 << 252) - 1]
    ps = [rand_point(r) for _ in ks]
    _check(ctx, ks, ps)
    for k, p in zip(ks, ps):
        _check(ctx, [k], [p])


def test_msm_edge_points(ctx):
    r = rng(12)
    p, q = rand_point(r), rand_point(r)
    k1, k2, k3 = rand_scalar(r), rand_scalar(r), rand_scalar(r)
    _check(ctx, [k1, k2], [p, p])            # duplicate points
    _check(ctx, [k1, k1], [p, -p])           # cancels to the identity
    _check(ctx, [k1, k2, k3], [G.IDENTITY, p, G.IDENTITY])  # identity as an input
    _check(ctx, [k1, G.L - k1], [p, p])      # scalars cancel
    _check(ctx, [5, 5], [p, q])


def test_msm_one_bucket(ctx):
    """All scalars equal: every term of a window lands in one bucket (the
    over-long-bucket path), as bit-valued witness vectors do in the reference's
    range gadget (tests/r1cs.rs:629-632)."""
    r = rng(13)
    n = 3000
    ps = [rand_point(r) for _ in range(n)]
    _check(ctx, [1] * n, ps)
    k = rand_scalar(r)
    _check(ctx, [k] * n, ps)
    _check(ctx, [i & 1 for i in range(n)], ps)


@pytest.mark.parametrize("c", [3, 4, 7, 8, 11, 13, 16])
def test_msm_all_windows(ctx, c):
    r = rng(20 + c)
    n = 300
    ks = [rand_scalar(r) for _ in range(n)]
    ps = [rand_point(r) for _ in range(n)]
    ctx.set_window(c)
    try:
        _check(ctx, ks, ps)
    finally:
        ctx.set_window(0)


def test_msm_table_sets_and_offset(ctx):
    from mpc_bulletproof_b200 import Table

    r = rng(31)
    n_tab, off, n, sets = 600, 100, 384, 3
    ps = [rand_point(r) for _ in range(n_tab)]
    t = Table(ctx, points_bytes(ps))
    assert len(t) == n_tab
    ks = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
    got = t.msm(b"".join(scalars_bytes(k) for k in ks), n_sets=sets, offset=off, n=n)
    for s in range(sets):
        assert got[s] == G.msm(ks[s], ps[off : off + n]).encode()
    t.close()


def test_invalid_point_rejected(ctx):
    from mpc_bulletproof_b200 import BpgError, msm
    from mpc_bulletproof_b200._lib import BPG_ERR_DECODE

    bad = bytes.fromhex("0100000000000000000000000000000000000000000000000000000000000000")
    with pytest.raises(BpgError) as e:
        msm(ctx, scalars_bytes([1]), bad)
    assert e.value.code == BPG_ERR_DECODE


def test_linearity_large(ctx):
    """Size-independent property at 2^18 terms: tiling a 2^12-point table 64 times
    must equal the MSM of the column-summed scalars over the 2^12 points."""
    import numpy as np

    from mpc_bulletproof_b200 import msm

    r = rng(41)
    m, reps = 4096, 64
    ps = [rand_point(r) for _ in range(m)]
    pb = points_bytes(ps)
    ks = [[rand_scalar(r) for _ in range(m)] for _ in range(reps)]
    big = msm(ctx, b"".join(scalars_bytes(k) for k in ks), pb * reps)
    summed = [sum(ks[j][i] for j in range(reps)) % G.L for i in range(m)]
    small = msm(ctx, scalars_bytes(summed), pb)
    assert big == small
    assert small == G.msm(summed, ps).encode()


@pytest.mark.parametrize("c", [0, 5, 8, 13])
def test_windowed_table(ctx, c):
    """Tables holding the precomputed window multiples 2^(c w) P_i give the same sums
    (whole table, sub-range, several scalar sets)."""
    from mpc_bulletproof_b200 import Table

    r = rng(50 + c)
    n_tab, off, n, sets = 500, 37, 400, 2
    ps = [rand_point(r) for _ in range(n_tab)]
    ps[5] = G.IDENTITY
    t = Table(ctx, points_bytes(ps)).set_windows(c)
    assert t.window == (c or t.window) and t.window >= 2
    ks = [rand_scalar(r) for _ in range(n_tab)]
    ks[0], ks[1], ks[2] = 0, G.L - 1, 2**252
    assert t.msm(scalars_bytes(ks))[0] == G.msm(ks, ps).encode()
    kk = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
    got = t.msm(b"".join(scalars_bytes(k) for k in kk), n_sets=sets, offset=off, n=n)
    for s in range(sets):
        assert got[s] == G.msm(kk[s], ps[off : off + n]).encode()
    t.close()


def test_msm_segmented_bucket(ctx):
    """A bucket far longer than one block's segment (2048 entries): 512 points tiled 16 times
    with one repeated scalar is 8192 entries per occupied bucket; the sum must be 16 k sum(P_i).
    Structured witnesses (bit vectors, reference tests/r1cs.rs:629-632) hit this path at scale."""
    from mpc_bulletproof_b200 import Table, msm

    r = rng(61)
    m, reps = 512, 16
    ps = [rand_point(r) for _ in range(m)]
    pb = points_bytes(ps)
    total = G.msm([1] * m, ps)
    for k in (1, rand_scalar(r)):
        want = (k * reps % G.L) * total
        assert msm(ctx, scalars_bytes([k]) * (m * reps), pb * reps) == want.encode()
    # same through a windowed table, bits as scalars
    t = Table(ctx, pb * reps).set_windows(0)
    bits = [i & 1 for i in range(m * reps)]
    want = G.msm([reps] * (m // 2), ps[1::2])
    assert t.msm(scalars_bytes(bits))[0] == want.encode()
    t.close()


@pytest.mark.parametrize("c,gsub", [(8, 1), (8, 3), (8, 32), (13, 1), (13, 2), (13, 20), (16, 1), (16, 5)])
def test_windowed_groups(ctx, c, gsub):
    """Windowed tables share one bucket array per set; the entries of a bucket are split into
    `gsub` groups by window index.  Every split must give the same sum."""
    from mpc_bulletproof_b200 import Table

    r = rng(70 + c)
    n, sets = 700, 2
    ps = [rand_point(r) for _ in range(n)]
    t = Table(ctx, points_bytes(ps)).set_windows(c)
    kk = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
    kk[0][0], kk[0][1], kk[1][0] = G.L - 1, 0, 2**252
    want = [G.msm(k, ps).encode() for k in kk]
    ctx.set_groups(gsub)
    try:
        got = t.msm(b"".join(scalars_bytes(k) for k in kk), n_sets=sets)
    finally:
        ctx.set_groups(0)
    assert got == want
    t.close()


@pytest.mark.parametrize("c", [18, 20])
def test_windowed_large_bucket_array(ctx, c):
    """Window widths used at the benchmark sizes (c = 20 at 2^20 points): one array of 2^(c-1)
    buckets per set, the thread-per-chunk reduction leaf and several k_reduce_pairs rounds.
    Checked on a table small enough for the oracle; sparse buckets and empty chunks included."""
    from mpc_bulletproof_b200 import Table

    r = rng(90 + c)
    n = 1500
    ps = [rand_point(r) for _ in range(n)]
    t = Table(ctx, points_bytes(ps)).set_windows(c)
    assert t.window == c
    ks = [rand_scalar(r) for _ in range(n)]
    ks[0], ks[1], ks[2], ks[3] = 0, 1, G.L - 1, 2**252
    assert t.msm(scalars_bytes(ks))[0] == G.msm(ks, ps).encode()
    # two sets, a sub-range, and a bucket schedule forced to several groups
    kk = [[rand_scalar(r) for _ in range(700)] for _ in range(2)]
    want = [G.msm(k, ps[100:800]).encode() for k in kk]
    for gsub in (0, 3):
        ctx.set_groups(gsub)
        try:
            assert t.msm(b"".join(scalars_bytes(k) for k in kk), n_sets=2, offset=100, n=700) == want
        finally:
            ctx.set_groups(0)
    t.close()


def test_full_size_linearity_windowed(ctx):
    """BASELINE.json's headline size (2^20 terms over a resident windowed table, c chosen by the
    library: 20) through a size-independent property: a 2^10-point set tiled 1024 times must give
    the MSM of the column-summed scalars over the 2^10 points, which the oracle confirms."""
    import numpy as np

    from mpc_bulletproof_b200 import Table

    r = rng(97)
    m, reps = 1024, 1024
    ps = [rand_point(r) for _ in range(m)]
    pb = points_bytes(ps)
    big_table = Table(ctx, pb * reps).set_windows(0)
    assert len(big_table) == 1 << 20 and big_table.window >= 18
    g = np.random.Generator(np.random.PCG64(5))
    raw = g.integers(0, 256, size=(m * reps, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0F  # < 2^252 < l
    big = big_table.msm(raw.tobytes())[0]
    big_table.close()
    # column sums mod l with exact integers (object dtype would be slow: split into 4 x 64-bit limbs)
    limbs = raw.view("<u8").reshape(reps, m, 4).astype(object)
    col = limbs.sum(axis=0)
    summed = [sum(int(col[i][k]) << (64 * k) for k in range(4)) % G.L for i in range(m)]
    small_table = Table(ctx, pb)
    small = small_table.msm(scalars_bytes(summed))[0]
    small_table.close()
    assert big == small
    assert small == G.msm(summed, ps).encode()


def test_host_scalars_in_pieces_ragged(ctx):
    """Host scalars of a large launch are uploaded in pieces with the digit histogram of each piece
    running behind its copy (bpg_msm_table, >= 2^18 terms).  A term count that is not a multiple of
    the piece granularity, two sets, and the same launch from device-resident scalars must agree,
    and the column-sum property ties the result to the oracle."""
    import numpy as np
    import torch

    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200.api import dev_sum_encode

    r = rng(98)
    m, reps = 1024, 257
    n = (1 << 18) + 777
    ps = [rand_point(r) for _ in range(m)]
    pb = points_bytes(ps)
    t = Table(ctx, pb * reps).set_windows(0)
    g = np.random.Generator(np.random.PCG64(6))
    raw = g.integers(0, 256, size=(2, n, 32), dtype=np.uint8)
    raw[:, :, 31] &= 0x0F
    got = t.msm(raw.tobytes(), n_sets=2, n=n)
    # device-resident scalars: no upload, one histogram launch
    dev = torch.device("cuda", 0)
    d_sc = torch.from_numpy(raw.reshape(-1).copy()).to(dev)
    d_ext = torch.empty(2 * 128, dtype=torch.uint8, device=dev)
    d_out = torch.empty(64, dtype=torch.uint8, device=dev)
    t.dev_msm(d_sc.data_ptr(), 2, d_ext.data_ptr(), n=n)
    dev_sum_encode(ctx, d_ext.data_ptr(), 1, 2, d_out.data_ptr())
    ctx.sync()
    res = bytes(d_out.cpu().tolist())
    assert got == [res[:32], res[32:]]
    # oracle: column sums over the tiled 1024 points
    for s in range(2):
        col = [0] * m
        limbs = raw[s].view("<u8").astype(object)
        for k in range(4):
            part = np.zeros(m, dtype=object)
            idx = np.arange(n) % m
            np.add.at(part, idx, limbs[:, k])
            for i in range(m):
                col[i] += int(part[i]) << (64 * k)
        summed = [c % G.L for c in col]
        assert got[s] == G.msm(summed, ps).encode()
    t.close()


@pytest.mark.parametrize("n", [31, 32, 33, 46, 1024, 1025])
def test_msm_few_terms_path(ctx, n):
    """Up to 1024 terms over a plain table take the quad-per-term path (k_msm_small: no sort, no
    buckets); 1025 is the first size of the bucket pipeline.  Block boundaries (32 terms per block),
    edge scalars and cancelling points included; the same inputs through the bucket pipeline
    (forced window) must give the same bytes."""
    r = rng(300 + n)
    ks = [rand_scalar(r) for _ in range(n)]
    ps = [rand_point(r) for _ in range(n)]
    ks[0], ks[1], ks[2] = 0, G.L - 1, 2**252
    ps[3] = G.IDENTITY
    ps[5], ks[5] = -ps[4], ks[4]  # k P + k (-P)
    _check(ctx, ks, ps)
    from mpc_bulletproof_b200 import msm

    direct = msm(ctx, scalars_bytes(ks), points_bytes(ps))
    ctx.set_window(6)
    try:
        assert msm(ctx, scalars_bytes(ks), points_bytes(ps)) == direct
    finally:
        ctx.set_window(0)


def test_few_terms_sets(ctx):
    """Several output sets over a plain table on the quad-per-term path (terms of a block belong to
    different sets; blocks are added by k_msm_small_fin)."""
    from mpc_bulletproof_b200 import Table

    r = rng(77)
    n = 75
    ps = [rand_point(r) for _ in range(n)]
    t = Table(ctx, points_bytes(ps))
    kk = [[rand_scalar(r) for _ in range(n)] for _ in range(3)]
    assert t.msm(b"".join(scalars_bytes(k) for k in kk), n_sets=3) == [G.msm(k, ps).encode() for k in kk]
    kk2 = [[rand_scalar(r) for _ in range(20)] for _ in range(2)]
    assert t.msm(b"".join(scalars_bytes(k) for k in kk2), n_sets=2, offset=30, n=20) == [G.msm(k, ps[30:50]).encode() for k in kk2]
    t.close()


@pytest.mark.parametrize("kind", ["all_equal", "bits", "two_values"])
def test_radix_sort_overfull_partition(ctx, kind):
    """Structured scalars put more pairs into ONE partition of the two-pass radix sort than its shared-memory
    placement holds (k_rs_finish stores the overflow directly), and into one bucket than a segment holds; the
    points are k_i*B, so the expected sum is one fixed-base multiplication.  Windowed and plain tables, both sorts."""
    import os

    from mpc_bulletproof_b200 import Comb, Table

    r = rng(91)
    n = 40000
    base = G.BASEPOINT.encode()
    ks = [rand_scalar(r) for _ in range(n)]
    pts = Comb(ctx, base).mul(scalars_bytes(ks))
    if kind == "all_equal":
        v = rand_scalar(r)
        ss = [v] * n
    elif kind == "bits":
        ss = [r.getrandbits(1) for _ in range(n)]
    else:
        a, b = rand_scalar(r), G.L - 1
        ss = [a if r.getrandbits(1) else b for _ in range(n)]
    want = (sum(s * k for s, k in zip(ss, ks)) % G.L * G.BASEPOINT).encode()
    plain, win = Table(ctx, pts), Table(ctx, pts).set_windows(0)
    old = os.environ.get("BPG_SORT")
    try:
        assert plain.msm(scalars_bytes(ss))[0] == want
        assert win.msm(scalars_bytes(ss))[0] == want
    finally:
        if old is None:
            os.environ.pop("BPG_SORT", None)
    plain.close()
    win.close()


def test_adhoc_prefetch_msm_mixed(ctx):
    """bpg_adhoc_prefetch + bpg_msm_mixed: the verifier's call shape (reference src/r1cs/verifier.rs:516-547: a few
    proof points, then ranges of the resident generator table).  The combs built ahead of the scalars must give the
    same bytes as the plain path and as the oracle; a prefetch of OTHER points must not be used; an invalid encoding
    among the prefetched points is reported by the MSM."""
    import ctypes

    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200._lib import lib

    r = rng(4242)
    n_tab = 96
    tab_pts = [rand_point(r) for _ in range(n_tab)]
    t = Table(ctx, points_bytes(tab_pts))
    t.set_windows(0)

    def mixed(adhoc, ks):
        out = ctypes.create_string_buffer(32)
        tabs = (ctypes.c_void_p * 1)(t._h)
        offs = (ctypes.c_size_t * 1)(8)
        lens = (ctypes.c_size_t * 1)(n_tab - 8)
        rc = lib().bpg_msm_mixed(ctx._h, adhoc, len(adhoc) // 32, tabs, offs, lens, 1, scalars_bytes(ks), out)
        return rc, out.raw

    for n_adhoc in (1, 7, 44, 130, 300):  # 300: more than a prefetch holds (256), the plain path
        pts = [rand_point(r) for _ in range(n_adhoc)]
        # edge scalars among the ad-hoc terms: 0, 1, l - 1
        ks = [rand_scalar(r) for _ in range(n_adhoc + n_tab - 8)]
        ks[0] = G.L - 1
        if n_adhoc > 2:
            ks[1], ks[2] = 0, 1
        want = G.msm(ks, pts + tab_pts[8:]).encode()
        enc = points_bytes(pts)
        rc, plain = mixed(enc, ks)
        assert rc == 0 and plain == want
        assert lib().bpg_adhoc_prefetch(ctx._h, enc, n_adhoc) == 0
        rc, got = mixed(enc, ks)
        assert rc == 0 and got == want, n_adhoc
        # the resident combs stay valid for a second call with other scalars
        ks2 = [rand_scalar(r) for _ in ks]
        rc, got = mixed(enc, ks2)
        assert rc == 0 and got == G.msm(ks2, pts + tab_pts[8:]).encode()
        # other points than the prefetched ones: plain path, same answer
        pts3 = [rand_point(r) for _ in range(n_adhoc)]
        rc, got = mixed(points_bytes(pts3), ks)
        assert rc == 0 and got == G.msm(ks, pts3 + tab_pts[8:]).encode()
    # an invalid encoding (a non-canonical field element) among prefetched points
    bad = bytearray(points_bytes([rand_point(r) for _ in range(5)]))
    bad[32 * 3 : 32 * 4] = b"\xff" * 32
    assert lib().bpg_adhoc_prefetch(ctx._h, bytes(bad), 5) == 0
    rc, _ = mixed(bytes(bad), [rand_scalar(r) for _ in range(5 + n_tab - 8)])
    assert rc != 0
    t.close()
