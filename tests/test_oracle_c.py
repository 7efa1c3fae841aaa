"""The C restatement (oracle/c/bp_oracle.c, also the timed CPU baseline) against the
big-integer oracle, across the Straus / Pippenger w=6,7,8 size classes."""
import pytest

from oracle import cbind
from oracle import group as G
from tests.util import points_bytes, rand_point, rand_scalar, rng, scalars_bytes


@pytest.mark.parametrize("n", [0, 1, 2, 50, 189, 190, 499, 500, 800, 2000])
def test_c_msm_matches_bigint(n):
    r = rng(n)
    ks = [rand_scalar(r) for _ in range(n)]
    ps = [rand_point(r) for _ in range(n)]
    if n > 3:
        ks[0], ks[1], ks[2] = 0, G.L - 1, 2**252
        ps[3] = G.IDENTITY
    want = G.msm(ks, ps).encode()
    assert cbind.msm(scalars_bytes(ks), points_bytes(ps), 1) == want
    assert cbind.msm(scalars_bytes(ks), points_bytes(ps), 4) == want


def test_c_basepoint_mul_and_rfc_vectors():
    from tests.test_oracle_group import MULTIPLES

    out = cbind.basepoint_mul(scalars_bytes(range(16)))
    assert [out[32 * i : 32 * i + 32].hex() for i in range(16)] == MULTIPLES
    r = rng(3)
    ks = [rand_scalar(r) for _ in range(50)]
    out = cbind.basepoint_mul(scalars_bytes(ks))
    assert all(out[32 * i : 32 * i + 32] == (k * G.BASEPOINT).encode() for i, k in enumerate(ks))


def test_c_rejects_bad_point():
    with pytest.raises(ValueError):
        cbind.msm(scalars_bytes([1]), bytes.fromhex("01" + "00" * 31))
