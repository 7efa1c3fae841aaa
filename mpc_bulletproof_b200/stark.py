"""Stark-curve instantiation of the group-level seam: the mounted fork's
`StarkPoint::msm_iter(scalars, points)` (mpc-stark; reference src/inner_product_proof.rs:90,
src/r1cs/verifier.rs:516) as `msm(ctx, scalars, points)`, generator tables as `StarkTable`.
Scalars are 32-byte little-endian (mod the group order), points are affine x || y, 32 bytes
little-endian each (reference src/util.rs:274-289), the identity 64 zero bytes."""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import check, lib
from .api import Context


class StarkTable:
    """Stark-curve points resident in HBM (affine, Montgomery form)."""

    def __init__(self, ctx: Context, points_xy: bytes):
        assert len(points_xy) % 64 == 0
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        check(lib().bpg_stark_table_upload(ctx._h, points_xy, len(points_xy) // 64, ctypes.byref(self._h)))
        ctx._children.add(self)

    def __len__(self):
        return int(lib().bpg_stark_table_len(self._h))

    def set_windows(self, c: int = 0):
        """Precompute 2^(c w) P_i for every window (c = 0: chosen from the table length)."""
        check(lib().bpg_stark_table_set_windows(self.ctx._h, self._h, c))
        return self

    @property
    def window(self) -> int:
        return int(lib().bpg_stark_table_window(self._h))

    def close(self):
        if self._h:
            lib().bpg_stark_table_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def msm(self, scalars: bytes, n_sets: int = 1, offset: int = 0, n: int | None = None) -> list[bytes]:
        if n is None:
            n = len(self) - offset
        if len(scalars) != 32 * n * n_sets:
            raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalar buffer length does not match n * n_sets")
        out = ctypes.create_string_buffer(64 * n_sets)
        check(lib().bpg_stark_msm_table(self.ctx._h, self._h, offset, n, scalars, n_sets, out))
        return [out.raw[64 * i : 64 * i + 64] for i in range(n_sets)]


def msm(ctx: Context, scalars: bytes, points_xy: bytes) -> bytes:
    """sum_i scalars[i] * points[i] -> 64-byte affine x || y."""
    if len(scalars) * 2 != len(points_xy) or len(scalars) % 32:
        raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalars and points differ in length")
    out = ctypes.create_string_buffer(64)
    check(lib().bpg_stark_msm(ctx._h, scalars, points_xy, len(scalars) // 32, out))
    return out.raw


class StarkIpp:
    """Device-resident state of `InnerProductProof::create` over the Stark curve (reference
    src/inner_product_proof.rs:49-193); the caller owns the transcript:

        st = StarkIpp(ctx, tG, tH, Q, Gf, Hf, a, b)
        while st.rounds_left():
            L, R = st.round_lr();  u = challenge(L, R);  st.round_fold(u, u_inv)
        a, b = st.finish()
    """

    def __init__(self, ctx: Context, G: StarkTable, H: StarkTable, Q_xy: bytes, G_factors, H_factors, a: bytes, b: bytes,
                 g_off: int = 0, h_off: int = 0):
        n = len(a) // 32
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        check(lib().bpg_stark_ipp_begin(ctx._h, G._h, g_off, H._h, h_off, n, Q_xy, G_factors, H_factors, a, b, ctypes.byref(self._h)))
        ctx._children.add(self)

    def rounds_left(self) -> int:
        return int(lib().bpg_stark_ipp_rounds_left(self._h))

    def round_lr(self):
        L, R = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        check(lib().bpg_stark_ipp_round_LR(self._h, L, R))
        return L.raw, R.raw

    def round_fold(self, u: bytes, u_inv: bytes):
        check(lib().bpg_stark_ipp_round_fold(self._h, u, u_inv))

    def finish(self):
        a, b = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        check(lib().bpg_stark_ipp_finish(self._h, a, b))
        return a.raw, b.raw

    def close(self):
        if self._h:
            lib().bpg_stark_ipp_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def keccak256(data: bytes) -> bytes:
    """legacy Keccak-256, the fork's `merlin::keccak256`"""
    out = ctypes.create_string_buffer(32)
    lib().bpg_keccak256(data, len(data), out)
    return out.raw


def hash_to_scalar(low: bytes) -> int:
    """reference src/util.rs:252-267"""
    assert len(low) == 32
    out = ctypes.create_string_buffer(32)
    lib().bpg_stark_hash_to_scalar(low, out)
    return int.from_bytes(out.raw, "little")


def gens_chain(ctx, state0: bytes, skip: int, n: int) -> bytes:
    """points [skip, skip + n) of the fork's generator chain (src/generators.rs:80-125) from its initial state"""
    assert len(state0) == 32
    out = ctypes.create_string_buffer(64 * max(n, 1))
    check(lib().bpg_stark_gens_chain(ctx._h, state0, skip, n, out))
    return out.raw[: 64 * n]
