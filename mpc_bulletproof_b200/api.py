"""Host-side mirror of the group-level seam of the reference:
`StarkPoint::msm_iter(scalars, points)` (reference src/inner_product_proof.rs:90,
src/r1cs/verifier.rs:516) becomes `msm(ctx, scalars, points)`; generator tables
(`BulletproofGens`, reference src/generators.rs:158-235) become `Table`.
Arguments are bytes (32-byte little-endian scalars, 32-byte compressed points).
"""
from __future__ import annotations

import ctypes
import weakref

from . import _lib
from ._lib import check, lib


class Context:
    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        check(lib().bpg_init(device, ctypes.byref(self._h)))
        self.device = device
        self._children = weakref.WeakSet()  # tables/combs/ipp states: freed before the context

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            lib().bpg_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t handle (0 = legacy default stream); None = the context's own."""
        if cuda_stream is None:
            check(lib().bpg_set_stream(self._h, None, 1))
        else:
            check(lib().bpg_set_stream(self._h, ctypes.c_void_p(cuda_stream), 0))

    def profile(self, on: bool):
        check(lib().bpg_profile_enable(self._h, int(on)))

    def profile_reset(self):
        check(lib().bpg_profile_reset(self._h))

    def profile_read(self) -> dict:
        """{phase: (total_ms, intervals)} since the last reset; waits for the stream."""
        n = 10
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_uint64 * n)()
        check(lib().bpg_profile_read(self._h, ms, cnt, n))
        return {lib().bpg_profile_phase_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n)}

    def sync(self):
        check(lib().bpg_sync(self._h))

    def set_window(self, c: int):
        check(lib().bpg_set_window(self._h, c))

    def set_groups(self, gsub: int):
        check(lib().bpg_set_groups(self._h, gsub))

    @property
    def launches(self) -> int:
        return int(lib().bpg_launch_count(self._h))


class Table:
    """Points resident in HBM in affine-Niels form."""

    def __init__(self, ctx: Context, points: bytes | None = None, *, dev_ptr: int | None = None, n: int | None = None):
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        if points is not None:
            assert len(points) % 32 == 0
            check(lib().bpg_table_upload(ctx._h, points, len(points) // 32, ctypes.byref(self._h)))
        else:
            check(lib().bpg_table_upload_dev(ctx._h, ctypes.c_void_p(dev_ptr), n, ctypes.byref(self._h)))
        ctx._children.add(self)

    def __len__(self):
        return int(lib().bpg_table_len(self._h))

    def set_windows(self, c: int = 0):
        """Precompute 2^(c w) P_i for every window (c = 0: chosen from the table length)."""
        check(lib().bpg_table_set_windows(self.ctx._h, self._h, c))
        return self

    def build_comb(self):
        """64 x 8 multiples per point (48 KB each): lets inner-product rounds over this table skip the bucket method."""
        check(lib().bpg_table_build_comb(self.ctx._h, self._h))
        return self

    @property
    def has_comb(self) -> bool:
        return bool(lib().bpg_table_has_comb(self._h))

    @property
    def entry_bytes(self) -> int:
        return int(lib().bpg_table_entry_bytes(self._h))

    @property
    def window(self) -> int:
        return int(lib().bpg_table_window(self._h))

    def close(self):
        if self._h:
            lib().bpg_table_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def msm(self, scalars: bytes, n_sets: int = 1, offset: int = 0, n: int | None = None) -> list[bytes]:
        """out[s] = sum_i scalars[s*n+i] * table[offset+i]; host buffers."""
        if n is None:
            n = len(self) - offset
        if len(scalars) != 32 * n * n_sets:
            raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalar buffer length does not match n * n_sets")
        out = ctypes.create_string_buffer(32 * n_sets)
        check(lib().bpg_msm_table(self.ctx._h, self._h, offset, n, scalars, n_sets, out))
        return [out.raw[32 * i : 32 * i + 32] for i in range(n_sets)]

    def msm_partial(self, scalars: bytes, n_sets: int = 1, offset: int = 0, n: int | None = None) -> bytes:
        """The sums as extended points (n_sets x 128 bytes), not encoded: one rank's / one MPC party's
        partial result, to be combined with `sum_encode` after the exchange."""
        if n is None:
            n = len(self) - offset
        if len(scalars) != 32 * n * n_sets:
            raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalar buffer length does not match n * n_sets")
        out = ctypes.create_string_buffer(128 * n_sets)
        check(lib().bpg_msm_table_partial(self.ctx._h, self._h, offset, n, scalars, n_sets, out))
        return out.raw

    def dev_msm(self, d_scalars: int, n_sets: int, d_out_ext: int, offset: int = 0, n: int | None = None):
        """Device-resident form; enqueues on the context's stream."""
        if n is None:
            n = len(self) - offset
        check(
            lib().bpg_dev_msm_table(
                self.ctx._h, self._h, offset, n, ctypes.c_void_p(d_scalars), n_sets, ctypes.c_void_p(d_out_ext)
            )
        )


def msm(ctx: Context, scalars: bytes, points: bytes) -> bytes:
    """sum_i scalars[i] * points[i] -> 32-byte compressed point."""
    if len(scalars) != len(points) or len(scalars) % 32:
        raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalars and points differ in length")
    out = ctypes.create_string_buffer(32)
    check(lib().bpg_msm(ctx._h, scalars, points, len(scalars) // 32, out))
    return out.raw


def sum_encode(ctx: Context, parts: bytes, n_parts: int, n_sets: int = 1) -> list[bytes]:
    """out[s] = encode(sum_p parts[p][s]) for partial sums laid out [part][set] x 128 bytes."""
    if len(parts) != 128 * n_parts * n_sets:
        raise _lib.BpgError(_lib.BPG_ERR_LEN, "partials buffer length is not n_parts * n_sets * 128")
    out = ctypes.create_string_buffer(32 * n_sets)
    check(lib().bpg_sum_encode(ctx._h, parts, n_parts, n_sets, out))
    return [out.raw[32 * i : 32 * i + 32] for i in range(n_sets)]


def dev_sum_encode(ctx: Context, d_parts: int, n_parts: int, n_sets: int, d_out_bytes: int | None, d_out_ext: int | None = None):
    check(
        lib().bpg_dev_sum_encode(
            ctx._h, ctypes.c_void_p(d_parts), n_parts, n_sets, ctypes.c_void_p(d_out_bytes or 0), ctypes.c_void_p(d_out_ext or 0)
        )
    )


class Comb:
    """Fixed-base comb tables for a few points (e.g. the Pedersen bases B, B_blinding)."""

    def __init__(self, ctx: Context, bases: bytes):
        assert len(bases) % 32 == 0 and bases
        self.ctx = ctx
        self.nbases = len(bases) // 32
        self._h = ctypes.c_void_p()
        check(lib().bpg_comb_create(ctx._h, bases, self.nbases, ctypes.byref(self._h)))
        ctx._children.add(self)

    def close(self):
        if self._h:
            lib().bpg_comb_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mul(self, scalars: bytes) -> bytes:
        """out[i] = sum_t scalars[t*n+i] * base_t; scalars laid out base-major."""
        n = len(scalars) // (32 * self.nbases)
        if len(scalars) != 32 * n * self.nbases:
            raise _lib.BpgError(_lib.BPG_ERR_LEN, "scalar buffer length is not nbases*n*32")
        out = ctypes.create_string_buffer(32 * max(n, 1))
        check(lib().bpg_comb_mul(self.ctx._h, self._h, scalars, n, out))
        return out.raw[: 32 * n]

    def dev_mul(self, d_scalars: int, n: int, d_out_bytes: int | None, d_out_ext: int | None = None):
        check(
            lib().bpg_dev_comb_mul(
                self.ctx._h, self._h, ctypes.c_void_p(d_scalars), n, ctypes.c_void_p(d_out_bytes or 0), ctypes.c_void_p(d_out_ext or 0)
            )
        )
