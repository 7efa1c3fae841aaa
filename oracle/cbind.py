"""ctypes binding of the C oracle (oracle/c/bp_oracle.c). Test/bench infrastructure only."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "c", "libbp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_PATH)
        L.oracle_msm.restype = ctypes.c_int
        L.oracle_msm.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int]
        L.oracle_points_decode.restype = ctypes.c_void_p
        L.oracle_points_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        L.oracle_points_free.argtypes = [ctypes.c_void_p]
        L.oracle_msm_decoded.restype = ctypes.c_int
        L.oracle_msm_decoded.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int]
        L.oracle_basepoint_mul.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_ipp_create.restype = ctypes.c_int
        L.oracle_ipp_create.argtypes = [ctypes.c_size_t] + [ctypes.c_char_p] * 10 + [ctypes.c_int]
        _lib = L
    return _lib


def msm(scalars: bytes, points: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    assert len(points) == 32 * n
    out = ctypes.create_string_buffer(32)
    rc = lib().oracle_msm(scalars, points, n, out, threads)
    if rc:
        raise ValueError(f"oracle_msm rc={rc}")
    return out.raw


def basepoint_mul(scalars: bytes, threads: int = 0) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(32 * n)
    lib().oracle_basepoint_mul(scalars, n, out, threads or max_threads())
    return out.raw


def max_threads() -> int:
    return int(lib().oracle_max_threads())


class DecodedPoints:
    def __init__(self, points: bytes, threads: int = 0):
        self.n = len(points) // 32
        self._h = lib().oracle_points_decode(points, self.n, threads or max_threads())
        if not self._h:
            raise ValueError("invalid point encoding")

    def msm(self, scalars: bytes, offset: int = 0, n: int | None = None, threads: int = 1) -> bytes:
        n = self.n - offset if n is None else n
        out = ctypes.create_string_buffer(32)
        rc = lib().oracle_msm_decoded(scalars, self._h, offset, n, out, threads)
        if rc:
            raise ValueError(f"oracle_msm_decoded rc={rc}")
        return out.raw

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_points_free(self._h)
            self._h = None


def ipp_create(Q: bytes, G_factors, H_factors, G: bytes, H: bytes, a: bytes, b: bytes, challenges: bytes, threads: int = 1):
    """CPU restatement of InnerProductProof::create with the challenges as inputs
    -> ([(L_j, R_j)], a, b) as bytes."""
    n = len(a) // 32
    lg = n.bit_length() - 1
    out_lr = ctypes.create_string_buffer(64 * max(lg, 1))
    out_ab = ctypes.create_string_buffer(64)
    rc = lib().oracle_ipp_create(n, Q, G_factors, H_factors, G, H, a, b, challenges, out_lr, out_ab, threads)
    if rc:
        raise ValueError(f"oracle_ipp_create rc={rc}")
    lr = [(out_lr.raw[64 * j : 64 * j + 32], out_lr.raw[64 * j + 32 : 64 * j + 64]) for j in range(lg)]
    return lr, out_ab.raw[:32], out_ab.raw[32:]
