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
