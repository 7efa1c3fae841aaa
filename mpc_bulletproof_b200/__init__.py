"""mpc_bulletproof_b200 — B200 (sm_100a) engine for the multiscalar
multiplications and inner-product-argument folding behind
renegade-fi/mpc-bulletproof's R1CS prover/verifier (ristretto255 instantiation).

The product is the C-ABI shared library `libbpgpu.so` (include/bpgpu.h); this
package is the thin Python host mirror used by the tests and bench.
"""
from .api import Comb, Context, Table, msm  # noqa: F401
from ._lib import BpgError  # noqa: F401
