"""ctypes binding of libbpgpu.so (include/bpgpu.h).

The library is built in-tree by `__graft_entry__.build()`; if it is missing or no
CUDA device is present every call fails loudly — there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbpgpu.so")

BPG_OK = 0
BPG_ERR_ARG, BPG_ERR_LEN, BPG_ERR_POW2, BPG_ERR_CAPACITY = -1, -2, -3, -4
BPG_ERR_DECODE, BPG_ERR_VERIFY, BPG_ERR_CUDA, BPG_ERR_NOMEM = -5, -6, -7, -8


class BpgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbpgpu error {code}: {msg}")
        self.code = code


_lib = None

_P = ctypes.c_void_p
_SZ = ctypes.c_size_t
_I = ctypes.c_int

_U64 = ctypes.c_uint64
_CP = ctypes.c_char_p


class Term(ctypes.Structure):
    _fields_ = [("var", ctypes.c_uint64), ("coeff", ctypes.c_uint8 * 32)]


RANDOMIZED_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)

_SIGNATURES = {
    "bpg_transcript_new": (_P, [_P, _SZ]),
    "bpg_transcript_clone": (_P, [_P]),
    "bpg_transcript_free": (None, [_P]),
    "bpg_transcript_append_message": (None, [_P, _CP, _P, _SZ]),
    "bpg_transcript_append_u64": (None, [_P, _CP, _U64]),
    "bpg_transcript_challenge_bytes": (None, [_P, _CP, _P, _SZ]),
    "bpg_transcript_challenge_scalar": (None, [_P, _CP, _P]),
    "bpg_transcript_build_rng": (_P, [_P]),
    "bpg_transcript_rng_rekey_with_witness_bytes": (None, [_P, _CP, _P, _SZ]),
    "bpg_transcript_rng_finalize": (None, [_P, _P]),
    "bpg_transcript_rng_fill_bytes": (None, [_P, _P, _SZ]),
    "bpg_transcript_rng_free": (None, [_P]),
    "bpg_gens_new": (_I, [_P, _P, _P, _SZ, _P, _P, ctypes.POINTER(_P)]),
    "bpg_points_from_uniform": (_I, [_P, _P, _SZ, _P]),
    "bpg_gens_chain": (_I, [_P, _P, _SZ, _SZ, _SZ, _P]),
    "bpg_gens_derive": (_I, [_P, _SZ, ctypes.c_uint32, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_gens_free": (None, [_P]),
    "bpg_gens_capacity": (_SZ, [_P]),
    "bpg_gens_table": (_P, [_P]),
    "bpg_pedersen_commit": (_I, [_P, _P, _P, _P, _SZ, _P]),
    "bpg_ipp_create": (_I, [_P, _P, _P, _P, _P, _P, _SZ, _P, _SZ, _SZ, _P, _P, _P, _SZ, ctypes.POINTER(_SZ)]),
    "bpg_ipp_verify": (_I, [_P, _P, _SZ, _P, _P, _P, _P, _P, _SZ, _P, _SZ, _P, _SZ]),
    "bpg_var_one": (_U64, []),
    "bpg_prover_new": (_I, [_P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_verifier_new": (_I, [_P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_cs_free": (None, [_P]),
    "bpg_prover_commit": (_I, [_P, _P, _P, _P, ctypes.POINTER(_U64)]),
    "bpg_prover_commit_batch": (_I, [_P, _P, _P, _SZ, _P, _P]),
    "bpg_verifier_commit": (_I, [_P, _P, ctypes.POINTER(_U64)]),
    "bpg_cs_commit_public": (_I, [_P, _P, ctypes.POINTER(_U64)]),
    "bpg_cs_multiply": (_I, [_P, ctypes.POINTER(Term), _SZ, ctypes.POINTER(Term), _SZ, ctypes.POINTER(_U64)]),
    "bpg_cs_allocate": (_I, [_P, _P, ctypes.POINTER(_U64)]),
    "bpg_cs_allocate_multiplier": (_I, [_P, _P, _P, ctypes.POINTER(_U64)]),
    "bpg_cs_constrain": (_I, [_P, ctypes.POINTER(Term), _SZ]),
    "bpg_cs_specify_randomized_constraints": (_I, [_P, RANDOMIZED_CB, _P]),
    "bpg_cs_challenge_scalar": (_I, [_P, _CP, _P]),
    "bpg_cs_eval": (_I, [_P, ctypes.POINTER(Term), _SZ, _P]),
    "bpg_cs_num_multipliers": (_SZ, [_P]),
    "bpg_gadget_random_circuit": (_I, [_P, ctypes.c_uint64, _SZ, _SZ, _P]),
    "bpg_gadget_square_chain": (_I, [_P, ctypes.c_uint64, _SZ, ctypes.POINTER(ctypes.c_uint64)]),
    "bpg_gadget_shuffle": (_I, [_P, _P, _P, _SZ]),
    "bpg_cs_num_constraints": (_SZ, [_P]),
    "bpg_prover_prove": (_I, [_P, _P, _SZ, ctypes.POINTER(_SZ)]),
    "bpg_prover_prove_with_rng_bytes": (_I, [_P, _P, _P, _SZ, ctypes.POINTER(_SZ)]),
    "bpg_prover_prove_deterministic": (_I, [_P, _U64, _P, _SZ, ctypes.POINTER(_SZ)]),
    "bpg_verifier_verify": (_I, [_P, _P, _SZ]),
    "bpg_verifier_verify_with_rng_bytes": (_I, [_P, _P, _SZ, _P]),
    "bpg_batch_verify": (_I, [_P, _P, _P, _SZ, _P]),
    "bpg_init": (_I, [_I, ctypes.POINTER(_P)]),
    "bpg_free": (None, [_P]),
    "bpg_set_stream": (_I, [_P, _P, _I]),
    "bpg_profile_enable": (_I, [_P, _I]),
    "bpg_profile_reset": (_I, [_P]),
    "bpg_profile_read": (_I, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64), _I]),
    "bpg_profile_phase_name": (ctypes.c_char_p, [_I]),
    "bpg_sync": (_I, [_P]),
    "bpg_strerror": (ctypes.c_char_p, [_I]),
    "bpg_last_cuda_error": (_I, [_P]),
    "bpg_launch_count": (ctypes.c_uint64, [_P]),
    "bpg_set_window": (_I, [_P, _I]),
    "bpg_set_groups": (_I, [_P, _I]),
    "bpg_table_upload": (_I, [_P, _P, _SZ, ctypes.POINTER(_P)]),
    "bpg_table_upload_dev": (_I, [_P, _P, _SZ, ctypes.POINTER(_P)]),
    "bpg_table_len": (_SZ, [_P]),
    "bpg_table_entry_bytes": (_SZ, [_P]),
    "bpg_table_set_windows": (_I, [_P, _P, _I]),
    "bpg_table_window": (_I, [_P]),
    "bpg_table_build_comb": (_I, [_P, _P]),
    "bpg_table_has_comb": (_I, [_P]),
    "bpg_table_free": (None, [_P]),
    "bpg_msm": (_I, [_P, _P, _P, _SZ, _P]),
    "bpg_msm_table": (_I, [_P, _P, _SZ, _SZ, _P, _I, _P]),
    "bpg_msm_table_submit": (_I, [_P, _P, _SZ, _SZ, _P, _I, ctypes.POINTER(_P)]),
    "bpg_msm_job_wait": (_I, [_P, _P]),
    "bpg_msm_table_indexed": (_I, [_P, _P, _P, _P, _P, _SZ, _I, _P]),
    "bpg_msm_mixed": (_I, [_P, _P, _SZ, _P, _P, _P, _I, _P, _P]),
    "bpg_adhoc_prefetch": (_I, [_P, _P, _SZ]),
    "bpg_dev_msm_table": (_I, [_P, _P, _SZ, _SZ, _P, _I, _P]),
    "bpg_dev_sum_encode": (_I, [_P, _P, _I, _I, _P, _P]),
    "bpg_msm_table_partial": (_I, [_P, _P, _SZ, _SZ, _P, _I, _P]),
    "bpg_sum_encode": (_I, [_P, _P, _I, _I, _P]),
    "bpg_peer_create": (_I, [_P, _I, _I, _I, ctypes.POINTER(_P), _P]),
    "bpg_peer_connect": (_I, [_P, _P]),
    "bpg_dev_exchange_sum_encode": (_I, [_P, _P, _P, _I, _P, _P]),
    "bpg_peer_status": (_I, [_P, ctypes.POINTER(_I)]),
    "bpg_peer_free": (None, [_P]),
    "bpg_ipp_begin": (_I, [_P, _P, _SZ, _P, _SZ, _SZ, _P, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_ipp_begin_dev": (_I, [_P, _P, _SZ, _P, _SZ, _SZ, _P, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_ipp_begin_shared": (_I, [_P, _P, _SZ, _SZ, _SZ, _P, _SZ, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_ipp_rounds_left": (_SZ, [_P]),
    "bpg_ipp_round_LR": (_I, [_P, _P, _P]),
    "bpg_ipp_round_fold": (_I, [_P, _P, _P]),
    "bpg_ipp_finish": (_I, [_P, _P, _P]),
    "bpg_ipp_free": (None, [_P]),
    "bpg_ipp_begin_shares": (_I, [_P, _P, _SZ, _SZ, _SZ, _P, _SZ, _I, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_ipp_lanes": (_I, [_P]),
    "bpg_ipp_len": (_SZ, [_P]),
    "bpg_ipp_read_ab": (_I, [_P, _P, _P]),
    "bpg_ipp_round_LR_shares": (_I, [_P, _P, _P, _P, _P]),
    "bpg_ipp_finish_shares": (_I, [_P, _P, _P]),
    "bpg_points_sum": (_I, [_P, _P, _I, _I, _P]),
    "bpg_r1cs_dev_new": (_I, [_P, _SZ, ctypes.POINTER(_P)]),
    "bpg_r1cs_dev_free": (None, [_P]),
    "bpg_r1cs_dev_reserve": (_I, [ctypes.POINTER(_P), _SZ]),
    "bpg_ipp_verify_msm": (_I, [_P, _P, _SZ, _P, _SZ, _P, _P, _SZ, _P, _P, _P, _P]),
    "bpg_r1cs_dev_commit": (_I, [_P, _P, _SZ, _SZ, _SZ, _SZ, _SZ, _P, _P, _P, ctypes.c_uint64, _P, _P]),
    "bpg_r1cs_dev_commit_keyed": (_I, [_P, _P, _SZ, _SZ, _SZ, _SZ, _SZ, _P, _P, _P, _P, _P, _P]),
    "bpg_host_alloc": (_P, [_SZ]),
    "bpg_host_free": (None, [_P]),
    "bpg_r1cs_dev_flatten": (_I, [_P, _SZ, _SZ, _SZ, _P, _P, _P, _P, _P]),
    "bpg_vbatch_new": (_I, [_P, _SZ, _SZ, _P]),
    "bpg_vbatch_free": (None, [_P]),
    "bpg_vbatch_put": (_I, [_P, _SZ, _P, _P, _P]),
    "bpg_vbatch_check": (_I, [_P, _P, _SZ, _SZ, _SZ, _P, _SZ, _P, _P, _P, _SZ, _P]),
    "bpg_r1cs_dev_flatten_terms": (_I, [_P, _SZ, _SZ, _P, _P, _P]),
    "bpg_r1cs_terms_prefetch": (_I, [_P, _P, _I]),
    "bpg_r1cs_terms_wait": (_I, [_P]),
    "bpg_r1cs_dev_poly_t": (_I, [_P, _SZ, _P, _P, _P]),
    "bpg_r1cs_dev_verify_msm": (_I, [_P, _P, _SZ, _SZ, _SZ, _P, _P, _SZ, _P, _P, _P]),
    "bpg_r1cs_dev_ipp_begin": (_I, [_P, _P, _SZ, _SZ, _SZ, _P, _SZ, _SZ, _SZ, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_stark_table_upload": (_I, [_P, _P, _SZ, ctypes.POINTER(_P)]),
    "bpg_stark_table_set_windows": (_I, [_P, _P, _I]),
    "bpg_stark_table_window": (_I, [_P]),
    "bpg_stark_table_len": (_SZ, [_P]),
    "bpg_stark_table_free": (None, [_P]),
    "bpg_stark_msm_table": (_I, [_P, _P, _SZ, _SZ, _P, _I, _P]),
    "bpg_stark_msm": (_I, [_P, _P, _P, _SZ, _P]),
    "bpg_keccak256": (None, [_P, _SZ, _P]),
    "bpg_stark_hash_to_scalar": (None, [_P, _P]),
    "bpg_stark_gens_chain": (_I, [_P, _P, _SZ, _SZ, _P]),
    "bpg_stark_wide_mul_generator": (_I, [_P, _P, _SZ, _P]),
    "bpg_stark_ipp_begin": (_I, [_P, _P, _SZ, _P, _SZ, _SZ, _P, _P, _P, _P, _P, ctypes.POINTER(_P)]),
    "bpg_stark_ipp_rounds_left": (_SZ, [_P]),
    "bpg_stark_ipp_round_LR": (_I, [_P, _P, _P]),
    "bpg_stark_ipp_round_fold": (_I, [_P, _P, _P]),
    "bpg_stark_ipp_finish": (_I, [_P, _P, _P]),
    "bpg_stark_ipp_free": (None, [_P]),
    "bpg_comb_create": (_I, [_P, _P, _I, ctypes.POINTER(_P)]),
    "bpg_comb_free": (None, [_P]),
    "bpg_comb_mul": (_I, [_P, _P, _P, _SZ, _P]),
    "bpg_dev_comb_mul": (_I, [_P, _P, _P, _SZ, _P, _P]),
}


def exported_symbols():
    return list(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no fallback implementation."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int):
    if code != BPG_OK:
        raise BpgError(code, lib().bpg_strerror(code).decode())
