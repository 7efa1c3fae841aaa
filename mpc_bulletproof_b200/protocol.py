"""Python face of the host mirror of the reference's protocol API (the C++ host code in
csrc/host/ bound through include/bpgpu.h): `Transcript`, `Gens` (PedersenGens +
BulletproofGens party 0 as resident tables), `InnerProductProof.create/verify`,
`Prover`, `Verifier` — same method names and argument meaning as the reference
(src/transcript.rs, src/generators.rs, src/inner_product_proof.rs,
src/r1cs/{prover,verifier}.rs), so parity tests read like the reference's tests.

Scalars are Python ints mod l, points are 32-byte compressed encodings.
Variables are tuples ("L", i) | ("R", i) | ("O", i) | ("V", j) | ("1",); a linear
combination is any object with a `.terms` dict {variable: coefficient}, a bare variable,
or an int (a constant).
"""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import BpgError, Term, check, lib
from .api import Context, Table

L = 2**252 + 27742317777372353535851937790883648493

_KIND = {"L": 1, "R": 2, "O": 3, "V": 4, "1": 5, "0": 6}
_KIND_INV = {v: k for k, v in _KIND.items()}


class VerificationError(Exception):
    """ProofError::VerificationError / R1CSError::VerificationError"""


class FormatError(Exception):
    """ProofError::FormatError / R1CSError::FormatError"""


class InvalidGeneratorsLength(Exception):
    """R1CSError::InvalidGeneratorsLength"""


def _raise(code: int):
    if code == _lib.BPG_OK:
        return
    if code == _lib.BPG_ERR_VERIFY:
        raise VerificationError()
    if code == _lib.BPG_ERR_DECODE:
        raise FormatError()
    if code == _lib.BPG_ERR_CAPACITY:
        raise InvalidGeneratorsLength()
    check(code)


def sc_bytes(x: int) -> bytes:
    return (x % L).to_bytes(32, "little")


def _scs(xs) -> bytes:
    """ints mod l -> canonical bytes; a bytes object is taken as already serialised scalars"""
    if isinstance(xs, (bytes, bytearray)):
        return bytes(xs)
    return b"".join(sc_bytes(x) for x in xs)


def _count(xs) -> int:
    return len(xs) // 32 if isinstance(xs, (bytes, bytearray)) else len(xs)


class Transcript:
    def __init__(self, label: bytes, _h=None):
        self._h = _h if _h is not None else ctypes.c_void_p(lib().bpg_transcript_new(label, len(label)))

    def clone(self) -> "Transcript":
        return Transcript(b"", ctypes.c_void_p(lib().bpg_transcript_clone(self._h)))

    def __del__(self):
        try:
            lib().bpg_transcript_free(self._h)
        except Exception:
            pass

    def append_message(self, label: bytes, msg: bytes):
        lib().bpg_transcript_append_message(self._h, label, msg, len(msg))

    def append_u64(self, label: bytes, x: int):
        lib().bpg_transcript_append_u64(self._h, label, x)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(n)
        lib().bpg_transcript_challenge_bytes(self._h, label, out, n)
        return out.raw

    def challenge_scalar(self, label: bytes) -> int:
        out = ctypes.create_string_buffer(32)
        lib().bpg_transcript_challenge_scalar(self._h, label, out)
        return int.from_bytes(out.raw, "little")

    def build_rng(self) -> "TranscriptRng":
        return TranscriptRng(ctypes.c_void_p(lib().bpg_transcript_build_rng(self._h)))


class TranscriptRng:
    """merlin::TranscriptRngBuilder / TranscriptRng over the host mirror's transcript."""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        try:
            lib().bpg_transcript_rng_free(self._h)
        except Exception:
            pass

    def rekey_with_witness_bytes(self, label: bytes, witness: bytes) -> "TranscriptRng":
        lib().bpg_transcript_rng_rekey_with_witness_bytes(self._h, label, witness, len(witness))
        return self

    def finalize(self, random_bytes: bytes) -> "TranscriptRng":
        assert len(random_bytes) == 32
        lib().bpg_transcript_rng_finalize(self._h, random_bytes)
        return self

    def fill_bytes(self, n: int) -> bytes:
        out = ctypes.create_string_buffer(n)
        lib().bpg_transcript_rng_fill_bytes(self._h, out, n)
        return out.raw


class Gens:
    """PedersenGens {B, B_blinding} and BulletproofGens (party 0) resident in HBM."""

    def __init__(self, ctx: Context, G: bytes, H: bytes, B: bytes, B_blinding: bytes):
        assert len(G) == len(H) and len(G) % 32 == 0
        self.ctx = ctx
        self.gens_capacity = len(G) // 32
        self.B, self.B_blinding = B, B_blinding
        self._h = ctypes.c_void_p()
        _raise(lib().bpg_gens_new(ctx._h, G, H, self.gens_capacity, B, B_blinding, ctypes.byref(self._h)))
        ctx._children.add(self)

    @classmethod
    def derive(cls, ctx: Context, gens_capacity: int, party: int = 0) -> "Gens":
        """BulletproofGens::new(gens_capacity, ..).share(party) + PedersenGens::default()
        (reference src/generators.rs:61-71, 182-235), derived on the device; `.G`, `.H` hold the
        compressed generators."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.gens_capacity = gens_capacity
        G = ctypes.create_string_buffer(32 * max(gens_capacity, 1))
        H = ctypes.create_string_buffer(32 * max(gens_capacity, 1))
        B, Bb = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        self._h = ctypes.c_void_p()
        _raise(lib().bpg_gens_derive(ctx._h, gens_capacity, party, G, H, B, Bb, ctypes.byref(self._h)))
        self.G, self.H = G.raw[: 32 * gens_capacity], H.raw[: 32 * gens_capacity]
        self.B, self.B_blinding = B.raw, Bb.raw
        ctx._children.add(self)
        return self

    def close(self):
        if self._h:
            lib().bpg_gens_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def table(self) -> Table:
        """The resident windowed table [G (capacity) | H (capacity) | B | B_blinding] (borrowed)."""
        t = Table.__new__(Table)
        t.ctx = self.ctx
        t._h = ctypes.c_void_p(lib().bpg_gens_table(self._h))
        t.close = lambda: None  # owned by the Gens
        return t

    def commit(self, value: int, blinding: int) -> bytes:
        """PedersenGens::commit (reference src/generators.rs:41-43)"""
        return self.commit_batch([value], [blinding])[0]

    def commit_batch(self, values, blindings):
        n = len(values)
        out = ctypes.create_string_buffer(32 * max(n, 1))
        _raise(lib().bpg_pedersen_commit(self.ctx._h, self._h, _scs(values), _scs(blindings), n, out))
        return [out.raw[32 * i : 32 * i + 32] for i in range(n)]


class InnerProductProof:
    """Proof bytes are (L_j R_j)_j || a || b (reference src/inner_product_proof.rs:388-397)."""

    def __init__(self, data: bytes):
        self.data = data

    def to_bytes(self) -> bytes:
        return self.data

    @staticmethod
    def from_bytes(b: bytes) -> "InnerProductProof":
        return InnerProductProof(bytes(b))

    @property
    def L_vec(self):
        k = (len(self.data) // 32 - 2) // 2
        return [self.data[64 * i : 64 * i + 32] for i in range(k)]

    @property
    def R_vec(self):
        k = (len(self.data) // 32 - 2) // 2
        return [self.data[64 * i + 32 : 64 * i + 64] for i in range(k)]

    @property
    def a(self) -> int:
        return int.from_bytes(self.data[-64:-32], "little")

    @property
    def b(self) -> int:
        return int.from_bytes(self.data[-32:], "little")

    @staticmethod
    def create(ctx: Context, transcript: Transcript, Q: bytes, G_factors, H_factors, G_vec: Table, H_vec: Table, a_vec, b_vec,
               g_off: int = 0, h_off: int = 0) -> "InnerProductProof":
        """reference src/inner_product_proof.rs:49-58 — G_vec/H_vec are resident tables."""
        n = _count(a_vec)
        if not (_count(b_vec) == n and _count(G_factors) == n and _count(H_factors) == n):
            raise BpgError(_lib.BPG_ERR_LEN, "vector lengths differ")
        cap = 64 * 32 + 64
        out = ctypes.create_string_buffer(cap)
        ln = ctypes.c_size_t()
        _raise(
            lib().bpg_ipp_create(
                ctx._h, transcript._h, Q, _scs(G_factors), _scs(H_factors), G_vec._h, g_off, H_vec._h, h_off, n,
                _scs(a_vec), _scs(b_vec), out, cap, ctypes.byref(ln),
            )
        )
        return InnerProductProof(out.raw[: ln.value])

    def verify(self, ctx: Context, n: int, transcript: Transcript, G_factors, H_factors, P: bytes, Q: bytes, G: Table, H: Table,
               g_off: int = 0, h_off: int = 0):
        """reference src/inner_product_proof.rs:317-372; raises VerificationError / FormatError."""
        _raise(
            lib().bpg_ipp_verify(
                ctx._h, transcript._h, n, _scs(G_factors), _scs(H_factors), P, Q, G._h, g_off, H._h, h_off, self.data, len(self.data)
            )
        )


def _terms(lc) -> "ctypes.Array":
    if isinstance(lc, tuple):
        items = [(lc, 1)]
    elif isinstance(lc, int):
        items = [(("1",), lc)]
    else:
        items = list(lc.terms.items())
    arr = (Term * max(len(items), 1))()
    for k, (var, c) in enumerate(items):
        idx = var[1] if len(var) > 1 else 0
        arr[k].var = (_KIND[var[0]] << 56) | idx
        arr[k].coeff[:] = sc_bytes(c)
    return arr, len(items)


def _var(h: int):
    kind = _KIND_INV[h >> 56]
    return (kind, h & ((1 << 56) - 1)) if kind in "LROV" else (kind,)


class _CS:
    """ConstraintSystem / RandomizableConstraintSystem (reference src/r1cs/constraint_system.rs:55-208)."""

    def __init__(self):
        self._cbs = []  # keep ctypes callbacks alive

    def __del__(self):
        try:
            lib().bpg_cs_free(self._h)
        except Exception:
            pass

    def multiply(self, left, right):
        la, nl = _terms(left)
        ra, nr = _terms(right)
        out = (ctypes.c_uint64 * 3)()
        _raise(lib().bpg_cs_multiply(self._h, la, nl, ra, nr, out))
        return _var(out[0]), _var(out[1]), _var(out[2])

    def allocate(self, assignment=None):
        out = ctypes.c_uint64()
        _raise(lib().bpg_cs_allocate(self._h, None if assignment is None else sc_bytes(assignment), ctypes.byref(out)))
        return _var(out.value)

    def allocate_multiplier(self, l=None, r=None):
        out = (ctypes.c_uint64 * 3)()
        _raise(
            lib().bpg_cs_allocate_multiplier(self._h, None if l is None else sc_bytes(l), None if r is None else sc_bytes(r), out)
        )
        return _var(out[0]), _var(out[1]), _var(out[2])

    def square_chain(self, var, n: int):
        """The reference's DummyCircuit (benches/r1cs.rs:24-32): n chained squarings of `var`."""
        out = ctypes.c_uint64()
        idx = var[1] if len(var) > 1 else 0
        _raise(lib().bpg_gadget_square_chain(self._h, (_KIND[var[0]] << 56) | idx, n, ctypes.byref(out)))
        return _var(out.value)

    def shuffle(self, x, y):
        """The shuffle gadget of the reference's bench and tests (benches/shuffle.rs:30-69): y is a permutation
        of x.  Native (bpg_gadget_shuffle): one library call instead of 2 (k - 1) multiplier calls in a callback."""
        assert len(x) == len(y) and len(x) >= 1
        k = len(x)

        def pack(vs):
            return (ctypes.c_uint64 * k)(*[(_KIND[v[0]] << 56) | (v[1] if len(v) > 1 else 0) for v in vs])

        _raise(lib().bpg_gadget_shuffle(self._h, pack(x), pack(y), k))

    def random_circuit(self, seed: int, n_mult: int, n_cons: int, c0: bytes | None = None) -> bytes:
        """BASELINE.json config 4: the synthetic random circuit over the variables committed so far.
        Prover: returns the public constants c0 (n_cons x 32 bytes); verifier: pass them in."""
        buf = ctypes.create_string_buffer(c0 if c0 is not None else b"", 32 * n_cons)
        _raise(lib().bpg_gadget_random_circuit(self._h, seed, n_mult, n_cons, buf))
        return buf.raw

    def constrain(self, lc):
        arr, n = _terms(lc)
        _raise(lib().bpg_cs_constrain(self._h, arr, n))

    def commit_public(self, value: int):
        out = ctypes.c_uint64()
        _raise(lib().bpg_cs_commit_public(self._h, sc_bytes(value), ctypes.byref(out)))
        return _var(out.value)

    def specify_randomized_constraints(self, callback):
        err = []

        def tramp(_cs, _user):
            try:
                callback(self)
                return 0
            except Exception as e:  # surfaced after prove/verify returns
                err.append(e)
                return _lib.BPG_ERR_ARG

        cb = _lib.RANDOMIZED_CB(tramp)
        self._cbs.append((cb, err))
        _raise(lib().bpg_cs_specify_randomized_constraints(self._h, cb, None))

    def challenge_scalar(self, label: bytes) -> int:
        out = ctypes.create_string_buffer(32)
        _raise(lib().bpg_cs_challenge_scalar(self._h, label, out))
        return int.from_bytes(out.raw, "little")

    def eval(self, lc) -> int:
        arr, n = _terms(lc)
        out = ctypes.create_string_buffer(32)
        _raise(lib().bpg_cs_eval(self._h, arr, n, out))
        return int.from_bytes(out.raw, "little")

    def num_multipliers(self) -> int:
        return int(lib().bpg_cs_num_multipliers(self._h))

    def num_constraints(self) -> int:
        return int(lib().bpg_cs_num_constraints(self._h))

    def _reraise(self):
        for _, err in self._cbs:
            if err:
                raise err[0]


class Prover(_CS):
    """reference src/r1cs/prover.rs: Prover::new / commit / prove"""

    def __init__(self, gens: Gens, transcript: Transcript):
        super().__init__()
        self.gens, self.transcript = gens, transcript
        self._h = ctypes.c_void_p()
        _raise(lib().bpg_prover_new(gens.ctx._h, gens._h, transcript._h, ctypes.byref(self._h)))

    def commit(self, v: int, v_blinding: int):
        V = ctypes.create_string_buffer(32)
        var = ctypes.c_uint64()
        _raise(lib().bpg_prover_commit(self._h, sc_bytes(v), sc_bytes(v_blinding), V, ctypes.byref(var)))
        return V.raw, _var(var.value)

    def commit_batch(self, values, blindings):
        """len(values) calls of `commit` in one library call (one batched fixed-base launch)"""
        n = len(values)
        V = ctypes.create_string_buffer(32 * max(n, 1))
        vs = (ctypes.c_uint64 * max(n, 1))()
        _raise(lib().bpg_prover_commit_batch(self._h, _scs(values), _scs(blindings), n, V, vs))
        return [(V.raw[32 * i : 32 * i + 32], _var(vs[i])) for i in range(n)]

    def prove(self, rng_seed: int | None = None, rng_bytes: bytes | None = None) -> bytes:
        """Prover::prove(bp_gens) -> R1CSProof bytes (reference src/r1cs/proof.rs:82-108).
        No argument: blindings from the transcript-bound RNG finalized with operating-system randomness
        (bpg_prover_prove, the production entry point).  rng_bytes (32): the same with the caller's
        bytes in place of the OS draw (reproducible).  rng_seed: TEST ONLY xoshiro256** blindings."""
        cap = 1 + 14 * 32 + 66 * 32
        out = ctypes.create_string_buffer(cap)
        ln = ctypes.c_size_t()
        if rng_seed is not None:
            code = lib().bpg_prover_prove_deterministic(self._h, rng_seed, out, cap, ctypes.byref(ln))
        elif rng_bytes is not None:
            assert len(rng_bytes) == 32
            code = lib().bpg_prover_prove_with_rng_bytes(self._h, rng_bytes, out, cap, ctypes.byref(ln))
        else:
            code = lib().bpg_prover_prove(self._h, out, cap, ctypes.byref(ln))
        self._reraise()
        _raise(code)
        return out.raw[: ln.value]


class Verifier(_CS):
    """reference src/r1cs/verifier.rs: Verifier::new / commit / verify"""

    def __init__(self, gens: Gens, transcript: Transcript):
        super().__init__()
        self.gens, self.transcript = gens, transcript
        self._h = ctypes.c_void_p()
        _raise(lib().bpg_verifier_new(gens.ctx._h, gens._h, transcript._h, ctypes.byref(self._h)))

    def commit(self, V: bytes):
        var = ctypes.c_uint64()
        _raise(lib().bpg_verifier_commit(self._h, V, ctypes.byref(var)))
        return _var(var.value)

    def verify(self, proof: bytes, hardened: bool = False, rng_bytes: bytes | None = None):
        """Verifier::verify.  Default: r = challenge_scalar("r") as the mounted fork (verifier.rs:506).
        hardened / rng_bytes: r from a TranscriptRng finalized with 32 bytes unknown to the prover."""
        if hardened or rng_bytes is not None:
            code = lib().bpg_verifier_verify_with_rng_bytes(self._h, proof, len(proof), rng_bytes)
        else:
            code = lib().bpg_verifier_verify(self._h, proof, len(proof))
        self._reraise()
        _raise(code)


def batch_verify(jobs) -> list[bool]:
    """jobs: [(Verifier with its constraint system built, proof bytes), ...] -> per-proof accept, the answers of
    the reference's proof-by-proof `Verifier::verify` (verifier.rs:393).  Proofs of equal padded size on one
    context share ONE multiscalar multiplication (random combination of their checks); a combination that does
    not pass is split until the failing proofs are isolated.  Across GPUs shard the jobs with
    `multi.batch_verify_sharded`."""
    n = len(jobs)
    if n == 0:
        return []
    hs = (ctypes.c_void_p * n)(*[v._h for v, _ in jobs])
    bufs = [ctypes.create_string_buffer(p, len(p)) for _, p in jobs]
    ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    lens = (ctypes.c_size_t * n)(*[len(p) for _, p in jobs])
    ok = ctypes.create_string_buffer(n)
    code = lib().bpg_batch_verify(hs, ptrs, lens, n, ok)
    for v, _ in jobs:
        v._reraise()
    _raise(code)
    return [b != 0 for b in ok.raw]


def points_from_uniform(ctx: Context, uniform: bytes) -> bytes:
    """from_uniform_bytes of every 64-byte block (RFC 9496 4.3.4), on the device."""
    n = len(uniform) // 64
    out = ctypes.create_string_buffer(32 * max(n, 1))
    _raise(lib().bpg_points_from_uniform(ctx._h, uniform, n, out))
    return out.raw[: 32 * n]


def gens_chain(ctx: Context, label: bytes, skip: int, n: int) -> bytes:
    """Points [skip, skip+n) of GeneratorsChain::new(label) (reference src/generators.rs:80-125)."""
    out = ctypes.create_string_buffer(32 * max(n, 1))
    _raise(lib().bpg_gens_chain(ctx._h, label, len(label), skip, n, out))
    return out.raw[: 32 * n]
