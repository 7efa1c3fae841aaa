"""ORACLE (test infrastructure).  Restatement of the reference's protocol layer —
what MSMs are formed from which scalars and points, in which order, what enters
the transcript and how proofs serialise — over the ristretto255 group layer of
oracle/group.py with the conventions of SURVEY.md Appendix A.6 (upstream
dalek-bulletproofs conventions for encodings, generators and Merlin).

Every function cites the reference lines it follows.  Proving randomness is an
explicit input (`Blindings`, drawn in the reference's RNG order) so that proofs are
reproducible; the reference itself draws it from `thread_rng()`
(src/r1cs/prover.rs:435-445) and holds no golden proof.
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass, field

from . import group as G
from .merlin import Transcript as Merlin

L = G.L


# ---------------------------------------------------------------- randomness
class Xoshiro256ss:
    """xoshiro256** seeded through SplitMix64 (SURVEY.md §8d synthetic inputs)."""

    M = (1 << 64) - 1

    def __init__(self, seed: int):
        s = seed & self.M
        self.s = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & self.M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
            self.s.append(z ^ (z >> 31))

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & Xoshiro256ss.M

    def next_u64(self) -> int:
        s = self.s
        result = (self._rotl((s[1] * 5) & self.M, 7) * 9) & self.M
        t = (s[1] << 17) & self.M
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return result

    def scalar(self) -> int:
        """64 PRNG bytes, little-endian, reduced mod l (uniform, like Scalar::random)."""
        b = b"".join(self.next_u64().to_bytes(8, "little") for _ in range(8))
        return int.from_bytes(b, "little") % L


class Blindings:
    """Source of the prover's random scalars, consumed in the reference's draw order
    (src/r1cs/prover.rs:457-462, 519-530, 621-625)."""

    def __init__(self, seed: int):
        self.rng = Xoshiro256ss(seed)

    def scalar(self) -> int:
        return self.rng.scalar()

    def vector_pair(self, n: int):
        """The two blinding vectors of a phase (s_L, s_R; src/r1cs/prover.rs:460-461, 523-528).
        One xoshiro draw keys a SplitMix64 stream whose consecutive 64-byte blocks are
        s_L[0], s_R[0], s_L[1], s_R[1], ... (each block reduced mod l like `scalar`), so that
        element j is a pure function of (key, j) and can be produced where it is consumed.
        No draw for n = 0."""
        if n == 0:
            return [], []
        M = (1 << 64) - 1
        key = self.rng.next_u64()

        def word(counter):
            z = (key + 0x9E3779B97F4A7C15 * (counter + 1)) & M
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            return z ^ (z >> 31)

        def block(idx):
            b = b"".join(word(8 * idx + k).to_bytes(8, "little") for k in range(8))
            return int.from_bytes(b, "little") % L

        return [block(2 * j) for j in range(n)], [block(2 * j + 1) for j in range(n)]


def chacha20_block(key: bytes, counter: int, nonce: bytes) -> bytes:
    """RFC 8439 section 2.3 (pinned by its 2.3.2 test vector in tests/test_oracle_group.py)."""
    M = 0xFFFFFFFF

    def rotl(v, c):
        return ((v << c) & M) | (v >> (32 - c))

    def qr(x, a, b, c, d):
        x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 7)

    init = list(struct.unpack("<4I", b"expand 32-byte k")) + list(struct.unpack("<8I", key)) + [counter & M] + list(struct.unpack("<3I", nonce))
    x = list(init)
    for _ in range(10):
        qr(x, 0, 4, 8, 12); qr(x, 1, 5, 9, 13); qr(x, 2, 6, 10, 14); qr(x, 3, 7, 11, 15)
        qr(x, 0, 5, 10, 15); qr(x, 1, 6, 11, 12); qr(x, 2, 7, 8, 13); qr(x, 3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & M for a, b in zip(x, init)])


BLIND_NONCE = b"bpg sLsR v01"


class TranscriptBlindings:
    """The production blinding source: merlin's TranscriptRng forked from the prover's transcript after
    "m", rekeyed with every v_blinding and finalized with 32 external random bytes (reference
    src/r1cs/prover.rs:435-445), drawn in the reference's order.  `Prover.prove` calls `bind` at the fork
    point.  The two blinding vectors of a phase come from ONE 32-byte draw that keys ChaCha20: block
    2j -> s_L[j], block 2j+1 -> s_R[j], each 64-byte block reduced mod l (the product expands them on
    the device instead of drawing 2n scalars from the sequential RNG)."""

    def __init__(self, random_bytes: bytes):
        assert len(random_bytes) == 32
        self.random_bytes = random_bytes
        self.rng = None

    def bind(self, transcript, v_blinding):
        rng = transcript.build_rng()
        for vb in v_blinding:
            rng.rekey_with_witness_bytes(b"v_blinding", G.sc_to_bytes(vb))
        self.rng = rng.finalize(self.random_bytes)

    def scalar(self) -> int:
        return int.from_bytes(self.rng.fill_bytes(64), "little") % L

    def vector_pair(self, n: int):
        if n == 0:
            return [], []
        key = self.rng.fill_bytes(32)

        def block(idx):
            return int.from_bytes(chacha20_block(key, idx, BLIND_NONCE), "little") % L

        return [block(2 * j) for j in range(n)], [block(2 * j + 1) for j in range(n)]


# ---------------------------------------------------------------- generators
@dataclass
class PedersenGens:
    """reference src/generators.rs:32-71 (B, B_blinding; commit = v*B + r*B_blinding).
    ristretto255 instantiation: B = basepoint, B_blinding = hash-to-group(SHA3-512(B))."""

    B: G.Point = field(default_factory=lambda: G.BASEPOINT)
    B_blinding: G.Point = field(
        default_factory=lambda: G.from_uniform_bytes(hashlib.sha3_512(G.BASEPOINT.encode()).digest())
    )

    def commit(self, value: int, blinding: int) -> G.Point:
        return G.msm_naive([value, blinding], [self.B, self.B_blinding])


class GeneratorsChain:
    """reference src/generators.rs:76-129 with the ristretto255 construction:
    SHAKE256("GeneratorsChain" || label) XOF, 64 bytes per point -> from_uniform_bytes."""

    def __init__(self, label: bytes):
        self.xof = hashlib.shake_256(b"GeneratorsChain" + label)
        self.pos = 0

    def take(self, skip: int, n: int):
        data = self.xof.digest(64 * (skip + n))
        return [G.from_uniform_bytes(data[64 * i : 64 * i + 64]) for i in range(skip, skip + n)]


class BulletproofGens:
    """reference src/generators.rs:158-235: per-party chains labelled 'G'/'H' || u32le(party)."""

    def __init__(self, gens_capacity: int, party_capacity: int = 1):
        self.gens_capacity = 0
        self.party_capacity = party_capacity
        self.G_vec = [[] for _ in range(party_capacity)]
        self.H_vec = [[] for _ in range(party_capacity)]
        self.increase_capacity(gens_capacity)

    def increase_capacity(self, new_capacity: int):
        if self.gens_capacity >= new_capacity:
            return
        for i in range(self.party_capacity):
            for tag, vec in ((b"G", self.G_vec), (b"H", self.H_vec)):
                label = tag + i.to_bytes(4, "little")
                vec[i].extend(GeneratorsChain(label).take(self.gens_capacity, new_capacity - self.gens_capacity))
        self.gens_capacity = new_capacity

    def G(self, n: int, party: int = 0):
        return self.G_vec[party][:n]

    def H(self, n: int, party: int = 0):
        return self.H_vec[party][:n]


# ---------------------------------------------------------------- transcript protocol
class Transcript(Merlin):
    """reference src/transcript.rs:25-121 (TranscriptProtocol); payload conventions of
    SURVEY.md A.6: points 32-byte compressed, scalars 32-byte LE, challenges = 64
    squeezed bytes reduced mod l."""

    def innerproduct_domain_sep(self, n: int):
        self.append_message(b"dom-sep", b"ipp v1")
        self.append_u64(b"n", n)

    def r1cs_domain_sep(self):
        self.append_message(b"dom-sep", b"r1cs v1")

    def r1cs_1phase_domain_sep(self):
        self.append_message(b"dom-sep", b"r1cs-1phase")

    def r1cs_2phase_domain_sep(self):
        self.append_message(b"dom-sep", b"r1cs-2phase")

    def append_scalar(self, label: bytes, s: int):
        self.append_message(label, G.sc_to_bytes(s))

    def append_point(self, label: bytes, p: G.Point):
        self.append_message(label, p.encode())

    def validate_and_append_point(self, label: bytes, p: G.Point):
        if p.is_identity():
            raise VerificationError("identity point")
        self.append_message(label, p.encode())

    def challenge_scalar(self, label: bytes) -> int:
        return G.sc_from_wide(self.challenge_bytes(label, 64))


class VerificationError(Exception):
    pass


class FormatError(Exception):
    pass


class InvalidGeneratorsLength(Exception):
    pass


# ---------------------------------------------------------------- inner product proof
def inner_product(a, b) -> int:
    """reference src/inner_product_proof.rs:463-472"""
    assert len(a) == len(b)
    return sum(x * y for x, y in zip(a, b)) % L


@dataclass
class InnerProductProof:
    L_vec: list
    R_vec: list
    a: int
    b: int

    # reference src/inner_product_proof.rs:49-193
    @staticmethod
    def create(transcript: Transcript, Q, G_factors, H_factors, G_vec, H_vec, a_vec, b_vec, trace=None):
        n = len(G_vec)
        assert len(H_vec) == len(a_vec) == len(b_vec) == len(G_factors) == len(H_factors) == n
        assert n & (n - 1) == 0 and n > 0
        G_vec, H_vec, a_vec, b_vec = list(G_vec), list(H_vec), list(a_vec), list(b_vec)
        transcript.innerproduct_domain_sep(n)
        L_vec, R_vec = [], []
        first = True
        while n != 1:
            n //= 2
            a_L, a_R = a_vec[:n], a_vec[n:]
            b_L, b_R = b_vec[:n], b_vec[n:]
            G_L, G_R = G_vec[:n], G_vec[n:]
            H_L, H_R = H_vec[:n], H_vec[n:]
            c_L = inner_product(a_L, b_R)
            c_R = inner_product(a_R, b_L)
            if first:  # :80-114 — factors folded into the scalars
                Ls = [a_L[i] * G_factors[n + i] % L for i in range(n)] + [b_R[i] * H_factors[i] % L for i in range(n)] + [c_L]
                Rs = [a_R[i] * G_factors[i] % L for i in range(n)] + [b_L[i] * H_factors[n + i] % L for i in range(n)] + [c_R]
            else:  # :149-172
                Ls = a_L + b_R + [c_L]
                Rs = a_R + b_L + [c_R]
            Lp = G.msm(Ls, G_R + H_L + [Q])
            Rp = G.msm(Rs, G_L + H_R + [Q])
            L_vec.append(Lp)
            R_vec.append(Rp)
            transcript.append_point(b"L", Lp)
            transcript.append_point(b"R", Rp)
            u = transcript.challenge_scalar(b"u")
            u_inv = G.sc_inv(u)
            if trace is not None:
                trace.append((Lp.encode(), Rp.encode(), u))
            if first:  # :125-134 — G_i <- g_i*G_i, H_i <- h_i*H_i before folding
                Gs = [(G_factors[i] % L) * G_vec[i] for i in range(2 * n)]
                Hs = [(H_factors[i] % L) * H_vec[i] for i in range(2 * n)]
                G_L, G_R, H_L, H_R = Gs[:n], Gs[n:], Hs[:n], Hs[n:]
                first = False
            # fold_witness :202-248
            a_vec = [(a_L[i] * u + u_inv * a_R[i]) % L for i in range(n)]
            b_vec = [(b_L[i] * u_inv + u * b_R[i]) % L for i in range(n)]
            G_vec = [G.msm_naive([u_inv, u], [G_L[i], G_R[i]]) for i in range(n)]
            H_vec = [G.msm_naive([u, u_inv], [H_L[i], H_R[i]]) for i in range(n)]
        return InnerProductProof(L_vec, R_vec, a_vec[0], b_vec[0])

    # reference src/inner_product_proof.rs:254-310
    def verification_scalars(self, n: int, transcript: Transcript):
        lg_n = len(self.L_vec)
        if lg_n >= 32 or n != (1 << lg_n):
            raise VerificationError("bad n")
        transcript.innerproduct_domain_sep(n)
        challenges = []
        for Lp, Rp in zip(self.L_vec, self.R_vec):
            transcript.validate_and_append_point(b"L", Lp)
            transcript.validate_and_append_point(b"R", Rp)
            challenges.append(transcript.challenge_scalar(b"u"))
        challenges_inv = [G.sc_inv(c) for c in challenges]
        allinv = 1
        for c in challenges_inv:
            allinv = allinv * c % L
        challenges_sq = [c * c % L for c in challenges]
        challenges_inv_sq = [c * c % L for c in challenges_inv]
        s = [allinv]
        for i in range(1, n):
            lg_i = i.bit_length() - 1
            k = 1 << lg_i
            s.append(s[i - k] * challenges_sq[(lg_n - 1) - lg_i] % L)
        return challenges_sq, challenges_inv_sq, s

    # reference src/inner_product_proof.rs:317-372
    def verify(self, n, transcript, G_factors, H_factors, P, Q, G_vec, H_vec):
        u_sq, u_inv_sq, s = self.verification_scalars(n, transcript)
        gs = [self.a * s[i] % L * G_factors[i] % L for i in range(n)]
        hs = [self.b * s[n - 1 - i] % L * H_factors[i] % L for i in range(n)]
        scalars = [self.a * self.b % L] + gs + hs + [(-x) % L for x in u_sq] + [(-x) % L for x in u_inv_sq]
        points = [Q] + list(G_vec) + list(H_vec) + self.L_vec + self.R_vec
        if not (G.msm(scalars, points) == P):
            raise VerificationError("inner product check failed")

    # reference src/inner_product_proof.rs:388-455 (scalars little-endian per A.6)
    def to_bytes(self) -> bytes:
        out = b"".join(l.encode() + r.encode() for l, r in zip(self.L_vec, self.R_vec))
        return out + G.sc_to_bytes(self.a) + G.sc_to_bytes(self.b)

    @staticmethod
    def from_bytes(b: bytes) -> "InnerProductProof":
        if len(b) % 32 or len(b) < 64:
            raise FormatError("length")
        num_elements = len(b) // 32
        if (num_elements - 2) % 2:
            raise FormatError("odd")
        lg_n = (num_elements - 2) // 2
        if lg_n >= 32:
            raise FormatError("too big")
        Lv, Rv = [], []
        for i in range(lg_n):
            l = G.decode(b[64 * i : 64 * i + 32])
            r = G.decode(b[64 * i + 32 : 64 * i + 64])
            if l is None or r is None:
                raise FormatError("point")
            Lv.append(l)
            Rv.append(r)
        pos = 64 * lg_n
        a = int.from_bytes(b[pos : pos + 32], "little")
        bb = int.from_bytes(b[pos + 32 : pos + 64], "little")
        if a >= L or bb >= L:
            raise FormatError("scalar")
        return InnerProductProof(Lv, Rv, a, bb)


# ---------------------------------------------------------------- constraint system
# Variables: ("L", i) | ("R", i) | ("O", i) | ("V", j) | ("1",) | ("0",)
# (reference src/r1cs/linear_combination.rs:15-28)
ONE = ("1",)
ZERO = ("0",)


class LC:
    """reference src/r1cs/linear_combination.rs:119-121: terms keyed by variable."""

    def __init__(self, terms=None):
        self.terms = {}
        for var, c in terms or []:
            self.add_term(var, c)

    def add_term(self, var, c):
        self.terms[var] = (self.terms.get(var, 0) + c) % L

    def __add__(self, o):
        r = LC(list(self.terms.items()))
        for var, c in _lc(o).terms.items():
            r.add_term(var, c)
        return r

    def __sub__(self, o):
        r = LC(list(self.terms.items()))
        for var, c in _lc(o).terms.items():
            r.add_term(var, -c)
        return r

    def __neg__(self):
        return LC([(v, -c) for v, c in self.terms.items()])

    def scale(self, k):
        return LC([(v, c * k) for v, c in self.terms.items()])


def _lc(x) -> LC:
    if isinstance(x, LC):
        return x
    if isinstance(x, tuple):
        return LC([(x, 1)])
    return LC([(ONE, int(x) % L)])


@dataclass
class R1CSProof:
    A_I1: G.Point
    A_O1: G.Point
    S1: G.Point
    A_I2: G.Point
    A_O2: G.Point
    S2: G.Point
    T_1: G.Point
    T_3: G.Point
    T_4: G.Point
    T_5: G.Point
    T_6: G.Point
    t_x: int
    t_x_blinding: int
    e_blinding: int
    ipp_proof: InnerProductProof

    # reference src/r1cs/proof.rs:82-123
    def to_bytes(self) -> bytes:
        one_phase = self.A_I2.is_identity() and self.A_O2.is_identity() and self.S2.is_identity()
        pts = [self.A_I1, self.A_O1, self.S1]
        if not one_phase:
            pts += [self.A_I2, self.A_O2, self.S2]
        pts += [self.T_1, self.T_3, self.T_4, self.T_5, self.T_6]
        out = bytes([0 if one_phase else 1]) + b"".join(p.encode() for p in pts)
        out += G.sc_to_bytes(self.t_x) + G.sc_to_bytes(self.t_x_blinding) + G.sc_to_bytes(self.e_blinding)
        return out + self.ipp_proof.to_bytes()

    # reference src/r1cs/proof.rs:128-207
    @staticmethod
    def from_bytes(b: bytes) -> "R1CSProof":
        if not b:
            raise FormatError("empty")
        version, body = b[0], b[1:]
        if len(body) % 32:
            raise FormatError("length")
        if version == 0:
            npts = 8
        elif version == 1:
            npts = 11
        else:
            raise FormatError("version")
        if len(body) < (npts + 3) * 32:
            raise FormatError("short")
        pts = []
        for i in range(npts):
            p = G.decode(body[32 * i : 32 * i + 32])
            if p is None:
                raise FormatError("point")
            pts.append(p)
        if version == 0:
            pts = pts[:3] + [G.IDENTITY] * 3 + pts[3:]
        off = 32 * npts
        sc = []
        for i in range(3):
            v = int.from_bytes(body[off + 32 * i : off + 32 * i + 32], "little")
            if v >= L:
                raise FormatError("scalar")
            sc.append(v)
        ipp = InnerProductProof.from_bytes(body[off + 96 :])
        return R1CSProof(*pts, sc[0], sc[1], sc[2], ipp)


class _CSBase:
    def multiply(self, left, right):
        raise NotImplementedError

    def specify_randomized_constraints(self, callback):
        self.deferred.append(callback)

    def constrain(self, lc):
        self.constraints.append(_lc(lc))


class Prover(_CSBase):
    """reference src/r1cs/prover.rs"""

    def __init__(self, pc_gens: PedersenGens, transcript: Transcript):
        transcript.r1cs_domain_sep()  # :286
        self.pc_gens, self.transcript = pc_gens, transcript
        self.constraints, self.a_L, self.a_R, self.a_O = [], [], [], []
        self.v, self.v_blinding = [], []
        self.deferred = []
        self.pending_multiplier = None
        self.randomizing = False

    # :319-329
    def commit(self, v: int, v_blinding: int):
        i = len(self.v)
        self.v.append(v % L)
        self.v_blinding.append(v_blinding % L)
        V = self.pc_gens.commit(v, v_blinding)
        self.transcript.append_point(b"V", V)
        return V, ("V", i)

    def commit_public(self, v):  # :169-171
        return self.commit(v, 1)[1]

    def eval(self, lc) -> int:  # :178-194
        tot = 0
        for var, c in _lc(lc).terms.items():
            k = var[0]
            val = {"L": self.a_L, "R": self.a_R, "O": self.a_O, "V": self.v}.get(k)
            x = val[var[1]] if val is not None else (1 if k == "1" else 0)
            tot += c * x
        return tot % L

    def multiply(self, left, right):  # :99-125
        left, right = _lc(left), _lc(right)
        l, r = self.eval(left), self.eval(right)
        i = len(self.a_L)
        self.a_L.append(l)
        self.a_R.append(r)
        self.a_O.append(l * r % L)
        lv, rv, ov = ("L", i), ("R", i), ("O", i)
        self.constrain(left - lv)
        self.constrain(right - rv)
        return lv, rv, ov

    def allocate(self, assignment):  # :127-146
        if self.pending_multiplier is None:
            i = len(self.a_L)
            self.pending_multiplier = i
            self.a_L.append(assignment % L)
            self.a_R.append(0)
            self.a_O.append(0)
            return ("L", i)
        i = self.pending_multiplier
        self.pending_multiplier = None
        self.a_R[i] = assignment % L
        self.a_O[i] = self.a_L[i] * self.a_R[i] % L
        return ("R", i)

    def allocate_multiplier(self, l, r):  # :148-164
        i = len(self.a_L)
        self.a_L.append(l % L)
        self.a_R.append(r % L)
        self.a_O.append(l * r % L)
        return ("L", i), ("R", i), ("O", i)

    def challenge_scalar(self, label: bytes) -> int:  # RandomizingProver :261
        assert self.randomizing
        return self.transcript.challenge_scalar(label)

    def num_multipliers(self):
        return len(self.a_O)

    # :342-379
    def flattened_constraints(self, z):
        n, m = len(self.a_L), len(self.v)
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        exp_z = z
        for lc in self.constraints:
            for var, c in lc.terms.items():
                k = var[0]
                if k == "L":
                    wL[var[1]] = (wL[var[1]] + exp_z * c) % L
                elif k == "R":
                    wR[var[1]] = (wR[var[1]] + exp_z * c) % L
                elif k == "O":
                    wO[var[1]] = (wO[var[1]] + exp_z * c) % L
                elif k == "V":
                    wV[var[1]] = (wV[var[1]] - exp_z * c) % L
            exp_z = exp_z * z % L
        return wL, wR, wO, wV

    # :412-727
    def prove(self, bp_gens: BulletproofGens, blind: Blindings, trace=None) -> R1CSProof:
        tr = self.transcript
        tr.append_u64(b"m", len(self.v))  # :420
        if hasattr(blind, "bind"):
            blind.bind(tr, self.v_blinding)  # :435-445
        n1 = len(self.a_L)
        if bp_gens.gens_capacity < n1:
            raise InvalidGeneratorsLength()
        Bb = self.pc_gens.B_blinding
        i_b1, o_b1, s_b1 = blind.scalar(), blind.scalar(), blind.scalar()  # :457-459
        s_L1, s_R1 = blind.vector_pair(n1)
        Gn, Hn = bp_gens.G(n1), bp_gens.H(n1)
        A_I1 = G.msm([i_b1] + self.a_L + self.a_R, [Bb] + Gn + Hn)  # :465
        A_O1 = G.msm([o_b1] + self.a_O, [Bb] + Gn)  # :477
        S1 = G.msm([s_b1] + s_L1 + s_R1, [Bb] + Gn + Hn)  # :485
        tr.append_point(b"A_I1", A_I1)
        tr.append_point(b"A_O1", A_O1)
        tr.append_point(b"S1", S1)
        # create_randomized_constraints :383-402
        self.pending_multiplier = None
        if not self.deferred:
            tr.r1cs_1phase_domain_sep()
        else:
            tr.r1cs_2phase_domain_sep()
            self.randomizing = True
            cbs, self.deferred = self.deferred, []
            for cb in cbs:
                cb(self)
        n = len(self.a_L)
        n2 = n - n1
        padded_n = 1 << (n - 1).bit_length() if n > 0 else 1  # usize::next_power_of_two (0 -> 1)
        pad = padded_n - n
        if bp_gens.gens_capacity < padded_n:
            raise InvalidGeneratorsLength()
        if n2 > 0:  # :519-530
            i_b2, o_b2, s_b2 = blind.scalar(), blind.scalar(), blind.scalar()
        else:
            i_b2 = o_b2 = s_b2 = 0
        s_L2, s_R2 = blind.vector_pair(n2)
        if n2 > 0:  # :532-565
            G2, H2 = bp_gens.G(n)[n1:], bp_gens.H(n)[n1:]
            A_I2 = G.msm([i_b2] + self.a_L[n1:] + self.a_R[n1:], [Bb] + G2 + H2)
            A_O2 = G.msm([o_b2] + self.a_O[n1:], [Bb] + G2)
            S2 = G.msm([s_b2] + s_L2 + s_R2, [Bb] + G2 + H2)
        else:
            A_I2 = A_O2 = S2 = G.IDENTITY
        tr.append_point(b"A_I2", A_I2)
        tr.append_point(b"A_O2", A_O2)
        tr.append_point(b"S2", S2)
        y = tr.challenge_scalar(b"y")
        z = tr.challenge_scalar(b"z")
        wL, wR, wO, wV = self.flattened_constraints(z)
        y_inv = G.sc_inv(y)
        exp_y_inv = [pow(y_inv, i, L) for i in range(padded_n)]
        s_L, s_R = s_L1 + s_L2, s_R1 + s_R2
        l1, l2, l3 = [0] * n, [0] * n, [0] * n
        r0, r1, r3 = [0] * n, [0] * n, [0] * n
        exp_y = 1
        for i in range(n):  # :596-617
            l1[i] = (self.a_L[i] + exp_y_inv[i] * wR[i]) % L
            l2[i] = self.a_O[i]
            l3[i] = s_L[i]
            r0[i] = (wO[i] - exp_y) % L
            r1[i] = (exp_y * self.a_R[i] + wL[i]) % L
            r3[i] = exp_y * s_R[i] % L
            exp_y = exp_y * y % L
        ip = inner_product  # util.rs:152-170
        t1 = ip(l1, r0)
        t2 = (ip(l1, r1) + ip(l2, r0)) % L
        t3 = (ip(l2, r1) + ip(l3, r0)) % L
        t4 = (ip(l1, r3) + ip(l3, r1)) % L
        t5 = ip(l2, r3)
        t6 = ip(l3, r3)
        tb1, tb3, tb4, tb5, tb6 = (blind.scalar() for _ in range(5))  # :621-625
        pc = self.pc_gens
        T_1, T_3, T_4, T_5, T_6 = pc.commit(t1, tb1), pc.commit(t3, tb3), pc.commit(t4, tb4), pc.commit(t5, tb5), pc.commit(t6, tb6)
        for lab, T in ((b"T_1", T_1), (b"T_3", T_3), (b"T_4", T_4), (b"T_5", T_5), (b"T_6", T_6)):
            tr.append_point(lab, T)
        u = tr.challenge_scalar(b"u")
        x = tr.challenge_scalar(b"x")
        tb2 = sum(c * vb for c, vb in zip(wV, self.v_blinding)) % L  # :644-648

        def poly6(c1, c2, c3, c4, c5, c6):  # util.rs:192-194
            return x * (c1 + x * (c2 + x * (c3 + x * (c4 + x * (c5 + x * c6))))) % L

        t_x = poly6(t1, t2, t3, t4, t5, t6)
        t_x_blinding = poly6(tb1, tb2, tb3, tb4, tb5, tb6)
        l_vec = [(x * (l1[i] + x * (l2[i] + x * l3[i]))) % L for i in range(n)] + [0] * pad  # util.rs:172-181
        r_vec = [(r0[i] + x * (r1[i] + x * (x * r3[i]))) % L for i in range(n)] + [0] * pad
        for i in range(n, padded_n):  # :661-672
            r_vec[i] = (-exp_y) % L
            exp_y = exp_y * y % L
        i_b = (i_b1 + u * i_b2) % L
        o_b = (o_b1 + u * o_b2) % L
        s_b = (s_b1 + u * s_b2) % L
        e_blinding = x * (i_b + x * (o_b + x * s_b)) % L
        tr.append_scalar(b"t_x", t_x)
        tr.append_scalar(b"t_x_blinding", t_x_blinding)
        tr.append_scalar(b"e_blinding", e_blinding)
        w = tr.challenge_scalar(b"w")
        Q = w * pc.B
        G_factors = [1] * n1 + [u] * (n2 + pad)
        H_factors = [exp_y_inv[i] * G_factors[i] % L for i in range(padded_n)]
        if trace is not None:
            trace.update(dict(y=y, z=z, u=u, x=x, w=w, l_vec=l_vec, r_vec=r_vec, G_factors=G_factors, H_factors=H_factors, Q=Q))
        ipp = InnerProductProof.create(tr, Q, G_factors, H_factors, bp_gens.G(padded_n), bp_gens.H(padded_n), l_vec, r_vec)
        return R1CSProof(A_I1, A_O1, S1, A_I2, A_O2, S2, T_1, T_3, T_4, T_5, T_6, t_x, t_x_blinding, e_blinding, ipp)


class Verifier(_CSBase):
    """reference src/r1cs/verifier.rs"""

    def __init__(self, pc_gens: PedersenGens, transcript: Transcript):
        transcript.r1cs_domain_sep()  # :271
        self.pc_gens, self.transcript = pc_gens, transcript
        self.num_vars = 0
        self.V = []
        self.constraints, self.deferred = [], []
        self.pending_multiplier = None
        self.randomizing = False

    def commit(self, V: G.Point):  # :298-308
        i = len(self.V)
        self.V.append(V)
        self.transcript.append_point(b"V", V)
        return ("V", i)

    def commit_public(self, value):  # :153-160
        return self.commit(self.pc_gens.commit(value, 1))

    def eval(self, lc):
        return 0

    def multiply(self, left, right):  # :100-118
        i = self.num_vars
        self.num_vars += 1
        lv, rv, ov = ("L", i), ("R", i), ("O", i)
        self.constrain(_lc(left) - lv)
        self.constrain(_lc(right) - rv)
        return lv, rv, ov

    def allocate(self, _assignment=None):  # :120-134
        if self.pending_multiplier is None:
            i = self.num_vars
            self.num_vars += 1
            self.pending_multiplier = i
            return ("L", i)
        i = self.pending_multiplier
        self.pending_multiplier = None
        return ("R", i)

    def allocate_multiplier(self, _l=None, _r=None):  # :136-150
        i = self.num_vars
        self.num_vars += 1
        return ("L", i), ("R", i), ("O", i)

    def challenge_scalar(self, label):
        assert self.randomizing
        return self.transcript.challenge_scalar(label)

    def num_multipliers(self):
        return self.num_vars

    # :323-362
    def flattened_constraints(self, z):
        n, m = self.num_vars, len(self.V)
        wL, wR, wO, wV, wc = [0] * n, [0] * n, [0] * n, [0] * m, 0
        exp_z = z
        for lc in self.constraints:
            for var, c in lc.terms.items():
                k = var[0]
                if k == "L":
                    wL[var[1]] = (wL[var[1]] + exp_z * c) % L
                elif k == "R":
                    wR[var[1]] = (wR[var[1]] + exp_z * c) % L
                elif k == "O":
                    wO[var[1]] = (wO[var[1]] + exp_z * c) % L
                elif k == "V":
                    wV[var[1]] = (wV[var[1]] - exp_z * c) % L
                elif k == "1":
                    wc = (wc - exp_z * c) % L
            exp_z = exp_z * z % L
        return wL, wR, wO, wV, wc

    def verification_inputs(self, proof: R1CSProof, bp_gens: BulletproofGens, rng_bytes: bytes | None = None):
        """Transcript replay and scalar preparation of `verify` (:398-514): returns the
        mega-MSM's (scalars, points) in the reference's order (:516-547)."""
        tr = self.transcript
        tr.append_u64(b"m", len(self.V))
        n1 = self.num_vars
        tr.validate_and_append_point(b"A_I1", proof.A_I1)
        tr.validate_and_append_point(b"A_O1", proof.A_O1)
        tr.validate_and_append_point(b"S1", proof.S1)
        self.pending_multiplier = None
        if not self.deferred:
            tr.r1cs_1phase_domain_sep()
        else:
            tr.r1cs_2phase_domain_sep()
            self.randomizing = True
            cbs, self.deferred = self.deferred, []
            for cb in cbs:
                cb(self)
        n = self.num_vars
        n2 = n - n1
        padded_n = 1 << (n - 1).bit_length() if n > 0 else 1
        pad = padded_n - n
        if bp_gens.gens_capacity < padded_n:
            raise InvalidGeneratorsLength()
        tr.append_point(b"A_I2", proof.A_I2)
        tr.append_point(b"A_O2", proof.A_O2)
        tr.append_point(b"S2", proof.S2)
        y = tr.challenge_scalar(b"y")
        z = tr.challenge_scalar(b"z")
        for lab, T in ((b"T_1", proof.T_1), (b"T_3", proof.T_3), (b"T_4", proof.T_4), (b"T_5", proof.T_5), (b"T_6", proof.T_6)):
            tr.validate_and_append_point(lab, T)
        u = tr.challenge_scalar(b"u")
        x = tr.challenge_scalar(b"x")
        tr.append_scalar(b"t_x", proof.t_x)
        tr.append_scalar(b"t_x_blinding", proof.t_x_blinding)
        tr.append_scalar(b"e_blinding", proof.e_blinding)
        w = tr.challenge_scalar(b"w")
        wL, wR, wO, wV, wc = self.flattened_constraints(z)
        u_sq, u_inv_sq, s = proof.ipp_proof.verification_scalars(padded_n, tr)
        a, b = proof.ipp_proof.a, proof.ipp_proof.b
        y_inv = G.sc_inv(y)
        y_inv_vec = [pow(y_inv, i, L) for i in range(padded_n)]
        yneg_wR = [wR[i] * y_inv_vec[i] % L for i in range(n)] + [0] * pad
        delta = inner_product(yneg_wR[:n], wL)
        U = [1] * n1 + [u] * (n2 + pad)
        wLp, wOp = wL + [0] * pad, wO + [0] * pad
        g_scalars = [U[i] * (x * yneg_wR[i] - a * s[i]) % L for i in range(padded_n)]
        h_scalars = [U[i] * (y_inv_vec[i] * (x * wLp[i] + wOp[i] - b * s[padded_n - 1 - i]) - 1) % L for i in range(padded_n)]
        if rng_bytes is None:
            r = tr.challenge_scalar(b"r")  # :506 (the mounted fork: a public function of the transcript)
        else:  # upstream: transcript.build_rng().finalize(&mut thread_rng()), then Scalar::random
            r = int.from_bytes(tr.build_rng().finalize(rng_bytes).fill_bytes(64), "little") % L
        xx = x * x % L
        rxx = r * xx % L
        xxx = x * xx % L
        T_scalars = [r * x % L, rxx * x % L, rxx * xx % L, rxx * xxx % L, rxx * xx % L * xx % L]
        scalars = (
            [x, xx, xxx, u * x % L, u * xx % L, u * xxx % L]
            + [wVi * rxx % L for wVi in wV]
            + T_scalars
            + [(w * (proof.t_x - a * b) + r * (xx * (wc + delta) - proof.t_x)) % L]
            + [(-proof.e_blinding - r * proof.t_x_blinding) % L]
            + g_scalars
            + h_scalars
            + u_sq
            + u_inv_sq
        )
        points = (
            [proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2]
            + self.V
            + [proof.T_1, proof.T_3, proof.T_4, proof.T_5, proof.T_6]
            + [self.pc_gens.B, self.pc_gens.B_blinding]
            + bp_gens.G(padded_n)
            + bp_gens.H(padded_n)
            + proof.ipp_proof.L_vec
            + proof.ipp_proof.R_vec
        )
        return scalars, points

    # :393-554
    def verify(self, proof: R1CSProof, bp_gens: BulletproofGens, rng_bytes: bytes | None = None):
        scalars, points = self.verification_inputs(proof, bp_gens, rng_bytes)
        if not G.msm(scalars, points).is_identity():
            raise VerificationError("mega check failed")
