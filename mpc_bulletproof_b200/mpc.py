"""Two-party collaborative proving over secret shares: each party's LOCAL-SHARE work of
`r1cs_mpc::MpcProver::prove` and `SharedInnerProductProof::create` (reference
src/r1cs_mpc/mpc_prover.rs:579-992, src/r1cs_mpc/mpc_inner_product.rs:52-228) on its own GPU through
the C ABI, around an emulation of what the reference delegates to the mpc-stark fabric.

What runs on the GPU (the hot path): the six commitments A_I, A_O, S (two phases)
as ONE multi-output MSM per phase over the value shares and the MAC shares (mpc_prover.rs:621-657,
717-750: `msm_authenticated_iter`), the Pedersen commitments of V_j and T_i on shares
(`commit_shared`), every inner-product round's L and R on shares with both lanes in one MSM
(mpc_inner_product.rs:104-126, 172-186), the folds, and the sums that open shared points.

What is emulated here in plain Python because it is the fabric's job, not the prover's (SURVEY.md
2.1 rows 9, 13, 16: out of scope as a product, needed to drive the path): additive sharing mod l with
SPDZ-style MACs (global key alpha = alpha_0 + alpha_1), Beaver multiplication, `open` /
`open_authenticated`, and a MOCK preprocessing source (`MockDealer`, the counterpart of the
reference's `PartyIDBeaverSource`, integration/main.rs:61-98) that both parties evaluate from a
common seed.  The party-to-party link is any object with `exchange(bytes) -> bytes`:
`TorchLink` (one process per party over `torch.distributed`) or `LocalLink` (two threads).

The transcript is run identically by both parties on opened values only, as in the reference
(mpc_prover.rs:571-577).  The opened proof has the single prover's byte layout
(`PartiallySharedR1CSProof::open`, src/r1cs_mpc/proof.rs:59-82) and verifies with the plain verifier.
"""
from __future__ import annotations

import ctypes
import queue
import random
from dataclasses import dataclass

from . import _lib
from ._lib import lib
from .protocol import Gens, L, Transcript, _raise, sc_bytes


def _inv(x: int) -> int:
    return pow(x, L - 2, L)


def _ints(buf: bytes) -> list[int]:
    return [int.from_bytes(buf[i : i + 32], "little") for i in range(0, len(buf), 32)]


def _bytes(xs) -> bytes:
    return b"".join(sc_bytes(x) for x in xs)


IDENTITY = bytes(32)


# ---------------------------------------------------------------------------------- the link
class LocalLink:
    """Two parties in one process (two threads): a pair of queues."""

    def __init__(self, tx: "queue.Queue", rx: "queue.Queue"):
        self.tx, self.rx = tx, rx
        self.bytes_sent = 0
        self.rounds = 0

    @staticmethod
    def pair():
        a, b = queue.Queue(), queue.Queue()
        return LocalLink(a, b), LocalLink(b, a)

    def exchange(self, payload: bytes) -> bytes:
        self.bytes_sent += len(payload)
        self.rounds += 1
        self.tx.put(payload)
        return self.rx.get(timeout=600)


class TorchLink:
    """One process per party: `torch.distributed` (gloo or nccl group of world size 2).  Both parties send
    payloads of the same length (the protocol is symmetric), so one all_gather is an exchange."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.device = device if device is not None else torch.device("cpu")
        self.bytes_sent = 0
        self.rounds = 0

    def exchange(self, payload: bytes) -> bytes:
        t = self.torch
        self.bytes_sent += len(payload)
        self.rounds += 1
        mine = t.frombuffer(bytearray(payload), dtype=t.uint8).to(self.device)
        out = [t.empty_like(mine), t.empty_like(mine)]
        self.dist.all_gather(out, mine, group=self.group)
        return bytes(out[1 - self.rank].cpu().numpy().tobytes())


# ---------------------------------------------------------------------------------- shared values
@dataclass
class AS:
    """AuthenticatedScalarResult, one party's view: value share and MAC share (sum of MAC shares = alpha * value)."""

    s: int
    m: int

    def __add__(self, o: "AS") -> "AS":
        return AS((self.s + o.s) % L, (self.m + o.m) % L)

    def __sub__(self, o: "AS") -> "AS":
        return AS((self.s - o.s) % L, (self.m - o.m) % L)

    def scale(self, k: int) -> "AS":
        return AS(self.s * k % L, self.m * k % L)


class MockDealer:
    """MOCK preprocessing (the reference's integration tests use one too, integration/main.rs:61-98): MAC key
    shares, sharings of given values, and Beaver triples, all derived from a seed both parties know, each
    party keeping its own half.  Stands in for the offline phase; offers no secrecy."""

    def __init__(self, seed: int, party: int):
        self.rng = random.Random(seed)
        self.party = party
        self.alpha = [self.rng.randrange(L), self.rng.randrange(L)]
        self.alpha_total = sum(self.alpha) % L

    def share(self, value: int) -> AS:
        r = self.rng.randrange(L)
        mr = self.rng.randrange(L)
        mac = self.alpha_total * value % L
        shares = [r, (value - r) % L]
        macs = [mr, (mac - mr) % L]
        return AS(shares[self.party], macs[self.party])

    def share_many(self, values) -> list[AS]:
        return [self.share(v) for v in values]

    def triples(self, k: int):
        xs, ys, zs = [], [], []
        for _ in range(k):
            x, y = self.rng.randrange(L), self.rng.randrange(L)
            xs.append(self.share(x))
            ys.append(self.share(y))
            zs.append(self.share(x * y % L))
        return xs, ys, zs


class MacCheckError(Exception):
    """an authenticated opening failed its MAC check (MpcError::AuthenticationError)"""


class Fabric:
    """What the reference gets from `MpcFabric`: linear algebra on shares is local, products and
    openings cross the link."""

    def __init__(self, party: int, link, dealer: MockDealer, gens: Gens):
        self.party, self.link, self.dealer, self.gens = party, link, dealer, gens
        self.ctx = gens.ctx if gens is not None else None
        self.alpha = dealer.alpha[party]
        self.mults = 0

    # ---- scalars
    def const(self, c: int) -> AS:
        """a public value as a shared one: party 0 holds it, every party holds alpha_p * c of its MAC"""
        return AS(c % L if self.party == 0 else 0, self.alpha * c % L)

    def open(self, xs: list[AS]) -> list[int]:
        theirs = _ints(self.link.exchange(_bytes(x.s for x in xs)))
        return [(x.s + t) % L for x, t in zip(xs, theirs)]

    def open_authenticated(self, xs: list[AS]) -> list[int]:
        vals = self.open(xs)
        checks = [(x.m - self.alpha * v) % L for x, v in zip(xs, vals)]
        theirs = _ints(self.link.exchange(_bytes(checks)))
        if any((c + t) % L for c, t in zip(checks, theirs)):
            raise MacCheckError()
        return vals

    def mul(self, xs: list[AS], ys: list[AS]) -> list[AS]:
        """Beaver multiplication, batched: one exchange opens d = x - a, e = y - b for all pairs."""
        k = len(xs)
        if k == 0:
            return []
        self.mults += k
        ta, tb, tc = self.dealer.triples(k)
        de = self.open([x - a for x, a in zip(xs, ta)] + [y - b for y, b in zip(ys, tb)])
        out = []
        for i in range(k):
            d, e = de[i], de[k + i]
            out.append(tc[i] + tb[i].scale(d) + ta[i].scale(e) + self.const(d * e % L))
        return out

    def inner_product(self, xs: list[AS], ys: list[AS]) -> AS:
        acc = AS(0, 0)
        for p in self.mul(xs, ys):
            acc = acc + p
        return acc

    # ---- points
    def points_sum(self, parts: list[list[bytes]]) -> list[bytes]:
        n_sets = len(parts[0])
        out = ctypes.create_string_buffer(32 * n_sets)
        _raise(lib().bpg_points_sum(self.ctx._h, b"".join(b"".join(p) for p in parts), len(parts), n_sets, out))
        return [out.raw[32 * i : 32 * i + 32] for i in range(n_sets)]

    def open_points(self, mine: list[bytes]) -> list[bytes]:
        theirs = self.link.exchange(b"".join(mine))
        other = [theirs[32 * i : 32 * i + 32] for i in range(len(mine))]
        parts = [mine, other] if self.party == 0 else [other, mine]
        return self.points_sum(parts)

    def open_points_authenticated(self, mine: list[bytes], mine_mac: list[bytes]) -> list[bytes]:
        """open, then every party publishes mac_share - alpha_p * opened; the shares must sum to the identity"""
        opened = self.open_points(mine)
        neg_alpha = sc_bytes(-self.alpha)
        checks = []
        for mac, pt in zip(mine_mac, opened):
            out = ctypes.create_string_buffer(32)
            _raise(lib().bpg_msm(self.ctx._h, sc_bytes(1) + neg_alpha, mac + pt, 2, out))
            checks.append(out.raw)
        total = self.open_points(checks)
        if any(t != IDENTITY for t in total):
            raise MacCheckError()
        return opened


# ---------------------------------------------------------------------------------- the prover
def _pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


class MpcProver:
    """r1cs_mpc::MpcProver, one party (reference src/r1cs_mpc/mpc_prover.rs).  Variables are the tuples of
    `protocol.py` (("L", i), ("R", i), ("O", i), ("V", j), ("1",)); a linear combination is a dict
    {variable: public coefficient}."""

    def __init__(self, fabric: Fabric, transcript: Transcript, blindings):
        """blindings: the preprocessing's stream of random values as plain scalars -- an object with
        `.scalar()` and `.vector_pair(n)`, evaluated identically by both parties and shared through the mock
        dealer (`random_shared_scalars_authenticated`, mpc_prover.rs:602, 695, 835)."""
        self.f, self.tr, self.blind = fabric, transcript, blindings
        self.gens = fabric.gens
        self.tr.append_message(b"dom-sep", b"r1cs v1")  # mpc_prover.rs:151 / transcript.rs r1cs_domain_sep
        self.a_L: list[AS] = []
        self.a_R: list[AS] = []
        self.a_O: list[AS] = []
        self.v: list[AS] = []
        self.v_blinding: list[AS] = []
        self.constraints: list[dict] = []

    # -- mpc_prover.rs:362-400: commit a pre-shared value; the commitment is opened before it enters the transcript
    def commit_preshared(self, v: AS, v_blinding: AS):
        V = self._pedersen_open([v], [v_blinding], authenticated=False)[0]
        self.tr.append_message(b"V", V)
        self.v.append(v)
        self.v_blinding.append(v_blinding)
        return V, ("V", len(self.v) - 1)

    def eval(self, lc: dict) -> AS:
        acc = AS(0, 0)
        for var, c in lc.items():
            k = var[0]
            if k == "1":
                acc = acc + self.f.const(c)
            else:
                src = {"L": self.a_L, "R": self.a_R, "O": self.a_O, "V": self.v}[k]
                acc = acc + src[var[1]].scale(c % L)
        return acc

    # -- mpc_prover.rs:196-224: the product of two shared evaluations is a fabric multiplication
    def multiply(self, left: dict, right: dict):
        l, r = self.eval(left), self.eval(right)
        o = self.f.mul([l], [r])[0]
        i = len(self.a_L)
        self.a_L.append(l)
        self.a_R.append(r)
        self.a_O.append(o)
        lv, rv, ov = ("L", i), ("R", i), ("O", i)
        lc_l, lc_r = dict(left), dict(right)
        lc_l[lv] = (lc_l.get(lv, 0) - 1) % L
        lc_r[rv] = (lc_r.get(rv, 0) - 1) % L
        self.constraints.append(lc_l)
        self.constraints.append(lc_r)
        return lv, rv, ov

    def constrain(self, lc: dict):
        self.constraints.append(dict(lc))

    # -- mpc_prover.rs:456-497 (public weights: every party computes them in the clear)
    def _flatten(self, z: int, n: int, m: int):
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        ez = z
        for lc in self.constraints:
            for var, c in lc.items():
                k = var[0]
                if k == "L":
                    wL[var[1]] = (wL[var[1]] + ez * c) % L
                elif k == "R":
                    wR[var[1]] = (wR[var[1]] + ez * c) % L
                elif k == "O":
                    wO[var[1]] = (wO[var[1]] + ez * c) % L
                elif k == "V":
                    wV[var[1]] = (wV[var[1]] - ez * c) % L
            ez = ez * z % L
        return wL, wR, wO, wV

    # -- commit_shared (generators.rs:46-52 on shares) for both lanes, then open
    def _pedersen_open(self, values: list[AS], blinds: list[AS], authenticated: bool) -> list[bytes]:
        k = len(values)
        pts = self.gens.commit_batch([x.s for x in values] + [x.m for x in values], [x.s for x in blinds] + [x.m for x in blinds])
        if authenticated:
            return self.f.open_points_authenticated(pts[:k], pts[k:])
        return self.f.open_points(pts[:k])

    # -- the three commitments of a phase on shares: ONE six-output MSM (value lane + MAC lane)
    def _phase_commit(self, first: int, cnt: int, ib: AS, ob: AS, sb: AS, sL: list[AS], sR: list[AS]):
        cap = self.gens.gens_capacity
        bb = 2 * cap + 1
        pid, sid, sc = [], [], []
        for lane, pick in ((0, lambda x: x.s), (1, lambda x: x.m)):
            base = 3 * lane
            pid += [bb, bb, bb]
            sid += [base, base + 1, base + 2]
            sc += [pick(ib), pick(ob), pick(sb)]
            for j in range(cnt):
                i = first + j
                pid += [i, cap + i, i, i, cap + i]
                sid += [base, base, base + 1, base + 2, base + 2]
                sc += [pick(self.a_L[i]), pick(self.a_R[i]), pick(self.a_O[i]), pick(sL[j]), pick(sR[j])]
        n = len(pid)
        out = ctypes.create_string_buffer(6 * 32)
        _raise(
            lib().bpg_msm_table_indexed(
                self.f.ctx._h, self.gens.table._h, (ctypes.c_uint32 * n)(*pid), (ctypes.c_uint8 * n)(*sid), _bytes(sc), n, 6, out
            )
        )
        pts = [out.raw[32 * i : 32 * i + 32] for i in range(6)]
        return pts[:3], pts[3:]

    def prove(self) -> bytes:
        f, tr, dealer = self.f, self.tr, self.f.dealer
        cap = self.gens.gens_capacity
        tr.append_u64(b"m", len(self.v))  # :591
        n1 = len(self.a_L)
        # :602-613 the 3 + 2 n1 blinding factors of the first phase, shared by the preprocessing
        draw = [self.blind.scalar() for _ in range(3)]
        sL1, sR1 = self.blind.vector_pair(n1)
        ib1, ob1, sb1 = dealer.share_many(draw)
        sL, sR = dealer.share_many(sL1), dealer.share_many(sR1)
        if cap < n1:
            raise ValueError("InvalidGeneratorsLength")
        sh, _mac = self._phase_commit(0, n1, ib1, ob1, sb1, sL, sR)  # :621-657
        A_I1, A_O1, S1 = f.open_points(sh)  # plain `.open()` in the reference (:630, 641, 657)
        for lab, p in ((b"A_I1", A_I1), (b"A_O1", A_O1), (b"S1", S1)):
            tr.append_message(lab, p)
        tr.append_message(b"dom-sep", b"r1cs-1phase")  # :672: no randomized constraints in this emulation
        n = len(self.a_L)
        N = _pow2(n)
        if cap < N:
            raise ValueError("InvalidGeneratorsLength")
        A_I2 = A_O2 = S2 = IDENTITY  # n2 = 0: identities, zero blindings (:695-699, 760-766)
        for lab, p in ((b"A_I2", A_I2), (b"A_O2", A_O2), (b"S2", S2)):
            tr.append_message(lab, p)
        y, z = tr.challenge_scalar(b"y"), tr.challenge_scalar(b"z")  # :777-778
        wL, wR, wO, wV = self._flatten(z, n, len(self.v))
        y_inv = _inv(y)
        # :785-826 l(X), r(X) on shares; public terms enter as constants
        l1, l2, l3, r0, r1, r3 = [], [], [], [], [], []
        ey, eyi = 1, 1
        for i in range(n):
            l1.append(self.a_L[i] + f.const(eyi * wR[i] % L))
            l2.append(self.a_O[i])
            l3.append(sL[i])
            r0.append(f.const((wO[i] - ey) % L))
            r1.append(self.a_R[i].scale(ey) + f.const(wL[i]))
            r3.append(sR[i].scale(ey))
            ey, eyi = ey * y % L, eyi * y_inv % L
        # :832 t(X) = <l(X), r(X)>: the six coefficients are sums of products of shared values
        prods = f.mul(l1 * 3 + l2 * 3 + l3 * 3, (r0 + r1 + r3) * 3)

        def ip(a_idx, b_idx):  # <l_a, r_b> from the batched products
            acc = AS(0, 0)
            for p in prods[(3 * a_idx + b_idx) * n : (3 * a_idx + b_idx + 1) * n]:
                acc = acc + p
            return acc

        # rows: l1, l2, l3; columns: r0, r1, r3
        t1 = ip(0, 0)
        t2 = ip(0, 1) + ip(1, 0)
        t3 = ip(1, 1) + ip(2, 0)
        t4 = ip(0, 2) + ip(2, 1)
        t5 = ip(1, 2)
        t6 = ip(2, 2)
        tb = dealer.share_many([self.blind.scalar() for _ in range(5)])  # :835
        T = self._pedersen_open([t1, t3, t4, t5, t6], tb, authenticated=True)  # :838-872
        for lab, p in zip((b"T_1", b"T_3", b"T_4", b"T_5", b"T_6"), T):
            tr.append_message(lab, p)
        u, x = tr.challenge_scalar(b"u"), tr.challenge_scalar(b"x")  # :885-886
        tb2 = AS(0, 0)
        for c, vb in zip(wV, self.v_blinding):  # :890-894
            tb2 = tb2 + vb.scale(c)

        def poly6(c):  # x (c1 + x (c2 + ...)) on shares
            acc = AS(0, 0)
            for ci in reversed(c):
                acc = (acc + ci).scale(x)
            return acc

        t_x = poly6([t1, t2, t3, t4, t5, t6])
        t_xb = poly6([tb[0], tb2, tb[1], tb[2], tb[3], tb[4]])
        x2, x3 = x * x % L, x * x % L * x % L
        l_vec = [l1[i].scale(x) + l2[i].scale(x2) + l3[i].scale(x3) for i in range(n)] + [AS(0, 0)] * (N - n)
        r_vec = [r0[i] + r1[i].scale(x) + r3[i].scale(x3) for i in range(n)]
        for i in range(n, N):  # :913-918 padding: r[i] = -y^i
            r_vec.append(f.const(-ey % L))
            ey = ey * y % L
        e_b = (ib1 + (ob1 + sb1.scale(x)).scale(x)).scale(x)  # i_blinding2 = ... = 0 (:921-925)
        t_x_o, t_xb_o, e_b_o = f.open([t_x, t_xb, e_b])  # `open_batch`, :930
        tr.append_message(b"t_x", sc_bytes(t_x_o))
        tr.append_message(b"t_x_blinding", sc_bytes(t_xb_o))
        tr.append_message(b"e_blinding", sc_bytes(e_b_o))
        w = tr.challenge_scalar(b"w")  # :946; Q = w * B
        Gf = [1] * n1 + [u] * (N - n1)
        Hf, eyi = [], 1
        for i in range(N):
            Hf.append(eyi * Gf[i] % L)
            eyi = eyi * y_inv % L
        ipp = self._shared_ipp(w, Gf, Hf, l_vec, r_vec)
        out = bytes([0]) + A_I1 + A_O1 + S1 + b"".join(T) + sc_bytes(t_x_o) + sc_bytes(t_xb_o) + sc_bytes(e_b_o) + ipp
        return out

    # -- SharedInnerProductProof::create + open (mpc_inner_product.rs:52-271)
    def _shared_ipp(self, w: int, Gf, Hf, a: list[AS], b: list[AS]) -> bytes:
        f, tr = self.f, self.tr
        N = len(a)
        cap = self.gens.gens_capacity
        tr.append_message(b"dom-sep", b"ipp v1")
        tr.append_u64(b"n", N)
        lanes = 2
        st = ctypes.c_void_p()
        av = _bytes(x.s for x in a) + _bytes(x.m for x in a)
        bv = _bytes(x.s for x in b) + _bytes(x.m for x in b)
        _raise(
            lib().bpg_ipp_begin_shares(
                f.ctx._h, self.gens.table._h, 0, cap, 2 * cap, sc_bytes(w), N, lanes, _bytes(Gf), _bytes(Hf), av, bv, ctypes.byref(st)
            )
        )
        out = b""
        try:
            while lib().bpg_ipp_rounds_left(st):
                m = lib().bpg_ipp_len(st)
                h = m // 2
                ab, bb = ctypes.create_string_buffer(lanes * m * 32), ctypes.create_string_buffer(lanes * m * 32)
                _raise(lib().bpg_ipp_read_ab(st, ab, bb))
                ai, bi = _ints(ab.raw), _ints(bb.raw)
                A = [AS(ai[i], ai[m + i]) for i in range(m)]
                B = [AS(bi[i], bi[m + i]) for i in range(m)]
                # c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> (:104-105, 172-173): Beaver products over the link
                pr = f.mul(A[:h] + A[h:], B[h:] + B[:h])
                cL, cR = AS(0, 0), AS(0, 0)
                for p in pr[:h]:
                    cL = cL + p
                for p in pr[h:]:
                    cR = cR + p
                Ls, Rs = ctypes.create_string_buffer(32 * lanes), ctypes.create_string_buffer(32 * lanes)
                _raise(
                    lib().bpg_ipp_round_LR_shares(
                        st, sc_bytes(cL.s) + sc_bytes(cL.m), sc_bytes(cR.s) + sc_bytes(cR.m), Ls, Rs
                    )
                )
                # `open_authenticated` of L and R (:131, 191)
                Lo, Ro = f.open_points_authenticated([Ls.raw[:32], Rs.raw[:32]], [Ls.raw[32:], Rs.raw[32:]])
                tr.append_message(b"L", Lo)
                tr.append_message(b"R", Ro)
                out += Lo + Ro
                uu = tr.challenge_scalar(b"u")
                _raise(lib().bpg_ipp_round_fold(st, sc_bytes(uu), sc_bytes(_inv(uu))))
            fa, fb = ctypes.create_string_buffer(32 * lanes), ctypes.create_string_buffer(32 * lanes)
            _raise(lib().bpg_ipp_finish_shares(st, fa, fb))
        finally:
            lib().bpg_ipp_free(st)
        ai, bi = _ints(fa.raw), _ints(fb.raw)
        a_o, b_o = f.open_authenticated([AS(ai[0], ai[1]), AS(bi[0], bi[1])])  # :245-256
        return out + sc_bytes(a_o) + sc_bytes(b_o)
