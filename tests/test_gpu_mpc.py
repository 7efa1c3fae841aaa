"""BASELINE.json config 5 / SURVEY.md 8a-9: the two-party MpcProver with each party's local-share MSMs,
inner-product rounds and folds on the GPU (through the C ABI), against the single prover.

The reference's own oracle for the MPC prover is the plain verifier (integration/mpc_prover.rs:143-190);
here the opened proof must, in addition, equal BYTE FOR BYTE the single prover's proof for the same
witness and blinding values (the product's deterministic prover and the oracle's), because additive
shares of the scalars give additive shares of every commitment.

Two parties = two threads with one context each (so the suite runs on one GPU), and once as two
processes over `torch.distributed` (party p on GPU p when there are two)."""
import os
import random
import socket
import threading

import pytest

from oracle import group as G
from oracle import protocol as O
from tests.util import points_bytes

pytestmark = pytest.mark.gpu
L = G.L
CAP = 64
LABEL = b"mpc-r1cs"  # integration/mpc_inner_product.rs TRANSCRIPT_SEED plays this role


def _gens_bytes():
    pc = O.PedersenGens()
    bp = O.BulletproofGens(CAP, 1)
    return pc, bp, (points_bytes(bp.G(CAP)), points_bytes(bp.H(CAP)), pc.B.encode(), pc.B_blinding.encode())


# ---- circuits: (values, blinding values) of the committed inputs and a builder over any prover-like object
def simple_circuit(c_out):
    """integration/mpc_prover.rs:46-80: (5 a1 + 10 a2) * (2 b1 + 3 b2) == c"""
    vals = [2, 3, 4, 5, c_out]

    def build(cs, vs, lc):
        a1, a2, b1, b2, c = vs
        _, _, o = cs.multiply(lc({a1: 5, a2: 10}), lc({b1: 2, b2: 3}))
        cs.constrain(lc({o: 1, c: L - 1}))

    return vals, build


def chain_circuit(k):
    """k dependent multipliers: x_{i+1} = (x_i + i) * (x_i + v1); the last output equals the committed v2.
    k = 5 pads to 8 (the padding path of mpc_prover.rs:913-918), k = 40 to 64 (six inner-product rounds)."""
    v0, v1 = 7, 11
    x = v0
    for i in range(k):
        x = (x + i) * (x + v1) % L
    vals = [v0, v1, x]

    def build(cs, vs, lc):
        cur = lc({vs[0]: 1})
        for i in range(k):
            left = dict(cur.terms if hasattr(cur, "terms") else cur)
            right = dict(left)
            one = ("1",)
            left[one] = (left.get(one, 0) + i) % L
            right[vs[1]] = (right.get(vs[1], 0) + 1) % L
            _, _, o = cs.multiply(lc(left), lc(right))
            cur = lc({o: 1})
        fin = dict(cur.terms if hasattr(cur, "terms") else cur)
        fin[vs[2]] = L - 1
        cs.constrain(lc(fin))

    return vals, build


def _single_prover_bytes(vals, vblinds, build, seed, gens):
    """the same statement through the product's single prover (deterministic blindings) and the oracle"""
    from mpc_bulletproof_b200 import protocol as P

    class LC:  # what protocol._terms accepts
        def __init__(self, t):
            self.terms = t

    p = P.Prover(gens, P.Transcript(LABEL))
    cv = [p.commit(v, b) for v, b in zip(vals, vblinds)]
    build(p, [var for _, var in cv], LC)
    proof = p.prove(seed)
    pc, bp, _ = _gens_bytes()
    op = O.Prover(pc, O.Transcript(LABEL))
    ocv = [op.commit(v, b) for v, b in zip(vals, vblinds)]
    build(op, [var for _, var in ocv], lambda t: O.LC(list(t.items())))
    want = op.prove(bp, O.Blindings(seed)).to_bytes()
    assert [c for c, _ in cv] == [c.encode() for c, _ in ocv]
    return proof, want, [c for c, _ in cv]


def _party(party, link, gens_bytes, vals, vblinds, build, seed, device, tamper=None):
    from mpc_bulletproof_b200 import Context
    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200.mpc import AS, Fabric, MockDealer, MpcProver

    ctx = Context(device)
    gens = P.Gens(ctx, *gens_bytes)
    try:
        f = Fabric(party, link, MockDealer(0xD0 + seed, party), gens)
        pr = MpcProver(f, P.Transcript(LABEL), O.Blindings(seed))
        coms, vs = [], []
        for v, b in zip(vals, vblinds):
            sv, sb = f.dealer.share(v), f.dealer.share(b)  # `allocate_preshared`: the inputs arrive as shares
            c, var = pr.commit_preshared(sv, sb)
            coms.append(c)
            vs.append(var)
        build(pr, vs, dict)
        if tamper:
            tamper(party, pr)
        proof = pr.prove()
        return proof, coms, {"launches": ctx.launches, "mults": f.mults, "link_bytes": link.bytes_sent, "link_rounds": link.rounds}
    finally:
        gens.close()
        ctx.close()


def _two_threads(gens_bytes, vals, vblinds, build, seed, tamper=None):
    from mpc_bulletproof_b200.mpc import LocalLink

    la, lb = LocalLink.pair()
    out, err = [None, None], [None, None]

    def run(p, link):
        try:
            out[p] = _party(p, link, gens_bytes, vals, vblinds, build, seed, 0, tamper)
        except Exception as e:  # noqa: BLE001
            err[p] = e
            link.tx.put(b"")  # unblock the peer

    ts = [threading.Thread(target=run, args=(0, la)), threading.Thread(target=run, args=(1, lb))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(600)
    return out, err


def _verify(gens, coms, build, proof):
    from mpc_bulletproof_b200 import protocol as P

    class LC:
        def __init__(self, t):
            self.terms = t

    v = P.Verifier(gens, P.Transcript(LABEL))
    build(v, [v.commit(c) for c in coms], LC)
    v.verify(proof)


@pytest.mark.parametrize("case", ["simple", "chain5", "chain40"])
def test_mpc_prover_equals_single_prover(ctx, case):
    from mpc_bulletproof_b200 import protocol as P

    vals, build = {"simple": lambda: simple_circuit(920), "chain5": lambda: chain_circuit(5), "chain40": lambda: chain_circuit(40)}[case]()
    seed = {"simple": 5, "chain5": 6, "chain40": 7}[case]
    r = random.Random(seed)
    vblinds = [r.randrange(L) for _ in vals]
    pc, bp, gb = _gens_bytes()
    gens = P.Gens(ctx, *gb)
    single, want, coms = _single_prover_bytes(vals, vblinds, build, seed, gens)
    assert single == want
    out, err = _two_threads(gb, vals, vblinds, build, seed)
    assert err == [None, None], err
    (p0, c0, st0), (p1, c1, st1) = out
    assert p0 == p1, "the parties opened different proofs"
    assert c0 == c1 == coms, "opened V commitments differ from the single prover's"
    assert p0 == single, "opened MPC proof differs from the single prover's proof bytes"
    _verify(gens, coms, build, p0)  # the reference's oracle for the MPC prover: the plain verifier accepts
    ov = O.Verifier(pc, O.Transcript(LABEL))
    build(ov, [ov.commit(G.decode(c)) for c in coms], lambda t: O.LC(list(t.items())))
    ov.verify(O.R1CSProof.from_bytes(p0), bp)
    assert st0["launches"] > 0 and st0["mults"] == st1["mults"] > 0
    gens.close()


def test_mpc_false_statement_is_rejected(ctx):
    """integration/mpc_prover.rs test_r1cs_false_statement: the parties produce a proof, the verifier rejects it"""
    from mpc_bulletproof_b200 import protocol as P

    vals, build = simple_circuit(921)
    vblinds = [11, 12, 13, 14, 1]
    _, _, gb = _gens_bytes()
    out, err = _two_threads(gb, vals, vblinds, build, 9)
    assert err == [None, None], err
    gens = P.Gens(ctx, *gb)
    with pytest.raises(P.VerificationError):
        _verify(gens, out[0][1], build, out[0][0])
    gens.close()


def test_mpc_mac_check_catches_a_corrupted_share():
    """a party that alters a witness share without fixing its MAC is caught at an authenticated opening"""
    from mpc_bulletproof_b200.mpc import AS, MacCheckError

    vals, build = simple_circuit(920)
    _, _, gb = _gens_bytes()

    def tamper(party, pr):
        if party == 1:
            x = pr.a_O[0]
            pr.a_O[0] = AS((x.s + 1) % L, x.m)

    out, err = _two_threads(gb, vals, [1, 2, 3, 4, 5], build, 10, tamper)
    assert any(isinstance(e, MacCheckError) for e in err), err


def test_ipp_rounds_on_shares_match_the_plain_rounds(ctx):
    """bpg_ipp_begin_shares / round_LR_shares / read_ab / finish_shares directly: random additive splits of
    a and b (lane 0) and of alpha*a, alpha*b (lane 1), caller-supplied cross-term shares; per round the sums
    of the two parties' L/R shares equal the oracle's L/R (lane 0) and alpha times them (lane 1)."""
    import ctypes

    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200._lib import lib
    from mpc_bulletproof_b200.protocol import _raise, sc_bytes

    n = 16
    r = random.Random(31)
    pc, bp, gb = _gens_bytes()
    gens = P.Gens(ctx, *gb)
    alpha, w = r.randrange(L), r.randrange(1, L)
    a = [r.randrange(L) for _ in range(n)]
    b = [r.randrange(L) for _ in range(n)]
    Gf = [r.randrange(1, L) for _ in range(n)]
    Hf = [r.randrange(1, L) for _ in range(n)]
    Q = w * pc.B
    tr_o = O.Transcript(b"shares")
    want = O.InnerProductProof.create(tr_o, Q, Gf, Hf, bp.G(n), bp.H(n), a, b)

    def split(v):
        s0 = [r.randrange(L) for _ in v]
        return s0, [(x - y) % L for x, y in zip(v, s0)]

    a_sh, b_sh = split(a), split(b)
    am_sh, bm_sh = split([alpha * x % L for x in a]), split([alpha * x % L for x in b])
    bts = lambda xs: b"".join(sc_bytes(x) for x in xs)  # noqa: E731
    sts = []
    for p in range(2):
        st = ctypes.c_void_p()
        _raise(lib().bpg_ipp_begin_shares(ctx._h, gens.table._h, 0, CAP, 2 * CAP, sc_bytes(w), n, 2, bts(Gf), bts(Hf),
                                          bts(a_sh[p]) + bts(am_sh[p]), bts(b_sh[p]) + bts(bm_sh[p]), ctypes.byref(st)))
        assert lib().bpg_ipp_lanes(st) == 2
        sts.append(st)
    tr = O.Transcript(b"shares")
    tr.innerproduct_domain_sep(n)
    ints = lambda raw: [int.from_bytes(raw[i : i + 32], "little") for i in range(0, len(raw), 32)]  # noqa: E731
    rnd = 0
    while lib().bpg_ipp_rounds_left(sts[0]):
        m = lib().bpg_ipp_len(sts[0])
        h = m // 2
        cur = []
        for p in range(2):
            ab, bb = ctypes.create_string_buffer(2 * m * 32), ctypes.create_string_buffer(2 * m * 32)
            _raise(lib().bpg_ipp_read_ab(sts[p], ab, bb))
            cur.append((ints(ab.raw), ints(bb.raw)))
        A = [(cur[0][0][i] + cur[1][0][i]) % L for i in range(m)]  # lane 0 opened (test only)
        B = [(cur[0][1][i] + cur[1][1][i]) % L for i in range(m)]
        assert [(cur[0][0][m + i] + cur[1][0][m + i]) % L for i in range(m)] == [alpha * x % L for x in A]
        cL = sum(A[i] * B[h + i] for i in range(h)) % L
        cR = sum(A[h + i] * B[i] for i in range(h)) % L
        shares = []
        for p in range(2):
            # this party's shares of the cross terms and of their MACs (any additive split)
            cl0, cr0, clm0, crm0 = (r.randrange(L) for _ in range(4))
            if p == 0:
                mine = (cl0, cr0, clm0, crm0)
                keep = mine
            else:
                mine = ((cL - keep[0]) % L, (cR - keep[1]) % L, (alpha * cL - keep[2]) % L, (alpha * cR - keep[3]) % L)
            Ls, Rs = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
            _raise(lib().bpg_ipp_round_LR_shares(sts[p], sc_bytes(mine[0]) + sc_bytes(mine[2]), sc_bytes(mine[1]) + sc_bytes(mine[3]), Ls, Rs))
            shares.append((Ls.raw, Rs.raw))
        out = ctypes.create_string_buffer(4 * 32)
        parts = b"".join(s[0] + s[1] for s in shares)  # [part][L0 Lmac R0 Rmac]
        _raise(lib().bpg_points_sum(ctx._h, parts, 2, 4, out))
        Lo, Lm, Ro, Rm = (out.raw[32 * i : 32 * i + 32] for i in range(4))
        assert Lo == want.L_vec[rnd].encode() and Ro == want.R_vec[rnd].encode()
        assert Lm == (alpha * want.L_vec[rnd]).encode() and Rm == (alpha * want.R_vec[rnd]).encode()
        tr.append_point(b"L", G.decode(Lo))
        tr.append_point(b"R", G.decode(Ro))
        u = tr.challenge_scalar(b"u")
        for p in range(2):
            _raise(lib().bpg_ipp_round_fold(sts[p], sc_bytes(u), sc_bytes(pow(u, L - 2, L))))
        rnd += 1
    fin = []
    for p in range(2):
        fa, fb = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        _raise(lib().bpg_ipp_finish_shares(sts[p], fa, fb))
        fin.append((ints(fa.raw), ints(fb.raw)))
        lib().bpg_ipp_free(sts[p])
    assert (fin[0][0][0] + fin[1][0][0]) % L == want.a and (fin[0][1][0] + fin[1][1][0]) % L == want.b
    assert (fin[0][0][1] + fin[1][0][1]) % L == alpha * want.a % L
    # invalid share encodings are a format error, not a wrong sum
    bad = bytes([0xFF] * 32) + bytes(32)
    with pytest.raises(P.FormatError):
        _raise(lib().bpg_points_sum(ctx._h, bad, 2, 1, ctypes.create_string_buffer(32)))
    gens.close()


# ---- two processes, party p on GPU p (or both on GPU 0): the link is torch.distributed
def _proc_party(rank, port, q, gb, vals, vblinds, seed, k):
    import torch
    import torch.distributed as dist

    from mpc_bulletproof_b200.mpc import TorchLink

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    dev = rank % max(torch.cuda.device_count(), 1)
    _, build = chain_circuit(k)
    try:
        q.put((rank, _party(rank, TorchLink(), gb, vals, vblinds, build, seed, dev), dev))
    except Exception as e:  # noqa: BLE001
        q.put((rank, e, dev))
    dist.destroy_process_group()


def test_mpc_two_processes(ctx):
    import torch.multiprocessing as mp

    from mpc_bulletproof_b200 import protocol as P

    k, seed = 12, 21
    vals, build = chain_circuit(k)
    vblinds = [101, 102, 103]
    pc, bp, gb = _gens_bytes()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpx = mp.get_context("spawn")
    q = mpx.Queue()
    ps = [mpx.Process(target=_proc_party, args=(r, port, q, gb, vals, vblinds, seed, k)) for r in range(2)]
    for p in ps:
        p.start()
    res = {}
    for _ in range(2):
        rank, val, dev = q.get(timeout=600)
        res[rank] = val
    for p in ps:
        p.join(120)
    for v in res.values():
        assert not isinstance(v, Exception), v
    gens = P.Gens(ctx, *gb)
    single, want, coms = _single_prover_bytes(vals, vblinds, build, seed, gens)
    assert res[0][0] == res[1][0] == single == want
    assert res[0][1] == coms
    _verify(gens, coms, build, res[0][0])
    gens.close()
