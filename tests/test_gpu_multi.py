"""GPU parity of the multi-GPU building blocks on ONE device (the driver's GPU box has one):
partial sums + combine (the two halves of a stride-sharded MSM), the MPC open of additive
shares (reference src/r1cs_mpc/mpc_prover.rs:621-657: msm_authenticated_iter on shares, then
open), and batch verification.  The two-rank exchange itself is covered on the CPU by
tests/test_multi_gloo.py and on two B200s by tools/multi_gpu_check.py."""
import pytest

from oracle import gadgets
from oracle import group as G
from oracle import protocol as O
from tests.util import points_bytes, rand_point, rand_scalar, rng, scalars_bytes

pytestmark = pytest.mark.gpu
L = G.L


def test_partials_combine_like_two_ranks(ctx):
    from mpc_bulletproof_b200 import Table, multi
    from mpc_bulletproof_b200.api import sum_encode

    r = rng(301)
    n, sets, world = 301, 2, 2
    ps = [rand_point(r) for _ in range(n)]
    ks = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
    parts = b""
    for rank in range(world):
        mine = list(multi.shard_indices(n, rank, world))
        t = Table(ctx, points_bytes([ps[i] for i in mine])).set_windows(0)
        parts += t.msm_partial(b"".join(scalars_bytes([k[i] for i in mine]) for k in ks), n_sets=sets)
        t.close()
    got = sum_encode(ctx, parts, world, sets)
    assert got == [G.msm(k, ps).encode() for k in ks]
    # world = 1 degenerates to the plain MSM through the same engine interface
    t = Table(ctx, points_bytes(ps))
    eng = multi.CudaEngine(ctx, t)
    assert multi.sharded_msm(eng, b"".join(scalars_bytes(k) for k in ks), n_sets=sets) == got
    t.close()


def test_mpc_open_of_shares(ctx):
    """Each party commits to its additive share (and its MAC share) of a_L, a_R and the blinding
    over the shared generators; opening adds the parties' partial sums.  The result must be the
    single prover's A_I (prover.rs:465-475) and key * A_I."""
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200.api import sum_encode

    r = rng(302)
    n = 64
    bp, pc = O.BulletproofGens(n, 1), O.PedersenGens()
    pts = bp.G(n) + bp.H(n) + [pc.B_blinding]
    t = Table(ctx, points_bytes(pts)).set_windows(0)
    aL = [r.randrange(2) for _ in range(n)]  # bits, as in the range gadget
    aR = [(x - 1) % L for x in aL]
    blind, key = rand_scalar(r), rand_scalar(r)
    plain = aL + aR + [blind]
    s0 = [rand_scalar(r) for _ in plain]
    m0 = [rand_scalar(r) for _ in plain]
    shares = [s0, [(x - a) % L for x, a in zip(plain, s0)]]
    macs = [m0, [(key * x - a) % L for x, a in zip(plain, m0)]]
    parts = b"".join(t.msm_partial(scalars_bytes(shares[p]) + scalars_bytes(macs[p]), n_sets=2) for p in range(2))
    opened = sum_encode(ctx, parts, 2, 2)
    A_I = G.msm(plain, pts)
    assert opened[0] == A_I.encode()
    assert opened[1] == (key * A_I).encode()
    # and it is what the single-prover path computes for the same vectors
    assert t.msm(scalars_bytes(plain))[0] == A_I.encode()
    t.close()


def test_batch_verify(ctx):
    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200.protocol import Gens

    pc = O.PedersenGens()
    bp = O.BulletproofGens(64, 1)
    gens = Gens(ctx, points_bytes(bp.G(64)), points_bytes(bp.H(64)), pc.B.encode(), pc.B_blinding.encode())

    def circuit(cs, val, n):
        cs.square_chain(cs.commit_public(val), n)

    cases = [(5, 9), (7, 33), (11, 16), (13, 1)]
    proofs = []
    for val, n in cases:
        p = P.Prover(gens, P.Transcript(b"batch"))
        circuit(p, val, n)
        proofs.append(p.prove(1000 + n))
    bad = bytearray(proofs[1])
    bad[40] ^= 1
    jobs, want = [], []
    for k, (val, n) in enumerate(cases):
        v = P.Verifier(gens, P.Transcript(b"batch"))
        circuit(v, val, n)
        jobs.append((v, proofs[k]))
        want.append(True)
    v = P.Verifier(gens, P.Transcript(b"batch"))  # tampered proof
    circuit(v, *cases[1])
    jobs.append((v, bytes(bad)))
    want.append(False)
    v = P.Verifier(gens, P.Transcript(b"batch"))  # wrong statement
    circuit(v, cases[2][0] + 1, cases[2][1])
    jobs.append((v, proofs[2]))
    want.append(False)
    v = P.Verifier(gens, P.Transcript(b"batch"))  # truncated bytes: FormatError in the reference, a reject here
    circuit(v, *cases[0])
    jobs.append((v, proofs[0][:-7]))
    want.append(False)
    assert P.batch_verify(jobs) == want
    assert P.batch_verify([]) == []
    gens.close()


def test_batch_verify_combined(ctx):
    """Proofs of equal padded size share ONE multiscalar multiplication (sum_k rho_k * check_k); a batch that does not
    pass -- or holds a point that does not decode -- falls back to each proof's own check, so the per-proof answers
    are `Verifier::verify`'s (reference src/r1cs/verifier.rs:393-549) in every case.  The oracle decides the truth."""
    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200.protocol import Gens

    pc = O.PedersenGens()
    bp = O.BulletproofGens(16, 1)
    gens = Gens(ctx, points_bytes(bp.G(16)), points_bytes(bp.H(16)), pc.B.encode(), pc.B_blinding.encode())
    n = 13  # padded to 16 for every proof

    def circuit(cs, val):
        cs.square_chain(cs.commit_public(val), n)

    def oracle_accepts(val, proof):
        ov = O.Verifier(pc, O.Transcript(b"batch2"))
        var = ov.commit_public(val)
        for _ in range(n):
            _, _, var = ov.multiply(var, var)
        try:
            ov.verify(O.R1CSProof.from_bytes(proof), bp)
            return True
        except (O.VerificationError, O.FormatError, ValueError):
            return False

    vals = [100 + 7 * k for k in range(10)]
    proofs = []
    for k, val in enumerate(vals):
        p = P.Prover(gens, P.Transcript(b"batch2"))
        circuit(p, val)
        proofs.append(p.prove(4000 + k))

    def run(ps):
        jobs = []
        for val, pr in zip(vals, ps):
            v = P.Verifier(gens, P.Transcript(b"batch2"))
            circuit(v, val)
            jobs.append((v, pr))
        return P.batch_verify(jobs)

    # every proof valid: one combined check accepts them all
    l0 = ctx.launches
    assert run(proofs) == [True] * len(vals)
    combined_launches = ctx.launches - l0
    # a tampered scalar (t_x) in one proof, a point that is not a valid encoding in another
    bad = list(proofs)
    b = bytearray(bad[3])
    b[11 * 32 + 5] ^= 2  # inside t_x (the 12th 32-byte field of R1CSProof::to_bytes)
    bad[3] = bytes(b)
    b = bytearray(bad[7])
    b[0:32] = b"\xff" * 32  # A_I1
    bad[7] = bytes(b)
    want = [oracle_accepts(val, pr) for val, pr in zip(vals, bad)]
    assert want == [k not in (3, 7) for k in range(len(vals))]
    l0 = ctx.launches
    assert run(bad) == want
    assert ctx.launches - l0 > combined_launches  # the fallback ran
    gens.close()


def test_mpc_share_commitments_equal_the_provers(ctx):
    """Config 5 in miniature (reference integration/mpc_prover.rs): two parties hold additive shares
    of the witness rows and of the blinding factors; each commits to its shares over the shared
    generator table (one partial-sum launch, sets = A_I and A_O), the partials are exchanged and
    added.  The opened A_I1, A_O1 must be the very points the single prover puts in its proof."""
    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200.api import sum_encode
    from mpc_bulletproof_b200.protocol import Gens

    n, seed, val = 16, 4242, 3
    pc, bp = O.PedersenGens(), O.BulletproofGens(n, 1)
    gens = Gens(ctx, points_bytes(bp.G(n)), points_bytes(bp.H(n)), pc.B.encode(), pc.B_blinding.encode())
    p = P.Prover(gens, P.Transcript(b"mpc"))
    p.square_chain(p.commit_public(val), n)
    proof = p.prove(seed)
    A_I1, A_O1 = proof[1:33], proof[33:65]
    # the witness of the squaring chain and the prover's first two blinding draws (prover.rs:457-458)
    aL, x = [], val
    for _ in range(n):
        aL.append(x)
        x = x * x % L
    aR, aO = list(aL), [v * v % L for v in aL]
    blind = O.Blindings(seed)
    i_b, o_b = blind.scalar(), blind.scalar()
    # table order [G | H | B | B_blinding]; set 0 = A_I terms, set 1 = A_O terms
    plain = [aL + aR + [0, i_b], aO + [0] * n + [0, o_b]]
    r = rng(303)
    share0 = [[rand_scalar(r) for _ in row] for row in plain]
    shares = [share0, [[(v - s) % L for v, s in zip(row, s0)] for row, s0 in zip(plain, share0)]]
    t = gens.table
    parts = b"".join(t.msm_partial(b"".join(scalars_bytes(row) for row in shares[party]), n_sets=2) for party in range(2))
    opened = sum_encode(ctx, parts, 2, 2)
    assert opened == [A_I1, A_O1]
    gens.close()


def test_fused_exchange_single_rank(ctx):
    """The fused exchange + combine kernel with world = 1 (push to its own buffer, wait on its own
    flag, add one partial, encode): same bytes as the plain path, step after step (both parities of
    the double buffer).  The multi-rank form is checked on real peers by bench.py --gpus N, whose
    result must equal the NCCL form's, and by tools/multi_gpu_check.py."""
    import torch

    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200.multi import PeerExchange

    r = rng(401)
    n, sets = 120, 2
    ps = [rand_point(r) for _ in range(n)]
    t = Table(ctx, points_bytes(ps)).set_windows(0)
    px = PeerExchange(ctx, max_sets=4)
    dev = torch.device("cuda", ctx.device)
    for step in range(3):
        ks = [[rand_scalar(r) for _ in range(n)] for _ in range(sets)]
        raw = t.msm_partial(b"".join(scalars_bytes(k) for k in ks), n_sets=sets)
        part = torch.frombuffer(bytearray(raw), dtype=torch.int32).to(dev)
        out = torch.zeros(32 * sets, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        px.exchange_sum_encode(part.data_ptr(), sets, out.data_ptr())
        ctx.sync()
        got = bytes(out.cpu().numpy().tobytes())
        assert [got[:32], got[32:]] == [G.msm(k, ps).encode() for k in ks]
    assert px.ok()
    px.close()
    t.close()
