"""GPU parity of the protocol layer: InnerProductProof and the R1CS Prover/Verifier of the
product (CUDA engine behind the C ABI) against the oracle's restatement of the reference,
on the same seeded inputs.  Bar: byte-identical proofs and identical accept/reject.

Mirrors the reference's own tests: make_ipp_{1,2,4,32,64}
(src/inner_product_proof.rs:507-583), the shuffle / example / range-proof gadget tests
(tests/r1cs.rs:136-214, 542-587, 655-703) and its serialization roundtrip."""
import random

import pytest

from oracle import gadgets
from oracle import group as G
from oracle import protocol as O
from tests.util import points_bytes, scalars_bytes

pytestmark = pytest.mark.gpu
L = G.L


@pytest.fixture(scope="module")
def env(ctx):
    from mpc_bulletproof_b200.protocol import Gens

    pc = O.PedersenGens()
    bp = O.BulletproofGens(128, 1)
    gens = Gens(ctx, points_bytes(bp.G(128)), points_bytes(bp.H(128)), pc.B.encode(), pc.B_blinding.encode())
    return pc, bp, gens


# ------------------------------------------------------------------ inner product proof
def _ipp_case(ctx, n, seed):
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200 import protocol as P

    r = random.Random(seed)
    bp = O.BulletproofGens(n, 1)
    Gs, Hs = bp.G(n), bp.H(n)
    Q = G.hash_to_group_sha512(b"test point")
    a = [r.randrange(L) for _ in range(n)]
    b = [r.randrange(L) for _ in range(n)]
    c = O.inner_product(a, b)
    y_inv = r.randrange(1, L)
    Gf = [1] * n
    Hf = [pow(y_inv, i, L) for i in range(n)]
    # P = <a,G> + <b',H> + c Q with b'_i = b_i y^-i  (helper of inner_product_proof.rs:481-500)
    Ppt = G.msm(a + [b[i] * Hf[i] % L for i in range(n)] + [c], Gs + Hs + [Q])
    want = O.InnerProductProof.create(O.Transcript(b"innerproducttest"), Q, Gf, Hf, Gs, Hs, a, b)
    tG, tH = Table(ctx, points_bytes(Gs)), Table(ctx, points_bytes(Hs))
    got = P.InnerProductProof.create(ctx, P.Transcript(b"innerproducttest"), Q.encode(), Gf, Hf, tG, tH, a, b)
    assert got.to_bytes() == want.to_bytes()
    assert len(got.to_bytes()) == 32 * (2 * (n.bit_length() - 1) + 2)
    # verify, then verify again from bytes (roundtrip), then a tampered proof must fail
    got.verify(ctx, n, P.Transcript(b"innerproducttest"), Gf, Hf, Ppt.encode(), Q.encode(), tG, tH)
    P.InnerProductProof.from_bytes(want.to_bytes()).verify(
        ctx, n, P.Transcript(b"innerproducttest"), Gf, Hf, Ppt.encode(), Q.encode(), tG, tH
    )
    bad = bytearray(got.to_bytes())
    bad[-1] ^= 1
    with pytest.raises((P.VerificationError, P.FormatError)):
        P.InnerProductProof(bytes(bad)).verify(ctx, n, P.Transcript(b"innerproducttest"), Gf, Hf, Ppt.encode(), Q.encode(), tG, tH)
    with pytest.raises(P.VerificationError):
        got.verify(ctx, n, P.Transcript(b"innerproducttest"), Gf, Hf, (Ppt + Q).encode(), Q.encode(), tG, tH)


@pytest.mark.parametrize("n", [1, 2, 4, 32, 64])
def test_make_ipp(ctx, n):
    _ipp_case(ctx, n, 1000 + n)


def test_ipp_general_factors(ctx):
    """G_factors other than one (the R1CS prover passes u for the second phase)."""
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200 import protocol as P

    n = 16
    r = random.Random(77)
    bp = O.BulletproofGens(n, 1)
    Gs, Hs = bp.G(n), bp.H(n)
    Q = G.hash_to_group_sha512(b"another point")
    a = [r.randrange(L) for _ in range(n)]
    b = [r.randrange(L) for _ in range(n)]
    Gf = [r.randrange(1, L) for _ in range(n)]
    Hf = [r.randrange(1, L) for _ in range(n)]
    want = O.InnerProductProof.create(O.Transcript(b"t"), Q, Gf, Hf, Gs, Hs, a, b)
    got = P.InnerProductProof.create(
        ctx, P.Transcript(b"t"), Q.encode(), Gf, Hf, Table(ctx, points_bytes(Gs)), Table(ctx, points_bytes(Hs)), a, b
    )
    assert got.to_bytes() == want.to_bytes()


def test_ipp_rejects_non_power_of_two(ctx):
    from mpc_bulletproof_b200 import BpgError, Table
    from mpc_bulletproof_b200 import protocol as P
    from mpc_bulletproof_b200._lib import BPG_ERR_POW2

    bp = O.BulletproofGens(4, 1)
    t = Table(ctx, points_bytes(bp.G(4)))
    with pytest.raises(BpgError) as e:
        P.InnerProductProof.create(ctx, P.Transcript(b"t"), G.BASEPOINT.encode(), [1] * 3, [1] * 3, t, t, [1, 2, 3], [4, 5, 6])
    assert e.value.code == BPG_ERR_POW2


# ------------------------------------------------------------------ commitments
def test_pedersen_commit(env):
    pc, bp, gens = env
    r = random.Random(3)
    vs = [0, 1, L - 1] + [r.randrange(L) for _ in range(5)]
    bs = [0, L - 1, 1] + [r.randrange(L) for _ in range(5)]
    got = gens.commit_batch(vs, bs)
    for v, b, c in zip(vs, bs, got):
        assert c == pc.commit(v, b).encode()


# ------------------------------------------------------------------ R1CS
def _both(env, label, build, seed):
    """Runs `build(prover, rng)` against the oracle Prover and the product Prover with the
    same blinding seed; returns (oracle proof bytes, product proof bytes, commitments)."""
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env
    out = []
    for prover_cls in ("oracle", "product"):
        r = random.Random(seed)
        if prover_cls == "oracle":
            p = O.Prover(pc, O.Transcript(label))
            coms = build(p, r, lambda c: c.encode())
            out.append((p.prove(bp, O.Blindings(seed)).to_bytes(), coms))
        else:
            p = P.Prover(gens, P.Transcript(label))
            coms = build(p, r, lambda c: c)
            out.append((p.prove(seed), coms))
    return out


def _verify_both(env, label, build_v, proof, coms):
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env
    ov = O.Verifier(pc, O.Transcript(label))
    build_v(ov, [G.decode(c) for c in coms])
    oracle_ok = True
    try:
        ov.verify(O.R1CSProof.from_bytes(proof), bp)
    except O.VerificationError:
        oracle_ok = False
    pv = P.Verifier(gens, P.Transcript(label))
    build_v(pv, coms)
    product_ok = True
    try:
        pv.verify(proof)
    except P.VerificationError:
        product_ok = False
    assert oracle_ok == product_ok
    return product_ok


@pytest.mark.parametrize("c2,ok", [(9, True), (10, False)])
def test_example_gadget(env, c2, ok):
    """tests/r1cs.rs:542-563: (3+4)*(6+1) = 40+9 holds, = 40+10 is rejected."""

    def build(p, r, enc):
        cv = [p.commit(x, r.randrange(L)) for x in (3, 4, 6, 1, 40)]
        v = [var for _, var in cv]
        gadgets.example_gadget(p, v[0], v[1], v[2], v[3], v[4], c2)
        return [enc(c) for c, _ in cv]

    def build_v(vf, coms):
        v = [vf.commit(c) for c in coms]
        gadgets.example_gadget(vf, v[0], v[1], v[2], v[3], v[4], c2)

    (op, oc), (pp, pcoms) = _both(env, b"R1CSExampleGadget", build, 11)
    assert oc == pcoms
    assert op == pp, "proof bytes differ from the oracle"
    assert pp[0] == 0 and len(pp) == 1 + 11 * 32 + 32 * 2  # one-phase proof, n = 1
    assert _verify_both(env, b"R1CSExampleGadget", build_v, pp, pcoms) == ok


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 24, 42])
def test_shuffle_gadget(env, k):
    """tests/r1cs.rs:136-214: k-shuffles incl. non-power-of-two multiplier counts (padding)."""
    label = b"ShuffleProofTest"

    def mk(bad):
        def build(p, r, enc):
            inp = [r.randrange(2**64) for _ in range(k)]
            outp = inp[:]
            r.shuffle(outp)
            if bad:
                outp[0] += 1
            p.transcript.append_message(b"dom-sep", b"ShuffleProof")
            p.transcript.append_u64(b"k", k)
            # re-create the prover after the transcript prefix, as the reference does
            ic = [p.commit(v, r.randrange(L)) for v in inp]
            oc = [p.commit(v, r.randrange(L)) for v in outp]
            gadgets.shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
            return [enc(c) for c, _ in ic + oc]

        return build

    def build_v(vf, coms):
        vf.transcript.append_message(b"dom-sep", b"ShuffleProof")
        vf.transcript.append_u64(b"k", k)
        vs = [vf.commit(c) for c in coms]
        gadgets.shuffle_gadget(vf, vs[:k], vs[k:])

    (op, oc), (pp, pcoms) = _both(env, label, mk(False), 100 + k)
    assert oc == pcoms and op == pp
    assert pp[0] == (0 if k == 1 else 1)
    assert _verify_both(env, label, build_v, pp, pcoms) is True
    if k in (2, 5):
        (op, oc), (pp, pcoms) = _both(env, label, mk(True), 200 + k)
        assert op == pp
        assert _verify_both(env, label, build_v, pp, pcoms) is False


@pytest.mark.parametrize("n", [2, 10, 32, 63])
def test_range_proof_gadget(env, n):
    """tests/r1cs.rs:655-703: in-range values verify, max+1 is rejected."""
    label = b"RangeProofTest"
    r0 = random.Random(n)
    mx = (1 << n) - 1
    for val, ok in [(r0.randrange(0, mx), True), (mx + 1, False)]:

        def build(p, r, enc):
            com, var = p.commit(val, r.randrange(L))
            gadgets.range_proof_gadget(p, var, val, n)
            return [enc(com)]

        def build_v(vf, coms):
            var = vf.commit(coms[0])
            gadgets.range_proof_gadget(vf, var, None, n)

        (op, oc), (pp, pcoms) = _both(env, label, build, 300 + n)
        assert oc == pcoms and op == pp
        assert _verify_both(env, label, build_v, pp, pcoms) == ok


def test_proof_format_errors(env):
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env

    def build(p, r, enc):
        cv = [p.commit(x, r.randrange(L)) for x in (3, 4, 6, 1, 40)]
        v = [var for _, var in cv]
        gadgets.example_gadget(p, v[0], v[1], v[2], v[3], v[4], 9)
        return [enc(c) for c, _ in cv]

    (_, _), (pp, coms) = _both(env, b"R1CSExampleGadget", build, 5)

    def verifier():
        vf = P.Verifier(gens, P.Transcript(b"R1CSExampleGadget"))
        v = [vf.commit(c) for c in coms]
        gadgets.example_gadget(vf, v[0], v[1], v[2], v[3], v[4], 9)
        return vf

    for bad in (b"", pp[:-1], bytes([7]) + pp[1:], pp[:40]):
        with pytest.raises(P.FormatError):
            verifier().verify(bad)
    # a point that is not a valid ristretto encoding -> FormatError (proof.rs:141-145)
    broken = bytearray(pp)
    broken[1:33] = bytes.fromhex("01" + "00" * 31)
    with pytest.raises(P.FormatError):
        verifier().verify(bytes(broken))
    # identity commitment point -> VerificationError (transcript.rs:101-113)
    ident = bytearray(pp)
    ident[1:33] = bytes(32)
    with pytest.raises(P.VerificationError):
        verifier().verify(bytes(ident))


def test_generators_capacity_error(ctx):
    from mpc_bulletproof_b200 import protocol as P

    pc = O.PedersenGens()
    bp = O.BulletproofGens(2, 1)
    gens = P.Gens(ctx, points_bytes(bp.G(2)), points_bytes(bp.H(2)), pc.B.encode(), pc.B_blinding.encode())
    p = P.Prover(gens, P.Transcript(b"cap"))
    _, var = p.commit(5, 7)
    out = var
    for _ in range(3):
        _, _, out = p.multiply(out, out)
    with pytest.raises(P.InvalidGeneratorsLength):
        p.prove(1)


def test_square_chain_gadget(ctx, env):
    """bpg_gadget_square_chain (the reference's bench circuit, benches/r1cs.rs:24-32) is n calls of
    multiply(var, var): byte-identical proofs, and the oracle accepts them."""
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env
    n, val = 37, 0xDEADBEEF

    def loop(cs):
        var = cs.commit_public(val)
        for _ in range(n):
            _, _, var = cs.multiply(var, var)

    def native(cs):
        cs.square_chain(cs.commit_public(val), n)

    proofs = []
    for build in (loop, native):
        p = P.Prover(gens, P.Transcript(b"chain"))
        build(p)
        proofs.append(p.prove(99))
    assert proofs[0] == proofs[1]
    for build in (loop, native):
        v = P.Verifier(gens, P.Transcript(b"chain"))
        build(v)
        v.verify(proofs[0])
    ov = O.Verifier(pc, O.Transcript(b"chain"))
    var = ov.commit_public(val)
    for _ in range(n):
        _, _, var = ov.multiply(var, var)
    ov.verify(O.R1CSProof.from_bytes(proofs[0]), bp)


@pytest.mark.parametrize("n", [2, 8, 64])
def test_ipp_on_shared_windowed_table(ctx, n):
    """G and H as two ranges of ONE windowed table (a resident BulletproofGens) and an arbitrary Q:
    the round MSMs run over the caller's table, c_L Q and c_R Q come from a comb of Q.  Same proof
    bytes as the oracle; verification through the same table."""
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200 import protocol as P

    r = random.Random(5000 + n)
    bp = O.BulletproofGens(n, 1)
    Gs, Hs = bp.G(n), bp.H(n)
    Q = G.hash_to_group_sha512(b"shared table q")
    a = [r.randrange(L) for _ in range(n)]
    b = [r.randrange(L) for _ in range(n)]
    Gf = [r.randrange(1, L) for _ in range(n)]
    Hf = [r.randrange(1, L) for _ in range(n)]
    want = O.InnerProductProof.create(O.Transcript(b"shared"), Q, Gf, Hf, Gs, Hs, a, b)
    pad = [G.BASEPOINT] * 3  # ranges need not start at 0 nor be adjacent
    t = Table(ctx, points_bytes(pad + Gs + pad + Hs)).set_windows(0)
    g_off, h_off = 3, 3 + n + 3
    got = P.InnerProductProof.create(ctx, P.Transcript(b"shared"), Q.encode(), Gf, Hf, t, t, a, b, g_off=g_off, h_off=h_off)
    assert got.to_bytes() == want.to_bytes()
    c = O.inner_product(a, b)
    Ppt = G.msm([a[i] * Gf[i] % L for i in range(n)] + [b[i] * Hf[i] % L for i in range(n)] + [c], Gs + Hs + [Q])
    got.verify(ctx, n, P.Transcript(b"shared"), Gf, Hf, Ppt.encode(), Q.encode(), t, t, g_off=g_off, h_off=h_off)
    with pytest.raises(P.VerificationError):
        got.verify(ctx, n, P.Transcript(b"shared"), Gf, Hf, (Ppt + Q).encode(), Q.encode(), t, t, g_off=g_off, h_off=h_off)
    t.close()


def test_random_circuit_gadget(ctx, env):
    """BASELINE.json config 4 in miniature: random linear constraints over (a_L, a_R, a_O, v) with
    public constants (the constant term of every row exercises the device flattening's w_c path).
    The native gadget and the oracle's build the same system: byte-identical proofs."""
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env
    seed, n_mult, n_cons, m = 77, 23, 51, 3
    r = random.Random(9)
    vs = [r.randrange(L) for _ in range(m)]
    vb = [r.randrange(L) for _ in range(m)]
    # oracle
    op = O.Prover(pc, O.Transcript(b"rand"))
    Vs = [op.commit(v, b)[0] for v, b in zip(vs, vb)]
    c0 = gadgets.random_circuit(op, seed, n_mult, n_cons, v_assignment=vs)
    want = op.prove(bp, O.Blindings(31)).to_bytes()
    # product, native gadget
    p = P.Prover(gens, P.Transcript(b"rand"))
    got_V = [p.commit(v, b)[0] for v, b in zip(vs, vb)]
    assert got_V == [V.encode() for V in Vs]
    c0_native = p.random_circuit(seed, n_mult, n_cons)
    assert c0_native == b"".join(G.sc_to_bytes(k) for k in c0)
    proof = p.prove(31)
    assert proof == want
    # verifier: native gadget with the public constants; the oracle's verifier agrees
    v = P.Verifier(gens, P.Transcript(b"rand"))
    for V in got_V:
        v.commit(V)
    v.random_circuit(seed, n_mult, n_cons, c0_native)
    v.verify(proof)
    ov = O.Verifier(pc, O.Transcript(b"rand"))
    for V in Vs:
        ov.commit(V)
    gadgets.random_circuit(ov, seed, n_mult, n_cons, c0=c0)
    ov.verify(O.R1CSProof.from_bytes(proof), bp)
    # a wrong constant is a false statement
    bad = bytearray(c0_native)
    bad[0] ^= 1
    v = P.Verifier(gens, P.Transcript(b"rand"))
    for V in got_V:
        v.commit(V)
    v.random_circuit(seed, n_mult, n_cons, bytes(bad))
    with pytest.raises(P.VerificationError):
        v.verify(proof)


def test_full_size_r1cs_roundtrip(ctx):
    """BASELINE.json's 2^16-multiplier size (too large for the Python oracle) through the properties
    the reference's own tests use (tests/r1cs.rs: prove -> verify, serialization roundtrip, a false
    statement and a tampered proof are rejected), for the bench circuit and the random circuit."""
    import hashlib

    from mpc_bulletproof_b200 import Comb, protocol as P

    n = 1 << 16
    comb = Comb(ctx, G.BASEPOINT.encode())

    def pts(seed, count):
        ks = b"".join(hashlib.sha512(seed + i.to_bytes(4, "little")).digest()[:31] + b"\x00" for i in range(count))
        return comb.mul(ks)

    gens = P.Gens(ctx, pts(b"G", n), pts(b"H", n), G.BASEPOINT.encode(), pts(b"B", 1))
    p = P.Prover(gens, P.Transcript(b"full"))
    p.square_chain(p.commit_public(123456789), n)
    proof = p.prove(2024)
    assert len(proof) == 1 + 11 * 32 + (2 * 16 + 2) * 32  # one phase, lg n = 16 (proof.rs:82-108)
    p2 = P.Prover(gens, P.Transcript(b"full"))
    p2.square_chain(p2.commit_public(123456789), n)
    assert p2.prove(2024) == proof  # deterministic given the seed

    def verifier(val):
        v = P.Verifier(gens, P.Transcript(b"full"))
        v.square_chain(v.commit_public(val), n)
        return v

    verifier(123456789).verify(proof)
    with pytest.raises(P.VerificationError):
        verifier(123456790).verify(proof)
    bad = bytearray(proof)
    bad[500] ^= 0x10
    with pytest.raises((P.VerificationError, P.FormatError)):
        verifier(123456789).verify(bytes(bad))
    # random circuit: 2^15 multipliers, 2^16 constraints, 4 commitments
    pr = P.Prover(gens, P.Transcript(b"full rand"))
    Vs = [pr.commit(1000 + j, 77 + j)[0] for j in range(4)]
    c0 = pr.random_circuit(11, 1 << 15, 1 << 16)
    proof2 = pr.prove(5)
    vr = P.Verifier(gens, P.Transcript(b"full rand"))
    for V in Vs:
        vr.commit(V)
    vr.random_circuit(11, 1 << 15, 1 << 16, c0)
    vr.verify(proof2)
    gens.close()
    comb.close()


# ------------------------------------------------------------------ transcript-bound blinding (prover.rs:435-445)
@pytest.mark.parametrize("k", [1, 4, 7])
def test_keyed_blinding_matches_oracle(env, k):
    """The production blinding path -- merlin TranscriptRng (transcript + v_blindings + 32 external bytes)
    in the reference's draw order, s_L / s_R expanded from one draw through ChaCha20 on the device --
    gives the oracle's proof bytes when the 32 external bytes are fixed; shuffle gadget, so two phases
    (two vector keys) and, for k = 7, the padding path."""
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env
    label = b"ShuffleProofTest"
    rb = bytes((37 * i + k) & 0xFF for i in range(32))

    def build(p, r, enc):
        inp = [r.randrange(2**64) for _ in range(k)]
        outp = inp[:]
        r.shuffle(outp)
        ic = [p.commit(v, r.randrange(L)) for v in inp]
        oc = [p.commit(v, r.randrange(L)) for v in outp]
        gadgets.shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
        return [enc(c) for c, _ in ic + oc]

    op = O.Prover(pc, O.Transcript(label))
    ocoms = build(op, random.Random(900 + k), lambda c: c.encode())
    want = op.prove(bp, O.TranscriptBlindings(rb)).to_bytes()
    pp = P.Prover(gens, P.Transcript(label))
    pcoms = build(pp, random.Random(900 + k), lambda c: c)
    got = pp.prove(rng_bytes=rb)
    assert ocoms == pcoms and got == want, "keyed-blinding proof bytes differ from the oracle"
    # other external bytes -> other blindings -> another proof of the same statement
    p2 = P.Prover(gens, P.Transcript(label))
    build(p2, random.Random(900 + k), lambda c: c)
    assert p2.prove(rng_bytes=bytes(32)) != got

    def verifier(cls, tr):
        vf = cls(gens if cls is P.Verifier else pc, tr(label))
        vs = [vf.commit(c if cls is P.Verifier else G.decode(c)) for c in pcoms]
        gadgets.shuffle_gadget(vf, vs[:k], vs[k:])
        return vf

    verifier(P.Verifier, P.Transcript).verify(got)
    verifier(O.Verifier, O.Transcript).verify(O.R1CSProof.from_bytes(got), bp)


def test_os_entropy_prove_and_hardened_verify(env):
    """bpg_prover_prove (operating-system entropy): two proofs of one statement differ and both verify,
    under the fork's deterministic r (verifier.rs:506) and under the hardened verifier whose r the
    prover cannot predict; a false statement is rejected by both; the hardened r equals the oracle's."""
    from mpc_bulletproof_b200 import protocol as P

    pc, bp, gens = env

    def stmt(cs, commit, c2):
        cv = [commit(cs, x) for x in (3, 4, 6, 1, 40)]
        gadgets.example_gadget(cs, cv[0], cv[1], cv[2], cv[3], cv[4], c2)

    proofs, coms = [], None
    for _ in range(2):
        p = P.Prover(gens, P.Transcript(b"R1CSExampleGadget"))
        cs_coms = []

        def commit(cs, x):
            c, var = cs.commit(x, 12345 + x)
            cs_coms.append(c)
            return var

        stmt(p, commit, 9)
        proofs.append(p.prove())
        coms = cs_coms
    assert proofs[0] != proofs[1]

    def verifier(c2):
        vf = P.Verifier(gens, P.Transcript(b"R1CSExampleGadget"))
        it = iter(coms)
        stmt(vf, lambda cs, x: cs.commit(next(it)), c2)
        return vf

    for pr in proofs:
        verifier(9).verify(pr)
        verifier(9).verify(pr, hardened=True)
        verifier(9).verify(pr, rng_bytes=bytes(range(32)))
        for kw in ({}, {"hardened": True}):
            with pytest.raises(P.VerificationError):
                verifier(10).verify(pr, **kw)
    # the oracle's hardened verifier (same TranscriptRng draw) accepts the same proof
    ov = O.Verifier(pc, O.Transcript(b"R1CSExampleGadget"))
    it = iter(coms)
    stmt(ov, lambda cs, x: cs.commit(G.decode(next(it))), 9)
    ov.verify(O.R1CSProof.from_bytes(proofs[0]), bp, rng_bytes=bytes(range(32)))


def test_config1_shuffle_k64_product_gadget(ctx):
    """BASELINE.json config 1 (benches/shuffle.rs, k = 64) through the product's own gadget module
    (mpc_bulletproof_b200/gadgets.py): same proof bytes as the oracle's prover running the oracle's gadget,
    the product verifier and the oracle verifier accept, a non-permutation is rejected."""
    from mpc_bulletproof_b200 import gadgets as PG
    from mpc_bulletproof_b200 import protocol as P

    k = 64
    pc = O.PedersenGens()
    bp = O.BulletproofGens(128, 1)
    gens = P.Gens(ctx, points_bytes(bp.G(128)), points_bytes(bp.H(128)), pc.B.encode(), pc.B_blinding.encode())
    r = random.Random(64)
    inp = [r.randrange(2**64) for _ in range(k)]
    outp = inp[:]
    r.shuffle(outp)
    blinds = [r.randrange(L) for _ in range(2 * k)]
    proof, ic, oc = PG.shuffle_prove(gens, P.Transcript, b"ShuffleBenchmark", inp, outp, blinds, rng_seed=640)
    otr = O.Transcript(b"ShuffleBenchmark")
    otr.append_message(b"dom-sep", b"ShuffleProof")
    otr.append_u64(b"k", k)
    op = O.Prover(pc, otr)
    oic = [op.commit(v, blinds[i]) for i, v in enumerate(inp)]
    ooc = [op.commit(v, blinds[k + i]) for i, v in enumerate(outp)]
    gadgets.shuffle_gadget(op, [v for _, v in oic], [v for _, v in ooc])
    want = op.prove(bp, O.Blindings(640)).to_bytes()
    assert proof == want and ic == [c.encode() for c, _ in oic]
    PG.shuffle_verify(gens, P.Transcript, b"ShuffleBenchmark", proof, ic, oc)
    bad_out = outp[:]
    bad_out[3] += 1
    bad, bic, boc = PG.shuffle_prove(gens, P.Transcript, b"ShuffleBenchmark", inp, bad_out, blinds, rng_seed=641)
    with pytest.raises(P.VerificationError):
        PG.shuffle_verify(gens, P.Transcript, b"ShuffleBenchmark", bad, bic, boc)
    # the native gadget (bpg_gadget_shuffle) at the small sizes with their own code paths (k = 1: one constraint,
    # k = 2: no chain), and against the callback form of the same gadget: identical proof bytes
    for ks in (1, 2, 3, 5):
        inp = [r.randrange(2**64) for _ in range(ks)]
        outp = inp[:]
        r.shuffle(outp)
        bl = [r.randrange(L) for _ in range(2 * ks)]
        proof, ic, oc = PG.shuffle_prove(gens, P.Transcript, b"ShuffleBenchmark", inp, outp, bl, rng_seed=700 + ks)
        otr = O.Transcript(b"ShuffleBenchmark")
        otr.append_message(b"dom-sep", b"ShuffleProof")
        otr.append_u64(b"k", ks)
        op = O.Prover(pc, otr)
        oic = [op.commit(v, bl[i]) for i, v in enumerate(inp)]
        ooc = [op.commit(v, bl[ks + i]) for i, v in enumerate(outp)]
        gadgets.shuffle_gadget(op, [v for _, v in oic], [v for _, v in ooc])
        assert proof == op.prove(bp, O.Blindings(700 + ks)).to_bytes(), ks
        PG.shuffle_verify(gens, P.Transcript, b"ShuffleBenchmark", proof, ic, oc)
        tr = P.Transcript(b"ShuffleBenchmark")
        tr.append_message(b"dom-sep", b"ShuffleProof")
        tr.append_u64(b"k", ks)
        p2 = P.Prover(gens, tr)
        cv = p2.commit_batch(list(inp) + list(outp), bl)
        PG.shuffle_gadget_callbacks(p2, [v for _, v in cv[:ks]], [v for _, v in cv[ks:]])
        assert p2.prove(700 + ks) == proof, ks
    gens.close()


def test_verification_scalars_directly(ctx):
    """`InnerProductProof::verification_scalars` (reference src/inner_product_proof.rs:254-310): the n-vector s that the
    device forms by its closed form, checked VALUE BY VALUE against the oracle's serial recurrence rather than
    through accept/reject: with (a, b) = (1, 0) and no other term, bpg_ipp_verify_msm returns sum_i s_i g_i G_i, and
    with (0, 1) it returns sum_i s_{N-1-i} h_i H_i (= 1/s_i); distinct random generators make the sums injective
    in s for all practical purposes."""
    import ctypes

    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200._lib import lib
    from mpc_bulletproof_b200.protocol import _raise

    class Params(ctypes.Structure):
        _fields_ = [("u_sq", (ctypes.c_uint32 * 8) * 32), ("allinv", ctypes.c_uint32 * 8), ("a", ctypes.c_uint32 * 8),
                    ("b", ctypes.c_uint32 * 8), ("lg_n", ctypes.c_uint32), ("N", ctypes.c_uint32)]

    def mont(x):
        v = x * (1 << 256) % L
        return (ctypes.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])

    for n in (1, 2, 16, 64):
        r = random.Random(500 + n)
        lg = n.bit_length() - 1
        bp = O.BulletproofGens(n, 1)
        Gs, Hs = bp.G(n), bp.H(n)
        # a proof-shaped object whose L, R are arbitrary valid points: only the challenges matter here
        Ls = [rand_pt(r) for _ in range(lg)]
        Rs = [rand_pt(r) for _ in range(lg)]
        proof = O.InnerProductProof(Ls, Rs, 1, 0)
        u_sq, u_inv_sq, s = proof.verification_scalars(n, O.Transcript(b"vs"))
        allinv = s[0] if n > 1 else 1
        Gf = [r.randrange(1, L) for _ in range(n)]
        Hf = [r.randrange(1, L) for _ in range(n)]
        t = Table(ctx, points_bytes(Gs + Hs))
        for a, b in ((1, 0), (0, 1), (r.randrange(L), r.randrange(L))):
            p = Params()
            for j in range(lg):
                p.u_sq[j] = mont(u_sq[j])
            p.allinv, p.a, p.b, p.lg_n, p.N = mont(allinv), mont(a), mont(b), lg, n
            out = ctypes.create_string_buffer(32)
            _raise(lib().bpg_ipp_verify_msm(ctx._h, t._h, 0, t._h, n, None, None, 0, scalars_bytes(Gf), scalars_bytes(Hf),
                                            ctypes.byref(p), out))
            want = G.msm([a * s[i] * Gf[i] % L for i in range(n)] + [b * s[n - 1 - i] * Hf[i] % L for i in range(n)], Gs + Hs)
            assert out.raw == want.encode(), (n, a, b)
        t.close()


def rand_pt(r):
    return r.randrange(1, L) * G.BASEPOINT
