"""Generates tests/golden/*.json from the oracle (oracle/group.py, oracle/protocol.py).

The reference holds no golden vector for any group-valued output and
cannot be built here (no Rust toolchain), so the fixtures are outputs of the oracle —
itself pinned by RFC 9496 / Merlin known-answer vectors (tests/test_oracle_group.py,
tests/test_host_mirror.py).  Re-run with:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gadgets  # noqa: E402
from oracle import group as G  # noqa: E402
from oracle import protocol as O  # noqa: E402

L = G.L
HERE = os.path.dirname(os.path.abspath(__file__))


def msm_vectors():
    out = []
    for name, n, seed in [("n1", 1, 1), ("n2", 2, 2), ("n17", 17, 3), ("n64", 64, 4), ("n200", 200, 5)]:
        r = random.Random(seed)
        ks = [r.randrange(L) for _ in range(n)]
        ms = [r.randrange(1, L) for _ in range(n)]
        ps = [m * G.BASEPOINT for m in ms]
        out.append(
            dict(name=name, scalars=[G.sc_to_bytes(k).hex() for k in ks], points=[p.encode().hex() for p in ps], result=G.msm(ks, ps).encode().hex())
        )
    # edge vectors of SURVEY.md §8c
    r = random.Random(9)
    p, q = r.randrange(1, L) * G.BASEPOINT, r.randrange(1, L) * G.BASEPOINT
    k1, k2 = r.randrange(L), r.randrange(L)
    edge = [
        ("zero_scalar", [0, k1], [p, q]),
        ("scalar_one", [1], [p]),
        ("scalar_l_minus_1", [L - 1], [p]),
        ("scalar_ge_2_252", [2**252, 2**252 + 5, L - 2], [p, q, p]),
        ("identity_point", [k1, k2], [G.IDENTITY, p]),
        ("duplicate_points", [k1, k2], [p, p]),
        ("p_minus_p", [k1, k1], [p, -p]),
        ("all_one_bucket", [k1] * 40, [(i + 1) * G.BASEPOINT for i in range(40)]),
        ("empty", [], []),
    ]
    for name, ks, ps in edge:
        out.append(
            dict(name=name, scalars=[G.sc_to_bytes(k).hex() for k in ks], points=[x.encode().hex() for x in ps], result=G.msm(ks, ps).encode().hex())
        )
    return out


def ipp_vectors():
    out = []
    for n in (1, 2, 4, 32):
        r = random.Random(4000 + n)
        bp = O.BulletproofGens(n, 1)
        Q = G.hash_to_group_sha512(b"test point")
        a = [r.randrange(L) for _ in range(n)]
        b = [r.randrange(L) for _ in range(n)]
        y_inv = r.randrange(1, L)
        Gf = [1] * n
        Hf = [pow(y_inv, i, L) for i in range(n)]
        c = O.inner_product(a, b)
        P = G.msm(a + [b[i] * Hf[i] % L for i in range(n)] + [c], bp.G(n) + bp.H(n) + [Q])
        proof = O.InnerProductProof.create(O.Transcript(b"innerproducttest"), Q, Gf, Hf, bp.G(n), bp.H(n), a, b)
        out.append(
            dict(
                n=n, label="innerproducttest", Q=Q.encode().hex(), P=P.encode().hex(),
                a=[G.sc_to_bytes(x).hex() for x in a], b=[G.sc_to_bytes(x).hex() for x in b],
                G_factors=[G.sc_to_bytes(x).hex() for x in Gf], H_factors=[G.sc_to_bytes(x).hex() for x in Hf],
                proof=proof.to_bytes().hex(),
            )
        )
    return out


def r1cs_vectors():
    pc, bp = O.PedersenGens(), O.BulletproofGens(128, 1)
    out = []
    # example gadget (tests/r1cs.rs:542-563)
    for c2 in (9, 10):
        seed = 50 + c2
        r = random.Random(seed)
        p = O.Prover(pc, O.Transcript(b"R1CSExampleGadget"))
        blind = [r.randrange(L) for _ in range(5)]
        cv = [p.commit(x, bl) for x, bl in zip((3, 4, 6, 1, 40), blind)]
        v = [var for _, var in cv]
        gadgets.example_gadget(p, v[0], v[1], v[2], v[3], v[4], c2)
        proof = p.prove(bp, O.Blindings(seed))
        out.append(
            dict(
                kind="example", c2=c2, seed=seed, values=[3, 4, 6, 1, 40], v_blinding=[G.sc_to_bytes(x).hex() for x in blind],
                commitments=[c.encode().hex() for c, _ in cv], proof=proof.to_bytes().hex(), valid=(c2 == 9),
            )
        )
    # shuffle gadget k = 3 (two-phase, padded) and k = 1 (one-phase)
    for k in (1, 3):
        seed = 70 + k
        r = random.Random(seed)
        inp = [r.randrange(2**64) for _ in range(k)]
        outp = inp[:]
        r.shuffle(outp)
        blind = [r.randrange(L) for _ in range(2 * k)]
        t = O.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        p = O.Prover(pc, t)
        cv = [p.commit(x, bl) for x, bl in zip(inp + outp, blind)]
        gadgets.shuffle_gadget(p, [v for _, v in cv[:k]], [v for _, v in cv[k:]])
        proof = p.prove(bp, O.Blindings(seed))
        out.append(
            dict(
                kind="shuffle", k=k, seed=seed, values=inp + outp, v_blinding=[G.sc_to_bytes(x).hex() for x in blind],
                commitments=[c.encode().hex() for c, _ in cv], proof=proof.to_bytes().hex(), valid=True,
            )
        )
    gens = dict(
        B=pc.B.encode().hex(), B_blinding=pc.B_blinding.encode().hex(),
        G=[x.encode().hex() for x in bp.G(8)], H=[x.encode().hex() for x in bp.H(8)],
    )
    return dict(generators_first8=gens, proofs=out)


if __name__ == "__main__":
    for name, fn in (("msm.json", msm_vectors), ("ipp.json", ipp_vectors), ("r1cs.json", r1cs_vectors)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(), f, indent=0)
        print("wrote", name)
