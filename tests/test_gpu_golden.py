"""The CUDA path against the committed golden fixtures (tests/golden/*.json)."""
import json
import os

import pytest

from oracle import gadgets
from oracle import group as G
from oracle import protocol as O
from tests.util import points_bytes

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


def test_msm_golden(ctx):
    from mpc_bulletproof_b200 import msm

    for v in load("msm.json"):
        got = msm(ctx, b"".join(bytes.fromhex(s) for s in v["scalars"]), b"".join(bytes.fromhex(p) for p in v["points"]))
        assert got.hex() == v["result"], v["name"]


def test_ipp_golden(ctx):
    from mpc_bulletproof_b200 import Table
    from mpc_bulletproof_b200 import protocol as P

    for v in load("ipp.json"):
        n = v["n"]
        bp = O.BulletproofGens(n, 1)
        tG, tH = Table(ctx, points_bytes(bp.G(n))), Table(ctx, points_bytes(bp.H(n)))
        dec = lambda xs: [int.from_bytes(bytes.fromhex(x), "little") for x in xs]  # noqa: E731
        got = P.InnerProductProof.create(
            ctx, P.Transcript(v["label"].encode()), bytes.fromhex(v["Q"]), dec(v["G_factors"]), dec(v["H_factors"]), tG, tH, dec(v["a"]), dec(v["b"])
        )
        assert got.to_bytes().hex() == v["proof"], n
        got.verify(ctx, n, P.Transcript(v["label"].encode()), dec(v["G_factors"]), dec(v["H_factors"]), bytes.fromhex(v["P"]), bytes.fromhex(v["Q"]), tG, tH)


def test_r1cs_golden(ctx):
    from mpc_bulletproof_b200 import protocol as P

    pc, bp = O.PedersenGens(), O.BulletproofGens(128, 1)
    gens = P.Gens(ctx, points_bytes(bp.G(128)), points_bytes(bp.H(128)), pc.B.encode(), pc.B_blinding.encode())
    for v in load("r1cs.json")["proofs"]:
        blind = [int.from_bytes(bytes.fromhex(x), "little") for x in v["v_blinding"]]
        if v["kind"] == "example":
            label = b"R1CSExampleGadget"
            p = P.Prover(gens, P.Transcript(label))
            cv = [p.commit(x, bl) for x, bl in zip(v["values"], blind)]
            vs = [var for _, var in cv]
            gadgets.example_gadget(p, vs[0], vs[1], vs[2], vs[3], vs[4], v["c2"])
            vf = P.Verifier(gens, P.Transcript(label))
            vv = [vf.commit(bytes.fromhex(c)) for c in v["commitments"]]
            gadgets.example_gadget(vf, vv[0], vv[1], vv[2], vv[3], vv[4], v["c2"])
        else:
            k = v["k"]
            t = P.Transcript(b"ShuffleProofTest")
            t.append_message(b"dom-sep", b"ShuffleProof")
            t.append_u64(b"k", k)
            p = P.Prover(gens, t)
            cv = [p.commit(x, bl) for x, bl in zip(v["values"], blind)]
            gadgets.shuffle_gadget(p, [x for _, x in cv[:k]], [x for _, x in cv[k:]])
            t2 = P.Transcript(b"ShuffleProofTest")
            t2.append_message(b"dom-sep", b"ShuffleProof")
            t2.append_u64(b"k", k)
            vf = P.Verifier(gens, t2)
            vv = [vf.commit(bytes.fromhex(c)) for c in v["commitments"]]
            gadgets.shuffle_gadget(vf, vv[:k], vv[k:])
        assert [c.hex() for c, _ in cv] == v["commitments"]
        assert p.prove(v["seed"]).hex() == v["proof"], v["kind"]
        if v["valid"]:
            vf.verify(bytes.fromhex(v["proof"]))
        else:
            with pytest.raises(P.VerificationError):
                vf.verify(bytes.fromhex(v["proof"]))
