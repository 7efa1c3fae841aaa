"""The oracles (big-integer and C) against the committed golden fixtures; the fixtures
themselves are reproducible with tests/golden/make_golden.py."""
import json
import os

from oracle import cbind
from oracle import group as G
from oracle import protocol as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


def test_msm_golden_oracles():
    for v in load("msm.json"):
        ks = [int.from_bytes(bytes.fromhex(s), "little") for s in v["scalars"]]
        ps = [G.decode(bytes.fromhex(p)) for p in v["points"]]
        assert G.msm_naive(ks, ps).encode().hex() == v["result"], v["name"]
        got = cbind.msm(b"".join(bytes.fromhex(s) for s in v["scalars"]), b"".join(bytes.fromhex(p) for p in v["points"]))
        assert got.hex() == v["result"], v["name"]


def test_ipp_golden_oracle_verifies():
    for v in load("ipp.json"):
        n = v["n"]
        bp = O.BulletproofGens(n, 1)
        proof = O.InnerProductProof.from_bytes(bytes.fromhex(v["proof"]))
        Gf = [int.from_bytes(bytes.fromhex(x), "little") for x in v["G_factors"]]
        Hf = [int.from_bytes(bytes.fromhex(x), "little") for x in v["H_factors"]]
        proof.verify(n, O.Transcript(v["label"].encode()), Gf, Hf, G.decode(bytes.fromhex(v["P"])), G.decode(bytes.fromhex(v["Q"])), bp.G(n), bp.H(n))


def test_generator_golden():
    g = load("r1cs.json")["generators_first8"]
    pc, bp = O.PedersenGens(), O.BulletproofGens(8, 1)
    assert pc.B.encode().hex() == g["B"] and pc.B_blinding.encode().hex() == g["B_blinding"]
    assert [x.encode().hex() for x in bp.G(8)] == g["G"] and [x.encode().hex() for x in bp.H(8)] == g["H"]
    # values recorded in SURVEY.md Appendix B.2
    assert g["G"][0] == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"
    assert g["H"][1] == "acf2d2b95428fac99b12da3bab92edf8ea3788c2fd16769e586397eede7b5052"
    assert g["B_blinding"] == "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"
