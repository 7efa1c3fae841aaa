"""Workload gadgets of the reference's benches over the host mirror's constraint system
(`protocol.Prover` / `protocol.Verifier`): what a user of the library writes.  Linear combinations are
`LC` objects ({variable: coefficient}); variables are the tuples of `protocol.py`."""
from __future__ import annotations

from .protocol import L


class LC:
    """reference src/r1cs/linear_combination.rs:119-121"""

    def __init__(self, terms=None):
        self.terms = dict(terms or {})

    @staticmethod
    def of(x) -> "LC":
        if isinstance(x, LC):
            return x
        if isinstance(x, tuple):
            return LC({x: 1})
        return LC({("1",): int(x) % L})

    def __add__(self, o):
        r = LC(self.terms)
        for v, c in LC.of(o).terms.items():
            r.terms[v] = (r.terms.get(v, 0) + c) % L
        return r

    def __sub__(self, o):
        r = LC(self.terms)
        for v, c in LC.of(o).terms.items():
            r.terms[v] = (r.terms.get(v, 0) - c) % L
        return r


def shuffle_gadget(cs, x, y):
    """reference benches/shuffle.rs:30-69 (tests/r1cs.rs:22-63): y is a permutation of x, through the
    randomized product check prod (x_i - z) = prod (y_i - z).  The library's native form of the gadget
    (bpg_gadget_shuffle); `shuffle_gadget_callbacks` builds the same system through the callback interface."""
    cs.shuffle(x, y)


def shuffle_gadget_callbacks(cs, x, y):
    """the same gadget through `specify_randomized_constraints` and per-multiplier calls (what a caller's own
    gadget code does)"""
    assert len(x) == len(y)
    k = len(x)
    if k == 1:
        cs.constrain(LC.of(y[0]) - x[0])
        return

    def cb(cs):
        z = cs.challenge_scalar(b"shuffle challenge")

        def product(v):
            _, _, out = cs.multiply(LC.of(v[k - 1]) - z, LC.of(v[k - 2]) - z)
            for i in reversed(range(k - 2)):
                _, _, out = cs.multiply(LC.of(out), LC.of(v[i]) - z)
            return out

        cs.constrain(LC.of(product(x)) - product(y))

    cs.specify_randomized_constraints(cb)


def shuffle_prove(gens, transcript_cls, label: bytes, inputs, outputs, blindings, rng_seed=None, rng_bytes=None):
    """benches/shuffle.rs:76-116: returns (proof bytes, input commitments, output commitments)"""
    from .protocol import Prover

    k = len(inputs)
    tr = transcript_cls(label)
    tr.append_message(b"dom-sep", b"ShuffleProof")
    tr.append_u64(b"k", k)
    p = Prover(gens, tr)
    cv = p.commit_batch(list(inputs) + list(outputs), blindings[: 2 * k])  # 2k commits, one launch
    ic, oc = cv[:k], cv[k:]
    shuffle_gadget(p, [v for _, v in ic], [v for _, v in oc])
    proof = p.prove(rng_seed, rng_bytes)
    return proof, [c for c, _ in ic], [c for c, _ in oc]


def shuffle_verify(gens, transcript_cls, label: bytes, proof: bytes, in_coms, out_coms):
    """benches/shuffle.rs:120-145"""
    from .protocol import Verifier

    k = len(in_coms)
    tr = transcript_cls(label)
    tr.append_message(b"dom-sep", b"ShuffleProof")
    tr.append_u64(b"k", k)
    v = Verifier(gens, tr)
    xs = [v.commit(c) for c in in_coms]
    ys = [v.commit(c) for c in out_coms]
    shuffle_gadget(v, xs, ys)
    v.verify(proof)
