"""ORACLE (test infrastructure).  The reference's workload gadgets, restated over
any constraint system exposing multiply / allocate_multiplier / constrain /
specify_randomized_constraints / challenge_scalar (oracle.protocol.Prover/Verifier
and the host mirror of the product share that interface)."""
from __future__ import annotations

from .protocol import LC, ONE, L, _lc


def shuffle_gadget(cs, x, y):
    """reference tests/r1cs.rs:22-63 / benches/shuffle.rs:30-69"""
    assert len(x) == len(y)
    k = len(x)
    if k == 1:
        cs.constrain(_lc(y[0]) - x[0])
        return

    def cb(cs):
        z = cs.challenge_scalar(b"shuffle challenge")
        _, _, last_x = cs.multiply(_lc(x[k - 1]) - z, _lc(x[k - 2]) - z)
        out = last_x
        for i in reversed(range(k - 2)):
            _, _, out = cs.multiply(_lc(out), _lc(x[i]) - z)
        first_x = out
        _, _, last_y = cs.multiply(_lc(y[k - 1]) - z, _lc(y[k - 2]) - z)
        out = last_y
        for i in reversed(range(k - 2)):
            _, _, out = cs.multiply(_lc(out), _lc(y[i]) - z)
        cs.constrain(_lc(first_x) - out)

    cs.specify_randomized_constraints(cb)


def example_gadget(cs, a1, a2, b1, b2, c1, c2):
    """reference tests/r1cs.rs:217-228: (a1 + a2) * (b1 + b2) = (c1 + c2)"""
    _, _, c_var = cs.multiply(_lc(a1) + a2, _lc(b1) + b2)
    cs.constrain(_lc(c1) + c2 - c_var)


def range_proof_gadget(cs, v, v_assignment, n):
    """reference tests/r1cs.rs:620-652: v in [0, 2^n)"""
    v = _lc(v)
    exp_2 = 1
    for i in range(n):
        if v_assignment is None:
            a, b, o = cs.allocate_multiplier(None, None)
        else:
            bit = (v_assignment >> i) & 1
            a, b, o = cs.allocate_multiplier(1 - bit, bit)
        cs.constrain(_lc(o))
        cs.constrain(_lc(a) + (_lc(b) - 1))
        v = v - _lc(b).scale(exp_2)
        exp_2 = exp_2 * 2 % L
    cs.constrain(v)


def dummy_circuit(cs, n_constraints, val):
    """reference benches/r1cs.rs:24-32: commit_public(val) then a chain of n squarings"""
    var = cs.commit_public(val)
    for _ in range(n_constraints):
        _, _, var = cs.multiply(_lc(var), _lc(var))
