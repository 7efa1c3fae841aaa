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


def random_circuit(cs, seed, n_mult, n_cons, v_assignment=None, c0=None):
    """SURVEY.md §8d config 4 (BASELINE.json "synthetic random R1CS circuit"): n_mult multipliers
    with uniform a_L, a_R (a_O = a_L o a_R) and n_cons linear constraints
        c1 a_L[i1] + c2 a_R[i2] + c3 a_O[i3] + c4 v[j] - c0 = 0
    over the m variables already committed, indices and coefficients from xoshiro256**(seed),
    c0 fixed from the witness so that the system is satisfied.  The prover passes the committed
    values and gets the public constants c0 back; the verifier passes those constants.
    The draw order is part of the definition (the product's native gadget follows it)."""
    from .protocol import Xoshiro256ss

    rng = Xoshiro256ss(seed)
    proving = v_assignment is not None
    m = len(v_assignment) if proving else len(cs.V)
    assert m >= 1
    aL, aR, aO, vars_ = [], [], [], []
    for _ in range(n_mult):
        l, r = rng.scalar(), rng.scalar()
        aL.append(l)
        aR.append(r)
        aO.append(l * r % L)
        vars_.append(cs.allocate_multiplier(l, r) if proving else cs.allocate_multiplier(None, None))
    out = []
    for q in range(n_cons):
        i1, i2, i3 = rng.next_u64() % n_mult, rng.next_u64() % n_mult, rng.next_u64() % n_mult
        j = rng.next_u64() % m
        c1, c2, c3, c4 = rng.scalar(), rng.scalar(), rng.scalar(), rng.scalar()
        if proving:
            k0 = (c1 * aL[i1] + c2 * aR[i2] + c3 * aO[i3] + c4 * v_assignment[j]) % L
        else:
            k0 = c0[q]
        out.append(k0)
        lc = LC([(vars_[i1][0], c1), (vars_[i2][1], c2), (vars_[i3][2], c3), (("V", j), c4), (ONE, (-k0) % L)])
        cs.constrain(lc)
    return out
