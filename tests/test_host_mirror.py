"""Host-side pieces of the product that need no GPU: the C++ Merlin transcript (bound
through the C ABI) against merlin's published test vector and against the oracle."""
from oracle import protocol as O


def test_transcript_equivalence_simple_vector():
    from mpc_bulletproof_b200 import protocol as P

    t = P.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_transcript_matches_oracle():
    from mpc_bulletproof_b200 import protocol as P

    t, o = P.Transcript(b"proto"), O.Transcript(b"proto")
    big = bytes(range(256)) * 3  # crosses the STROBE rate boundary
    for i in range(6):
        t.append_u64(b"n", i)
        o.append_u64(b"n", i)
        t.append_message(b"blob", big[: 50 * i])
        o.append_message(b"blob", big[: 50 * i])
        assert t.challenge_scalar(b"c") == o.challenge_scalar(b"c")
        assert t.challenge_bytes(b"bytes", 200) == o.challenge_bytes(b"bytes", 200)
    c = t.clone()
    assert c.challenge_scalar(b"x") == t.challenge_scalar(b"x") == o.challenge_scalar(b"x")


def test_transcript_rng_matches_oracle():
    """merlin TranscriptRng (the prover's blinding source, reference src/r1cs/prover.rs:435-445):
    host mirror against the oracle's restatement, across the STROBE rate boundary, and bound to the
    transcript, the witness and the external bytes."""
    from mpc_bulletproof_b200 import protocol as P

    t, o = P.Transcript(b"rng"), O.Transcript(b"rng")
    for tr in (t, o):
        tr.append_message(b"stmt", bytes(range(200)))
    w1, w2, ext = bytes(range(32)), bytes(range(100, 132)), bytes([7] * 32)
    r = t.build_rng().rekey_with_witness_bytes(b"v_blinding", w1).rekey_with_witness_bytes(b"v_blinding", w2).finalize(ext)
    q = o.build_rng().rekey_with_witness_bytes(b"v_blinding", w1).rekey_with_witness_bytes(b"v_blinding", w2).finalize(ext)
    for n in (64, 32, 1, 400, 64):
        assert r.fill_bytes(n) == q.fill_bytes(n)
    # the fork leaves the transcript itself untouched
    assert t.challenge_bytes(b"c", 32) == o.challenge_bytes(b"c", 32)
    base = o.build_rng().rekey_with_witness_bytes(b"v_blinding", w1).finalize(ext).fill_bytes(64)
    assert base != o.build_rng().rekey_with_witness_bytes(b"v_blinding", w2).finalize(ext).fill_bytes(64)
    assert base != o.build_rng().rekey_with_witness_bytes(b"v_blinding", w1).finalize(bytes(32)).fill_bytes(64)
    o2 = O.Transcript(b"rng")
    assert base != o2.build_rng().rekey_with_witness_bytes(b"v_blinding", w1).finalize(ext).fill_bytes(64)


def test_fork_keccak_and_hash_to_scalar():
    """legacy Keccak-256 (known answer: keccak256("") = c5d246...a470) and the fork's
    hash_to_scalar (reference src/util.rs:252-267): host mirror against the oracle."""
    import random

    from mpc_bulletproof_b200 import stark as PS
    from oracle import stark as S

    assert PS.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    r = random.Random(12)
    from oracle.merlin import keccak256

    for n in (1, 31, 32, 135, 136, 137, 300):
        data = bytes(r.getrandbits(8) for _ in range(n))
        assert PS.keccak256(data) == keccak256(data)
    for low in [bytes(32), bytes([255] * 32)] + [bytes(r.getrandbits(8) for _ in range(32)) for _ in range(20)]:
        assert PS.hash_to_scalar(low) == S.hash_to_scalar(low) < S.N
