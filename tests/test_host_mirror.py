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
