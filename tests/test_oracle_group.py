"""Pins oracle/group.py against the known-answer vectors of RFC 9496 (ristretto255)
and against libsodium at the Edwards level.  The reference holds no golden vector
for any group-valued output, so these published vectors are
what anchors the whole parity chain: RFC vectors -> oracle/group.py -> oracle/c ->
CUDA path."""
import hashlib

import pytest

from oracle import group as G

# RFC 9496 Appendix A.1: multiples 0..15 of the generator
MULTIPLES = """
0000000000000000000000000000000000000000000000000000000000000000
e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76
6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919
94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259
da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57
e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e
f64746d3c92b13050ed8d80236a7f0007c3b3f962f5ba793d19a601ebb1df403
44f53520926ec81fbd5a387845beb7df85a96a24ece18738bdcfa6a7822a176d
903293d8f2287ebe10e2374dc1a53e0bc887e592699f02d077d5263cdd55601c
02622ace8f7303a31cafc63f8fc48fdc16e1c8c8d234b2f0d6685282a9076031
20706fd788b2720a1ed2a5dad4952b01f413bcf0e7564de8cdc816689e2db95f
bce83f8ba5dd2fa572864c24ba1810f9522bc6004afe95877ac73241cafdab42
e4549ee16b9aa03099ca208c67adafcafa4c3f3e4e5303de6026e3ca8ff84460
aa52e000df2e16f55fb1032fc33bc42742dad6bd5a8fc0be0167436c5948501f
46376b80f409b29dc2b5f6f0c52591990896e5716f41477cd30085ab7f10301e
e0c418f7c8d9c4cdd7395b93ea124f3ad99021bb681dfc3302a9d99a2e53e64e
""".split()

# RFC 9496 Appendix A.2: encodings that must be rejected
BAD = """
00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff
ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f
f3ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f
edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f
0100000000000000000000000000000000000000000000000000000000000000
01ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f
ed57ffd8c914fb201471d1c3d245ce3c746fcbe63a3679d51b6a516ebebe0e20
c34c4e1826e5d403b78e246e88aa051c36ccf0aafebffe137d148a2bf9104562
c940e5a4404157cfb1628b108db051a8d439e1a421394ec4ebccb9ec92a8ac78
47cfc5497c53dc8e61c91d17fd626ffb1c49e2bca94eed052281b510b1117a24
f1c6165d33367351b0da8f6e4511010c68174a03b6581212c71c0e1d026c3c72
87260f7a2f12495118360f02c26a470f450dadf34a413d21042b43b9d93e1309
26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371
4eac077a713c57b4f4397629a4145982c661f48044dd3f96427d40b147d9742f
de6a7b00deadc788eb6b6c8d20c0ae96c2f2019078fa604fee5b87d6e989ad7b
bcab477be20861e01e4a0e295284146a510150d9817763caf1a6f4b422d67042
2a292df7e32cababbd9de088d1d1abec9fc0440f637ed2fba145094dc14bea08
f4a9e534fc0d216c44b218fa0c42d99635a0127ee2e53c712f70609649fdff22
8268436f8c4126196cf64b3c7ddbda90746a378625f9813dd9b8457077256731
2810e5cbc2cc4d4eece54f61c6f69758e289aa7ab440b3cbeaa21995c2f4232b
3eb858e78f5a7254d8c9731174a94f76755fd3941c0ac93735c07ba14579630e
a45fdc55c76448c049a1ab33f17023edfb2be3581e9c7aade8a6125215e04220
d483fe813c6ba647ebbfd3ec41adca1c6130c2beeee9d9bf065c8d151c5f396e
8a2e1d30050198c65a54483123960ccc38aef6848e1ec8f5f780e8523769ba32
32888462f8b486c68ad7dd9610be5192bbeaf3b443951ac1a8118419d9fa097b
227142501b9d4355ccba290404bde41575b037693cef1f438c47f8fbf35d1165
5c37cc491da847cfeb9281d407efc41e15144c876e0170b499a96a22ed31e01e
445425117cb8c90edcbc7c1cc0e74f747f2c1efa5630a967c64f287792a48a4b
""".split()

# RFC 9496 Appendix A.3: element derivation, SHA-512 of the label -> from_uniform_bytes
HASH_TO_GROUP = [
    ("Ristretto is traditionally a short shot of espresso coffee", "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"),
    ("made with ten grams of coffee instead of the usual seven", None),
]


def test_generator_multiples():
    acc = G.IDENTITY
    for want in MULTIPLES:
        assert acc.encode().hex() == want
        dec = G.decode(bytes.fromhex(want))
        assert dec is not None and dec == acc and dec.encode().hex() == want
        acc = acc + G.BASEPOINT
    assert (G.L * G.BASEPOINT).is_identity()


def test_bad_encodings_rejected():
    for h in BAD:
        assert G.decode(bytes.fromhex(h)) is None, h


def test_hash_to_group_vector():
    label, want = HASH_TO_GROUP[0]
    assert G.hash_to_group_sha512(label.encode()).encode().hex() == want


def test_pedersen_blinding_generator():
    # curve25519-dalek's RISTRETTO_BASEPOINT hashed with SHA3-512: the B_blinding of upstream
    # dalek-bulletproofs PedersenGens::default
    got = G.from_uniform_bytes(hashlib.sha3_512(G.BASEPOINT.encode()).digest()).encode().hex()
    assert got == "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"


def test_edwards_layer_against_libsodium():
    nb = pytest.importorskip("nacl.bindings")
    import random

    r = random.Random(99)
    for _ in range(20):
        k = r.randrange(1, G.L)
        enc = nb.crypto_scalarmult_ed25519_base_noclamp(k.to_bytes(32, "little"))
        x, y = (k * G.BASEPOINT).affine()
        assert enc == (y | ((x & 1) << 255)).to_bytes(32, "little")
        a, b = r.randrange(G.L), r.randrange(G.L)
        assert int.from_bytes(nb.crypto_core_ed25519_scalar_mul(a.to_bytes(32, "little"), b.to_bytes(32, "little")), "little") == a * b % G.L


def test_msm_bucket_method_equals_definition():
    import random

    r = random.Random(5)
    for n in (0, 1, 7, 40):
        ks = [r.randrange(G.L) for _ in range(n)]
        ps = [r.randrange(G.L) * G.BASEPOINT for _ in range(n)]
        assert G.msm(ks, ps).encode() == G.msm_naive(ks, ps).encode()


def test_chacha20_block_rfc8439_vector():
    """RFC 8439 section 2.3.2: the block function behind the keyed blinding vectors."""
    from oracle import protocol as O

    key, nonce = bytes(range(32)), bytes.fromhex("000000090000004a00000000")
    want = (
        "10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
        "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e"
    )
    assert O.chacha20_block(key, 1, nonce).hex() == want
