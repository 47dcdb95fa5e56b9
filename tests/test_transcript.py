"""CPU: the host-side Fiat-Shamir transcript (RollingKeccakTranscript, prove/src/lib.rs:3211-3519).
Keccak-256 is pinned by the published known answers; the sponge/permutation additionally by hashlib's SHA3-256
(same permutation, padding byte 0x06) on multi-block inputs; the transcript layout by a direct restatement."""
import hashlib

from tokamak_b200 import transcript as TR


def _sha3_via_our_permutation(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x06)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = TR._keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def test_native_keccak_matches_the_restatement():
    """tkm_host_keccak256 (host-side function of the library, no device needed) against the pure-Python restatement on every
    length around the rate boundaries (135, 136, 137, 271, 272, ...) and on random data; keccak256 uses it when the library loads."""
    import random

    rng = random.Random(5)
    for n in list(range(0, 300)) + [1000, 1087, 1088, 1089, 4096]:
        data = bytes(rng.randrange(256) for _ in range(n))
        assert TR.keccak256(data) == TR.keccak256_py(data), n
    assert TR._native is not TR.keccak256_py, "the shared library should be loadable in this test environment"


def test_keccak256_known_answers():
    assert TR.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert TR.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    for n in (0, 1, 71, 72, 100, 135, 136, 137, 272, 1000):
        data = bytes((7 * i + 3) & 0xFF for i in range(n))
        assert _sha3_via_our_permutation(data) == hashlib.sha3_256(data).digest()
        assert TR.keccak256(data) != hashlib.sha3_256(data).digest()  # Keccak-256 is not SHA3-256


def test_rolling_transcript_layout():
    t = TR.RollingKeccakTranscript()
    v = bytes(range(1, 33))
    t.update(v)
    body = bytes(32) + bytes(32) + v
    assert t.state0 == TR.keccak256(b"\x00\x00\x00\x00" + body) and t.state1 == TR.keccak256(b"\x00\x00\x00\x01" + body)
    s0, s1 = t.state0, t.state1
    t.update(b"\xAB\xCD")  # short values are right-aligned in the 32-byte slot
    body = s0 + s1 + bytes(30) + b"\xAB\xCD"
    assert t.state0 == TR.keccak256(b"\x00\x00\x00\x00" + body) and t.state1 == TR.keccak256(b"\x00\x00\x00\x01" + body)
    raw = TR.keccak256(b"\x00\x00\x00\x02" + t.state0 + t.state1 + (0).to_bytes(4, "big"))
    c0 = t.get_challenge()
    assert c0 == int.from_bytes(bytes([raw[0] & 0x1F]) + raw[1:], "big") and 0 < c0 < TR.R_MOD
    raw1 = TR.keccak256(b"\x00\x00\x00\x02" + t.state0 + t.state1 + (1).to_bytes(4, "big"))
    assert t.get_challenge() == int.from_bytes(bytes([raw1[0] & 0x1F]) + raw1[1:], "big")


def test_manager_schedule_and_coordinate_split():
    import pyref as P

    pts = [P.g1_mul(P.G1_GEN, k) for k in range(2, 11)]
    m = TR.TranscriptManager()
    m.add_proof0(*pts[:6])
    thetas = m.get_thetas()
    assert len(thetas) == 3 and len(set(thetas)) == 3
    # the same absorbs done by hand: each coordinate = two updates, 16 high bytes zero-padded then 32 low bytes
    t = TR.RollingKeccakTranscript()
    for pt in pts[:6]:
        for coord in pt:
            be = coord.to_bytes(48, "big")
            t.update(bytes(16) + be[:16])
            t.update(be[16:])
    assert t.get_challenges(3) == thetas
    m.add_proof1(pts[6])
    k0 = m.get_kappa0()
    m.add_proof2(pts[7], pts[8])
    chi, zeta = m.get_chi_zeta()
    m.add_proof3(1, 2, 3, P.R_MOD - 1)
    k1 = m.get_kappa1()
    assert len({k0, chi, zeta, k1}) == 4 and m.transcript.challenge_counter == 7
    m2 = TR.TranscriptManager()  # identity commits as (0, 0)
    m2.transcript.commit_g1_point(None)
    m3 = TR.TranscriptManager()
    m3.transcript.commit_g1_point((0, 0))
    assert m2.transcript.state0 == m3.transcript.state0
