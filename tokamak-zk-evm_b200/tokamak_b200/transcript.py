"""Fiat-Shamir transcript of the Tokamak prover/verifier, host side (the reference computes it on the host with
tiny_keccak): RollingKeccakTranscript and the TranscriptManager absorb/squeeze schedule
(packages/backend/prove/src/lib.rs:3211-3730).  Hash = Keccak-256 with the original 0x01 padding (not SHA3-256).
A proof needs ~90 hashes of 100 bytes between its stages, while the device has nothing queued: keccak256 goes through the
library's host-side tkm_host_keccak256 (40 ms of pure Python per proof otherwise); keccak256_py is the same function in pure
Python (the KAT-pinned restatement, and what runs when the shared library is not built)."""

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256_py(data: bytes) -> bytes:
    """Keccak-256 (rate 136, original padding 0x01 .. 0x80), as tiny_keccak::Keccak::new_keccak256."""
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = _keccak_f(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


_native = None


def keccak256(data: bytes) -> bytes:
    """keccak256_py through the library's host function when the shared library is there (no device needed)."""
    global _native
    if _native is None:
        try:
            import ctypes

            from . import ffi

            lib = ffi.load()

            def _native(d, _lib=lib, _mk=ctypes.create_string_buffer):
                buf = _mk(32)  # per call: the transcript may be driven from more than one thread
                if _lib.tkm_host_keccak256(d, len(d), buf) != 0:
                    raise RuntimeError("tkm_host_keccak256 failed")
                return buf.raw

            if _native(b"") != keccak256_py(b""):  # paranoia: never trade a wrong transcript for speed
                raise RuntimeError("native Keccak disagrees with the restatement")
        except Exception:
            _native = keccak256_py
    return _native(bytes(data))


class RollingKeccakTranscript:
    """prove/src/lib.rs:3211-3519.  Two 32-byte states; update() hashes the 100-byte Solidity memory image
    [0,0,0,tag | state0 | state1 | value right-aligned in 32 bytes] with tag 0 -> state0 and tag 1 -> state1 (both from
    the OLD states); a challenge hashes [0,0,0,2 | state0 | state1 | counter_be32], masks the top 3 bits, maps 0 -> 1."""

    def __init__(self):
        self.state0 = bytes(32)
        self.state1 = bytes(32)
        self.challenge_counter = 0

    def update(self, value: bytes):
        if len(value) > 32:
            raise ValueError("Input must be 32 bytes or less")
        body = self.state0 + self.state1 + bytes(32 - len(value)) + value
        self.state0, self.state1 = keccak256(b"\x00\x00\x00\x00" + body), keccak256(b"\x00\x00\x00\x01" + body)

    def get_challenge_raw(self) -> bytes:
        buf = b"\x00\x00\x00\x02" + self.state0 + self.state1 + self.challenge_counter.to_bytes(4, "big")
        self.challenge_counter += 1
        return keccak256(buf)

    def get_challenge(self) -> int:
        raw = bytearray(self.get_challenge_raw())
        raw[0] &= 0x1F
        v = int.from_bytes(raw, "big")
        # ScalarField::from_bytes_le of a 253-bit value: already below r; "never zero"
        return v if v else 1

    def get_challenges(self, count):
        return [self.get_challenge() for _ in range(count)]

    def commit_field_as_bytes(self, fr: int):
        """32-byte big-endian scalar (commit_field_as_bytes, :3415-3426)."""
        self.update((fr % R_MOD).to_bytes(32, "big"))

    def commit_bls12_381_field_element(self, fq: int):
        """48-byte big-endian coordinate split 16 | 32, the 16-byte part zero-padded to 32 (:3429-3479)."""
        be = fq.to_bytes(48, "big")
        self.update(bytes(16) + be[:16])
        self.update(be[16:])

    def commit_g1_point(self, pt):
        """pt = (x, y) ints, identity = (0, 0) (commit_g1_point, :3482-3500)."""
        x, y = pt if pt is not None else (0, 0)
        self.commit_bls12_381_field_element(x)
        self.commit_bls12_381_field_element(y)


class TranscriptManager:
    """Absorb/squeeze schedule of the prover (prove/src/lib.rs:3521-3722): proof0 (U,V,W,Q_AX,Q_AY,B) -> thetas;
    proof1 (R) -> kappa0; proof2 (Q_CX,Q_CY) -> chi, zeta; proof3 (V_eval, R_eval, R_omegaX_eval, R_omegaX_omegaY_eval) -> kappa1."""

    def __init__(self):
        self.transcript = RollingKeccakTranscript()

    def add_proof0(self, U, V, W, Q_AX, Q_AY, B):
        for pt in (U, V, W, Q_AX, Q_AY, B):
            self.transcript.commit_g1_point(pt)

    def get_thetas(self):
        return self.transcript.get_challenges(3)

    def add_proof1(self, R):
        self.transcript.commit_g1_point(R)

    def get_kappa0(self):
        return self.transcript.get_challenge()

    def add_proof2(self, Q_CX, Q_CY):
        self.transcript.commit_g1_point(Q_CX)
        self.transcript.commit_g1_point(Q_CY)

    def get_chi_zeta(self):
        return self.transcript.get_challenges(2)

    def add_proof3(self, V_eval, R_eval, R_omegaX_eval, R_omegaX_omegaY_eval):
        for v in (V_eval, R_eval, R_omegaX_eval, R_omegaX_omegaY_eval):
            self.transcript.commit_field_as_bytes(v)

    def get_kappa1(self):
        return self.transcript.get_challenge()
