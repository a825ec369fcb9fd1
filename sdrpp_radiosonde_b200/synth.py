"""Synthetic radiosonde signal generators (test / bench input only — not on the product path).

Recipes follow SURVEY.md Appendix C: build the payload, apply the *inverse* of each
decoder's frame pipeline, NRZ-modulate at the raw baud rate, Gaussian-smooth, and
either emit the float "already FM-demodulated" stream the reference consumes
(SD/include/rs41.h:31-34) or frequency-modulate it to complex64 IQ for the
discriminator entry point.

Frame layouts cited from the reference (paths relative to /root/reference,
SD/ = src/decode/sondedump/):
  RS41    SD/sonde/rs41/protocol.h:56-61, frame.c:9-36,39-78
  DFM     SD/sonde/dfm09/protocol.h:33-52, frame.c:10-39,88-109
  M10     SD/sonde/m10/protocol.h:20-25, frame.c:7-54
  iMS-100 SD/sonde/ims100/protocol.h:120-153, frame.c:10-68
  MRZ-N1  SD/sonde/mrz-n1/protocol.h:20-48, frame.c:6-21
  iMet-4  SD/sonde/imet4/protocol.h:29-38, frame.c:5-24, subframe.c:8-30
  C50     SD/sonde/c50/protocol.h:33-42, frame.c:9-39
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# enum sonde_type (include/sonde_b200.h)
RS41, DFM09, M10, IMS100, MRZN1, IMET4, C50 = range(7)
TYPE_NAMES = ["rs41", "dfm09", "m10", "ims100", "mrzn1", "imet4", "c50"]


@dataclass(frozen=True)
class Modem:
    baud: int
    frame_bits: int
    syncword: int
    sync_len: int
    afsk: bool = False
    f_mark: float = 0.0
    f_space: float = 0.0


MODEMS = {
    RS41: Modem(4800, 4144, 0x086D53884469481F, 64),
    DFM09: Modem(2500, 560, 0x9A995A55, 32),
    M10: Modem(9600, 1664, 0x66666666B366, 48),
    IMS100: Modem(2400, 1200, 0xAAA56A659A99, 48),
    MRZN1: Modem(2400, 816, 0x666666666555A599, 64),
    IMET4: Modem(1200, 600, 0xFF40, 16, True, 2200.0, 1200.0),
    C50: Modem(2380, 90, 0x005FF, 20, True, 4700.0, 2900.0),
}

_GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# ----------------------------------------------------------------------------------------
# bit helpers
# ----------------------------------------------------------------------------------------

def bytes_to_bits(b) -> np.ndarray:
    """MSB-first bit expansion."""
    return np.unpackbits(np.asarray(bytearray(b), dtype=np.uint8))


def bits_to_bytes(bits) -> bytes:
    return np.packbits(np.asarray(bits, dtype=np.uint8)).tobytes()


def int_to_bits(v: int, n: int) -> np.ndarray:
    return np.array([(v >> (n - 1 - i)) & 1 for i in range(n)], dtype=np.uint8)


def manchester_encode(bits: np.ndarray) -> np.ndarray:
    """data bit b -> raw pair (not b, b); the decoder keeps the 2nd bit (SD/decode/manchester.c:17)."""
    bits = np.asarray(bits, dtype=np.uint8)
    out = np.empty(bits.size * 2, dtype=np.uint8)
    out[0::2] = bits ^ 1
    out[1::2] = bits
    return out


def bitrev8(x: int) -> int:
    return int(f"{x:08b}"[::-1], 2)


# ----------------------------------------------------------------------------------------
# GF(2^m) helpers (RS41 parity, iMS-100 BCH parity)
# ----------------------------------------------------------------------------------------

class GF:
    def __init__(self, n: int, poly: int):
        self.n = n
        self.exp = [0] * (2 * n + 2)
        self.log = [0] * (n + 1)
        x = 1
        for i in range(n):
            self.exp[i] = x
            self.log[x] = i
            x <<= 1
            if x > n:
                x ^= poly
        for i in range(n, 2 * n + 2):
            self.exp[i] = self.exp[i - n]

    def mul(self, a, b):
        if a == 0 or b == 0:
            return 0
        return self.exp[self.log[a] + self.log[b]]

    def inv(self, a):
        return self.exp[(self.n - self.log[a]) % self.n]

    def pow_alpha(self, e):
        return self.exp[e % self.n]

    def solve(self, A, b):
        """Gaussian elimination over the field: A x = b."""
        m = len(A)
        M = [row[:] + [b[i]] for i, row in enumerate(A)]
        for col in range(m):
            piv = next(r for r in range(col, m) if M[r][col])
            M[col], M[piv] = M[piv], M[col]
            iv = self.inv(M[col][col])
            M[col] = [self.mul(v, iv) for v in M[col]]
            for r in range(m):
                if r != col and M[r][col]:
                    f = M[r][col]
                    M[r] = [v ^ self.mul(f, w) for v, w in zip(M[r], M[col])]
        return [M[i][m] for i in range(m)]


_gf256 = None
_gf64 = None


def gf256() -> GF:
    global _gf256
    if _gf256 is None:
        _gf256 = GF(255, 0x11D)
    return _gf256


def gf64() -> GF:
    global _gf64
    if _gf64 is None:
        _gf64 = GF(63, 0x61)
    return _gf64


def rs255_parity(msg231) -> list:
    """Parity c[231..254] so that sum c[i]*alpha^(i*j) = 0 for j = 0..23 (SD/decode/ecc/rs.c:120-128)."""
    g = gf256()
    synd = []
    for j in range(24):
        s = 0
        for i, v in enumerate(msg231):
            if v:
                s ^= g.mul(v, g.pow_alpha(i * j))
        synd.append(s)
    A = [[g.pow_alpha((231 + k) * j) for k in range(24)] for j in range(24)]
    return g.solve(A, synd)


_bch_cache = None


def bch_parity(data34) -> list:
    """12 parity bits at symbol indices 51..62 for data bits at 17..50, roots alpha^1..alpha^4 of GF(64)/0x61."""
    global _bch_cache
    g = gf64()
    if _bch_cache is None:
        # parity is linear in the data bits: precompute the parity of each unit vector
        rows = []
        # unknown parity bits p_k in GF(2); solve over GF(2) using the bit expansion of the 4 syndromes
        roots = [2, 4, 8, 16]
        logs = [g.log[r] for r in roots]

        def synd_bits(pos):
            out = []
            for lr in logs:
                v = g.pow_alpha(lr * pos)
                out.extend([(v >> b) & 1 for b in range(6)])
            return out

        P = np.array([synd_bits(51 + k) for k in range(12)], dtype=np.uint8).T  # 24 x 12
        D = np.array([synd_bits(17 + k) for k in range(34)], dtype=np.uint8).T  # 24 x 34
        # solve P p = D d over GF(2) for each unit d: least-squares style elimination
        aug = np.concatenate([P, D], axis=1) % 2
        r = 0
        piv_cols = []
        for c in range(12):
            pr = next((i for i in range(r, aug.shape[0]) if aug[i, c]), None)
            if pr is None:
                continue
            aug[[r, pr]] = aug[[pr, r]]
            for i in range(aug.shape[0]):
                if i != r and aug[i, c]:
                    aug[i] ^= aug[r]
            piv_cols.append(c)
            r += 1
        assert piv_cols == list(range(12)), "BCH parity system is rank deficient"
        _bch_cache = aug[:12, 12:].copy()  # 12 x 34: p = M d
    M = _bch_cache
    d = np.asarray(data34, dtype=np.uint8)
    return list((M @ d) % 2)


# ----------------------------------------------------------------------------------------
# CRCs (restated for the generator only)
# ----------------------------------------------------------------------------------------

def crc16_msb(data: bytes, init: int, poly: int = 0x1021) -> int:
    crc = init
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ poly) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def crc16_modbus(data: bytes) -> int:
    crc = 0xFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0xA001 if crc & 1 else crc >> 1
    return crc


def m10_checksum(data: bytes) -> int:
    c = 0
    for b in data:
        c1 = c & 0xFF
        b = ((b >> 1) | ((b & 1) << 7)) & 0xFF
        b ^= (b >> 2) & 0xFF
        t6 = (c & 1) ^ ((c >> 2) & 1) ^ ((c >> 4) & 1)
        t7 = ((c >> 1) & 1) ^ ((c >> 3) & 1) ^ ((c >> 5) & 1)
        t = (c & 0x3F) | (t6 << 6) | (t7 << 7)
        s = (c >> 7) & 0xFF
        s ^= (s >> 2) & 0xFF
        c0 = b ^ t ^ s
        c = ((c1 << 8) | c0) & 0xFFFF
    return c


# ----------------------------------------------------------------------------------------
# frame builders: return the RAW on-air bit sequence (np.uint8 0/1) of one frame
# ----------------------------------------------------------------------------------------

RS41_PRN = bytes([
    0x96, 0x83, 0x3e, 0x51, 0xb1, 0x49, 0x08, 0x98, 0x32, 0x05, 0x59, 0x0e, 0xf9, 0x44, 0xc6, 0x26,
    0x21, 0x60, 0xc2, 0xea, 0x79, 0x5d, 0x6d, 0xa1, 0x54, 0x69, 0x47, 0x0c, 0xdc, 0xe8, 0x5c, 0xf1,
    0xf7, 0x76, 0x82, 0x7f, 0x07, 0x99, 0xa2, 0x2c, 0x93, 0x7c, 0x30, 0x63, 0xf5, 0x10, 0x2e, 0x61,
    0xd0, 0xbc, 0xb4, 0xb6, 0x06, 0xaa, 0xf4, 0x23, 0x78, 0x6e, 0x3b, 0xae, 0xbf, 0x7b, 0x4c, 0xc1,
])
RS41_HEADER = bytes([0x86, 0x35, 0xF4, 0x40, 0x93, 0xDF, 0x1A, 0x60])

_golden_rs41 = None


def golden_rs41_frame() -> bytes:
    """518-byte descrambled RS41 frame: header + the 510 bytes of SD/scripts/rs_bruteforce.py:6
    (committed as tests/golden/rs41_frame.hex by tests/golden/make_golden.py)."""
    global _golden_rs41
    if _golden_rs41 is None:
        with open(os.path.join(_GOLDEN_DIR, "rs41_frame.hex")) as f:
            body = bytes.fromhex(f.read().strip())
        assert len(body) == 510
        _golden_rs41 = RS41_HEADER + body
    return _golden_rs41


def rs41_reencode(frame: bytearray) -> None:
    """Recompute both interleaved RS(255,231) parities in place (SD/sonde/rs41/frame.c:52-75)."""
    extended = frame[56] == 0xF0
    chunk = 231 if extended else 132
    for block in range(2):
        msg = [0] * 231
        for i in range(chunk):
            msg[i] = frame[57 + 2 * i + block - 1]
        par = rs255_parity(msg)
        frame[8 + 24 * block: 8 + 24 * (block + 1)] = bytes(par)


def rs41_frame_bytes(seq: int | None = None, serial: str | None = None, fix_golden: bool = True) -> bytes:
    """A valid 518-byte descrambled RS41 frame derived from the golden one."""
    fr = bytearray(golden_rs41_frame())
    if fix_golden:
        fr[8 + 509] = 0x14          # the golden frame carries one byte error
    if seq is not None or serial is not None:
        # first subframe: type 0x79 (status), len 0x28 at data[0..1]; seq at +2, serial at +4
        base = 57
        assert fr[base] == 0x79
        ln = fr[base + 1]
        if seq is not None:
            fr[base + 2] = seq & 0xFF
            fr[base + 3] = (seq >> 8) & 0xFF
        if serial is not None:
            s = serial.encode()[:8].ljust(8, b"0")
            fr[base + 4: base + 12] = s
        crc = crc16_msb(bytes(fr[base + 2: base + 2 + ln]), 0xFFFF)
        fr[base + 2 + ln] = crc & 0xFF
        fr[base + 3 + ln] = crc >> 8
        rs41_reencode(fr)
    return bytes(fr)


def rs41_raw_bits(frame518: bytes) -> np.ndarray:
    """on-air byte = F[i] ^ prn[i%64], LSB first (inverse of SD/sonde/rs41/frame.c:21-36)."""
    air = bytes(bitrev8(b ^ RS41_PRN[i % 64]) for i, b in enumerate(frame518))
    return bytes_to_bits(air)


def _hamming84(nib: int) -> int:
    """DFM codeword: data nibble in the high 4 bits + low nibble making the four masked parities even."""
    for low in range(16):
        cw = (nib << 4) | low
        if all(bin(cw & m).count("1") % 2 == 0 for m in (0xAA, 0x66, 0x1E, 0xFF)):
            return cw
    raise AssertionError


def _dfm_interleave(codewords, depth) -> np.ndarray:
    """air bit p*D + k = codeword k bit p (MSB first) (inverse of SD/sonde/dfm09/frame.c:18-26)."""
    out = np.zeros(8 * depth, dtype=np.uint8)
    for k, cw in enumerate(codewords):
        for p in range(8):
            out[p * depth + k] = (cw >> (7 - p)) & 1
    return out


def dfm_raw_bits(ptu_type: int, ptu_data: bytes, gps: list) -> np.ndarray:
    """gps = [(type, 6 data bytes), (type, 6 data bytes)]."""
    nibs = [ptu_type & 0xF]
    for b in ptu_data[:3]:
        nibs += [b >> 4, b & 0xF]
    bits = [int_to_bits(0x45CF, 16), _dfm_interleave([_hamming84(n) for n in nibs], 7)]
    for typ, data in gps:
        nibs = []
        for b in data[:6]:
            nibs += [b >> 4, b & 0xF]
        nibs.append(typ & 0xF)
        bits.append(_dfm_interleave([_hamming84(n) for n in nibs], 13))
    payload = np.concatenate(bits)
    assert payload.size == 280
    return manchester_encode(payload)


def m10_raw_bits(payload97: bytes, ftype: int = 0x9F) -> np.ndarray:
    D = bytearray([0x64, ftype]) + bytearray(payload97[:97].ljust(97, b"\0"))
    chk = m10_checksum(bytes(D))
    D += bytes([chk >> 8, chk & 0xFF])
    assert len(D) == 101
    dbits = bytes_to_bits(bytes(D))
    # pre-descramble stream E: first 24 bits = sync (AA AA 5A pattern in the descrambled domain is
    # irrelevant: they are overwritten by the literal sync word after Manchester coding)
    E = np.zeros(24 + dbits.size, dtype=np.uint8)
    E[:24] = bytes_to_bits(bytes([0xAA, 0xAA, 0x5A]))
    for k in range(dbits.size):
        E[24 + k] = E[24 + k - 1] ^ dbits[k] ^ 1
    raw = manchester_encode(E)
    assert raw.size == 1664
    raw[:48] = int_to_bits(0x66666666B366, 48)
    return raw


def ims100_raw_bits(payload48: bytes) -> np.ndarray:
    vals = [(payload48[2 * i] << 8) | payload48[2 * i + 1] for i in range(24)]
    out = []
    for sf in range(2):
        out.append(np.zeros(24, dtype=np.uint8))          # filler where the sync word sits
        for m in range(6):
            pair = []
            for v in vals[sf * 12 + m * 2: sf * 12 + m * 2 + 2]:
                b = int_to_bits(v, 16)
                par = (int(b.sum()) & 1) ^ 1              # odd parity: validity requires ones&1 != parity bit
                pair.append(np.concatenate([b, [par]]))
            d34 = np.concatenate(pair).astype(np.uint8)
            out.append(np.concatenate([d34, np.array(bch_parity(d34), dtype=np.uint8)]))
    o = np.concatenate(out)
    assert o.size == 600
    # differential pre-code: in[k] = out[k] ^ in[k+1], in[600] = 0 (SD/sonde/ims100/frame.c:10-19)
    inn = np.zeros(601, dtype=np.uint8)
    for k in range(599, -1, -1):
        inn[k] = o[k] ^ inn[k + 1]
    raw = manchester_encode(inn[:600])
    raw[:48] = int_to_bits(0xAAA56A659A99, 48)
    return raw


def mrzn1_raw_bits(body45: bytes) -> np.ndarray:
    body = bytes(body45[:45].ljust(45, b"\0"))
    crc = crc16_modbus(body)
    fr = bytes([0xAA, 0xAA, 0xBF, 0x35]) + body + bytes([crc & 0xFF, crc >> 8])
    assert len(fr) == 51
    raw = manchester_encode(bytes_to_bits(fr))
    raw[:64] = int_to_bits(0x666666666555A599, 64)
    return raw


def uart_8n1(data: bytes) -> np.ndarray:
    out = np.empty(10 * len(data), dtype=np.uint8)
    for i, b in enumerate(data):
        out[10 * i] = 0
        for j in range(8):
            out[10 * i + 1 + j] = (b >> j) & 1
        out[10 * i + 9] = 1
    return out


def imet4_subframe(sftype: int, body: bytes) -> bytes:
    sf = bytes([0x01, sftype]) + body
    crc = crc16_msb(sf, 0x1D0F)
    return sf + bytes([crc >> 8, crc & 0xFF])


def imet4_raw_bits(rng: np.random.Generator, idle_bits: int = 60) -> np.ndarray:
    """Idle marks, then a GPS (type 2, 18 B) and a PTU (type 1, 14 B) subframe burst."""
    gps = imet4_subframe(0x02, bytes(rng.integers(0, 256, 14, dtype=np.uint8)))   # sizeof(GPS)=16 -> +2 crc
    ptu = imet4_subframe(0x01, bytes(rng.integers(0, 256, 10, dtype=np.uint8)))   # sizeof(PTU)=12 -> +2 crc
    return np.concatenate([np.ones(idle_bits, dtype=np.uint8), uart_8n1(ptu), np.ones(8, dtype=np.uint8),
                           uart_8n1(gps)])


def c50_raw_bits(ftype: int, data4: bytes) -> np.ndarray:
    s0 = s1 = 0
    for b in bytes([ftype]) + data4[:4]:
        s0 = (s0 + b) & 0xFF
        s1 = (s1 + s0) & 0xFF
    return uart_8n1(bytes([0x00, 0xFF, ftype]) + data4[:4] + bytes([s0, s1 ^ 0xFF]))


def random_frame_bits(stype: int, rng: np.random.Generator, index: int = 0) -> np.ndarray:
    """One valid frame of raw on-air bits for `stype` with pseudo-random payload."""
    if stype == RS41:
        return rs41_raw_bits(rs41_frame_bytes())
    if stype == DFM09:
        g = [(int(rng.integers(0, 9)), bytes(rng.integers(0, 256, 6, dtype=np.uint8))) for _ in range(2)]
        return dfm_raw_bits(int(rng.integers(0, 16)), bytes(rng.integers(0, 256, 3, dtype=np.uint8)), g)
    if stype == M10:
        return m10_raw_bits(bytes(rng.integers(0, 256, 97, dtype=np.uint8)), 0x9F if index % 2 == 0 else 0x20)
    if stype == IMS100:
        p = bytearray(rng.integers(0, 256, 48, dtype=np.uint8))
        p[14] = 0x30
        p[15] = 0xC1 if index % 2 == 0 else 0xA2
        return ims100_raw_bits(bytes(p))
    if stype == MRZN1:
        body = bytearray(rng.integers(0, 256, 45, dtype=np.uint8))
        # calib_frag_seq (frame byte 44) is 1..16 on air; the reference indexes its 64-byte calibration
        # image with it unchecked (mrzn1.c:136-141), so anything else corrupts the reference's own state
        body[40] = 1 + index % 16
        return mrzn1_raw_bits(bytes(body))
    if stype == IMET4:
        return imet4_raw_bits(rng)
    if stype == C50:
        return c50_raw_bits(int(rng.choice([0x03, 0x10, 0x14, 0x15, 0x16, 0x17, 0x18, 0x64])),
                            bytes(rng.integers(0, 256, 4, dtype=np.uint8)))
    raise ValueError(stype)


# ----------------------------------------------------------------------------------------
# telemetry "flights": consecutive frames with running counters and a complete calibration,
# so that the host parsers' stateful paths (calibration assembly, serial shards, climb rate) run
# ----------------------------------------------------------------------------------------

def _mbf_le(x: float) -> bytes:
    """float -> Microsoft binary format, little endian (inverse of SD/bitops.c:132-152)."""
    u = int(np.float32(x).view(np.uint32))
    sign, exp, man = u >> 31, (u >> 23) & 0xFF, u & 0x7FFFFF
    return bytes([man & 0xFF, (man >> 8) & 0xFF, (sign << 7) | (man >> 16), (exp + 2) & 0xFF])


def meisei_calibration(rs11g: bool) -> np.ndarray:
    """64 plausible calibration floats laid out like IMS100Calibration / RS11GCalibration (ims100/protocol.h:155-195)."""
    c = np.zeros(64, dtype=np.float32)
    c[0] = 4120537.0 if not rs11g else 2310644.0                  # serial
    if not rs11g:
        c[17:29] = [60, 50, 40, 30, 20, 10, 0, -20, -40, -60, -75, -85]                  # temps
        c[33:45] = [5.6, 7.9, 11.4, 16.9, 25.7, 40.2, 64.9, 187.0, 640.0, 2730.0, 9600.0, 24500.0]   # kOhm
        c[49:53] = [-120.0, 310.0, -170.0, 35.0]                  # rh_poly
        c[53:57] = [-3.2, 41.0, 2.5, 0.7]                         # temp_poly
        c[57:60] = [2.1e-7, 2.9e-4, 1.9e-3]                       # rh_temp_poly
    else:
        c[17:28] = [40, 30, 20, 10, 0, -20, -40, -55, -65, -75, -85]
        c[33:37] = [-2.9, 40.0, 2.0, 0.1]                         # temp_poly
        c[37:48] = [11.4, 16.9, 25.7, 40.2, 64.9, 187.0, 640.0, 1900.0, 4300.0, 9600.0, 24500.0]
        c[49:53] = [-110.0, 300.0, -160.0, 30.0]
    return c


def meisei_flight_bits(rs11g: bool, n_frames: int, rng: np.random.Generator, seq0: int = 0) -> np.ndarray:
    cal = meisei_calibration(rs11g)
    out = []
    for k in range(n_frames):
        seq = seq0 + k
        p = bytearray(rng.integers(0, 256, 48, dtype=np.uint8))
        p[0:2] = seq.to_bytes(2, "big")
        if rs11g:
            p[4:8] = _mbf_le(float(cal[seq % 64]))
        else:
            b = np.float32(cal[seq % 64]).view(np.uint32).item().to_bytes(4, "big")
            p[4:8] = bytes([b[2], b[3], b[0], b[1]])
        ref = 31000 + int(rng.integers(0, 50))
        p[10:12] = int(rng.integers(9000, 22000)).to_bytes(2, "big")                      # temperature count
        if rs11g or seq % 4 != 3:
            p[2:4] = ref.to_bytes(2, "big")
            p[12:14] = int(rng.integers(9000, 14000)).to_bytes(2, "big")                  # humidity count
        else:
            p[2:4] = int(rng.integers(9000, 22000)).to_bytes(2, "big")                    # humidity-sensor temperature
            p[12:14] = ref.to_bytes(2, "big")
        p[14] = 0x30 if k % 2 == 0 else 0x31
        p[15] = 0xA2 if rs11g else 0xC1
        if not rs11g and p[14] == 0x30:
            p[20:22] = int((k % 60) * 1000).to_bytes(2, "big")
            p[22], p[23] = 11, (k // 60) % 60
            p[24:26] = (17 * 1000 + 10 * 10 + 6).to_bytes(2, "big")                       # 17 Oct, year ...6
            p[26:30] = (35123456 + 20 * k).to_bytes(4, "big")
            p[30:34] = (139456789 + 31 * k).to_bytes(4, "big")
            p[34:37] = (120000 + 550 * k).to_bytes(3, "big")
        elif rs11g and p[14] == 0x30:
            p[26:30] = (351234560 + 200 * k).to_bytes(4, "big")
            p[30:34] = (1394567890 + 310 * k).to_bytes(4, "big")
            p[34:38] = (120000 + 550 * k).to_bytes(4, "big")
            p[45], p[46], p[47] = (2026 - 0x700) & 0xFF, 10, 17
        elif rs11g:
            p[20:22] = int((k % 60) * 1000).to_bytes(2, "little")
            p[22], p[23] = 11, (k // 60) % 60
        out.append(ims100_raw_bits(bytes(p)))
    return np.concatenate(out)


def dfm_flight_bits(n_frames: int, rng: np.random.Generator, dfm06: bool = False) -> np.ndarray:
    """PTU channels cycle 0..7 (channel 6 all zero, channel 7 = serial shards), GPS slots cycle through 0,1 / 2,3 / 4,8."""
    out = []
    serial = 0x0017_4C2B
    zero_ch = 5 if dfm06 else 6
    for k in range(n_frames):
        t = k % 8
        if t == 0:
            ptu = bytes([0x25, 0x13, 0x40 + k % 16])                 # thermistor reading
        elif t in (3, 4):
            ptu = bytes([0x27 if t == 3 else 0x2B, 0x10 + t, 0x55])  # the two reference channels
        elif t == zero_ch:
            ptu = bytes([0x40, 0, 0])                                # low 16 bits zero: the serial follows
        elif t == zero_ch + 1:
            if dfm06:
                ptu = bytes([0x00, 0x61, 0x23])
            else:
                shard_idx = (k // 8) % 2
                shard = (serial >> (16 * (1 - shard_idx))) & 0xFFFF
                ptu = ((shard << 4) | shard_idx).to_bytes(3, "big")
        else:
            ptu = bytes(rng.integers(1, 256, 3, dtype=np.uint8))
        ptu_type = t
        pairs = [(0, 1), (2, 3), (4, 8)][k % 3]
        gps = []
        for g in pairs:
            if g == 0:
                d = bytes([0, 0, 0, k & 0xFF, 0, 0])
            elif g == 1:
                d = bytes([0, 0, 0, 0]) + int((k % 60) * 1000).to_bytes(2, "big")
            elif g == 2:
                d = (481234567 + 35 * k).to_bytes(4, "big") + (1234 + k).to_bytes(2, "big")
            elif g == 3:
                d = (116543210 + 51 * k).to_bytes(4, "big") + (9000 + 7 * k).to_bytes(2, "big")
            elif g == 4:
                d = (1500000 + 480 * k).to_bytes(4, "big") + ((520 - 3 * k) & 0xFFFF).to_bytes(2, "big")
            else:
                raw = (2026 << 20) | (10 << 16) | (17 << 11) | (11 << 6) | (k // 60 % 60)
                d = raw.to_bytes(4, "big") + bytes([0, 0])
            gps.append((g, d))
        out.append(dfm_raw_bits(ptu_type, ptu, gps))
    return np.concatenate(out)


# ----------------------------------------------------------------------------------------
# modulation
# ----------------------------------------------------------------------------------------

def _gauss_taps(sps: float, bt: float = 0.5) -> np.ndarray:
    sigma = np.sqrt(np.log(2.0)) / (2.0 * np.pi * bt) * sps
    half = int(np.ceil(4 * sigma))
    t = np.arange(-half, half + 1, dtype=np.float64)
    h = np.exp(-0.5 * (t / sigma) ** 2)
    return h / h.sum()


def nrz_baseband(bits: np.ndarray, n_samples: int, baud: float, fs: float, timing_offset: float = 0.0,
                 ppm: float = 0.0, bt: float = 0.5) -> np.ndarray:
    """Gaussian-filtered +/-1 NRZ of `bits` (repeated cyclically), float64, n_samples long."""
    rate = baud * (1.0 + ppm * 1e-6) / fs
    idx = np.floor((np.arange(n_samples, dtype=np.float64)) * rate + timing_offset).astype(np.int64)
    nrz = bits[idx % bits.size].astype(np.float64) * 2.0 - 1.0
    return np.convolve(nrz, _gauss_taps(fs / baud, bt), mode="same")


def afsk_audio(bits: np.ndarray, n_samples: int, modem: Modem, fs: float, timing_offset: float = 0.0,
               ppm: float = 0.0) -> np.ndarray:
    rate = modem.baud * (1.0 + ppm * 1e-6) / fs
    idx = np.floor(np.arange(n_samples, dtype=np.float64) * rate + timing_offset).astype(np.int64)
    b = bits[idx % bits.size]
    f = np.where(b == 1, modem.f_mark, modem.f_space)
    ph = 2.0 * np.pi * np.cumsum(f) / fs
    return np.sin(ph)


@dataclass
class ChannelSpec:
    stype: int
    seed: int
    snr_db: float = 20.0
    cfo_hz: float = 0.0
    timing_offset: float = 0.0
    ppm: float = 0.0
    n_distinct_frames: int = 4
    bit_errors: int = 0          # random raw-bit flips per frame (FEC exercise)
    gap_bits: int = 0            # idle bits between frames
    custom_bits: np.ndarray | None = None   # on-air bits to send instead of the random frames (telemetry "flights")


def channel_bits(spec: ChannelSpec) -> np.ndarray:
    if spec.custom_bits is not None:
        return spec.custom_bits
    rng = np.random.default_rng(spec.seed)
    frames = []
    for k in range(spec.n_distinct_frames):
        fb = random_frame_bits(spec.stype, rng, k).copy()
        if spec.bit_errors:
            pos = rng.choice(fb.size - MODEMS[spec.stype].sync_len, spec.bit_errors, replace=False)
            fb[pos + MODEMS[spec.stype].sync_len] ^= 1
        frames.append(fb)
        if spec.gap_bits:
            frames.append(rng.integers(0, 2, spec.gap_bits, dtype=np.uint8))
    return np.concatenate(frames)


def make_fm(spec: ChannelSpec, n_samples: int, fs: float = 48000.0, amp: float = 0.3) -> np.ndarray:
    """Float32 'already FM-demodulated' stream (what xxx_decode() consumes)."""
    m = MODEMS[spec.stype]
    rng = np.random.default_rng(spec.seed ^ 0x5EED)
    bits = channel_bits(spec)
    if m.afsk:
        x = amp * afsk_audio(bits, n_samples, m, fs, spec.timing_offset, spec.ppm)
    else:
        x = amp * nrz_baseband(bits, n_samples, m.baud, fs, spec.timing_offset, spec.ppm)
        x = x + spec.cfo_hz / 2400.0 * amp           # a CFO shows up as a DC offset after the discriminator
    sigma = amp * 10.0 ** (-spec.snr_db / 20.0)
    x = x + rng.normal(0.0, sigma, n_samples)
    return x.astype(np.float32)


def make_iq(spec: ChannelSpec, n_samples: int, fs: float = 48000.0, f_dev: float = 2400.0) -> np.ndarray:
    """complex64 IQ: x[n] = exp(j*2*pi*cumsum(f_dev*m[n] + cfo)/fs) + AWGN."""
    m = MODEMS[spec.stype]
    rng = np.random.default_rng(spec.seed ^ 0x1C0FFEE)
    bits = channel_bits(spec)
    if m.afsk:
        msg = afsk_audio(bits, n_samples, m, fs, spec.timing_offset, spec.ppm)
        dev = 3000.0
    else:
        msg = nrz_baseband(bits, n_samples, m.baud, fs, spec.timing_offset, spec.ppm)
        dev = f_dev
    ph = 2.0 * np.pi * np.cumsum(dev * msg + spec.cfo_hz) / fs
    sigma = 10.0 ** (-spec.snr_db / 20.0) / np.sqrt(2.0)
    iq = np.exp(1j * ph) + rng.normal(0.0, sigma, n_samples) + 1j * rng.normal(0.0, sigma, n_samples)
    return iq.astype(np.complex64)


def default_spec(stype: int, channel: int, base_seed: int = 0xB200, impaired: bool = True) -> ChannelSpec:
    """Per-channel impairments of SURVEY.md §8d: seed 0xB200+c, CFO U(-500,500) Hz, timing U(0,1) symbol,
    clock U(-50,50) ppm, SNR U(12,25) dB."""
    rng = np.random.default_rng(base_seed + channel)
    if not impaired:
        return ChannelSpec(stype, base_seed + channel)
    return ChannelSpec(stype, base_seed + channel,
                       snr_db=float(rng.uniform(12, 25)), cfo_hz=float(rng.uniform(-500, 500)),
                       timing_offset=float(rng.uniform(0, 1)), ppm=float(rng.uniform(-50, 50)))


def make_batch(types, n_samples: int, kind: str = "iq", base_seed: int = 0xB200, impaired: bool = True,
               **kw) -> np.ndarray:
    """[C][n_samples] batch, channel c built from default_spec(types[c], c)."""
    rows = []
    for c, t in enumerate(types):
        spec = default_spec(int(t), c, base_seed, impaired)
        for k, v in kw.items():
            setattr(spec, k, v)
        rows.append(make_iq(spec, n_samples) if kind == "iq" else make_fm(spec, n_samples))
    return np.stack(rows)
