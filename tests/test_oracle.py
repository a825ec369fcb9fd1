"""CPU: pins the C restatement (oracle/sonde_oracle.c) against
  (1) the committed fixtures the UNMODIFIED reference produced (tests/golden/, make_golden.py),
  (2) the known-answer vectors harvested from the reference's own scripts
      (SD/scripts/rs_bruteforce.py:6 RS41 frame, :8-26 BCH messages),
  (3) the compiled reference itself (oracle/_ref) when it is present in this checkout.
"""
import json
import os
import zlib

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import synth
from tests import reflib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def orc():
    if not reflib.have_oracle():
        pytest.fail("oracle/_build/libsonde_oracle.so missing: run __graft_entry__.build()")
    return reflib.OracleLib()


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_rs41_golden_frame_known_answer(orc):
    ka = load("fec_known_answers.json")
    frame = bytearray(synth.golden_rs41_frame())
    assert orc.rs41_correct(frame) == ka["rs41_first_pass"] == 1
    assert bytes(frame).hex() == ka["rs41_corrected_hex"]
    assert [i for i in range(518) if frame[i] != synth.golden_rs41_frame()[i]] == ka["rs41_changed_bytes"] == [517]
    assert frame[517] == 0x14
    assert orc.rs41_correct(frame) == ka["rs41_second_pass"] == 0
    # first subframe CRC16-CCITT-FALSE: type 0x79 len 0x28
    sub = bytes(frame[57:])
    assert sub[0] == 0x79
    ln = sub[1]
    assert orc.checksum("crc16_ccitt_false", sub[2:2 + ln]) == sub[2 + ln] | sub[3 + ln] << 8 == 0x1B72


def test_bch_known_answers(orc):
    ka = load("fec_known_answers.json")
    assert len(ka["bch"]) == 24
    for case in ka["bch"]:
        m = bytearray(bytes.fromhex(case["before_hex"]))
        assert orc.bch_fix(m) == case["ret"]
        assert bytes(m).hex() == case["after_hex"]
        if len(case["flip"]) < 3:               # 3 flips exceed t=2: -1 or a mis-correction, as the reference says
            assert case["ret"] == len(case["flip"])


@pytest.mark.parametrize("name", synth.TYPE_NAMES)
def test_frames_match_reference_fixtures(orc, name):
    gold = load(f"ref_{name}.json")
    stype = gold["type"]
    for case in gold["cases"]:
        spec = synth.default_spec(stype, case["channel"])
        if stype == synth.RS41:
            spec.bit_errors = 20
        fm = synth.make_fm(spec, case["n"])
        assert int(np.uint32(zlib.crc32(fm.tobytes()))) == case["fm_crc32"], "signal generator drifted"
        recs = orc.frames_run(stype, fm, case["chunk"])
        assert len(recs) == len(case["frames"])
        rb = (synth.MODEMS[stype].frame_bits + 7) // 8
        for r, g in zip(recs, case["frames"]):
            assert (r.chunk, r.sync_offset, r.inverted, r.status, r.ok, r.aux) == \
                   (g["chunk"], g["sync_offset"], g["inverted"], g["status"], g["ok"], g["aux"])
            assert bytes(r.data[:max(r.data_len, 132)]).hex() == g["data_hex"]
            assert bytes(r.raw[:rb]).hex() == g["raw_hex"]
        assert sum(g["ok"] for g in case["frames"]) > 0


def test_modem_tables_match_reference(orc):
    tabs = load("modem_tables.json")
    for name, t in tabs.items():
        assert [int(v) for v in orc.gfsk_taps(t["baud"]).view(np.uint32)] == t["taps_u32"], name
        assert [int(v) for v in orc.gfsk_timing(t["baud"]).view(np.uint32)] == t["timing_u32"], name


def test_soft_symbols_bit_exact_vs_fixture(orc):
    gold = load("soft_symbols.json")
    for name, g in gold.items():
        stype = synth.TYPE_NAMES.index(name)
        fm = synth.make_fm(synth.default_spec(stype, g["channel"]), g["n"])
        assert int(np.uint32(zlib.crc32(fm.tobytes()))) == g["fm_crc32"]
        soft, state = orc.gfsk_soft(synth.MODEMS[stype].baud, fm, g["chunk"])
        assert [int(v) for v in soft.view(np.uint32)] == g["soft_u32"]
        assert [int(v) for v in state[:6].view(np.uint32)] == g["state_u32"]
        bits = orc.demod_bits(stype, fm, g["chunk"])
        assert bits.size == g["nbits"] and np.packbits(bits).tobytes().hex() == g["bits_hex"]


def test_rs_error_injection_round_trip(orc):
    """encode -> corrupt -> decode: up to 12 byte errors per interleaved block are corrected, more are not."""
    rng = np.random.default_rng(7)
    clean = bytearray(synth.rs41_frame_bytes(seq=1234, serial="T1234567"))
    assert orc.rs41_correct(bytearray(clean)) == 0
    for nerr in (1, 5, 12, 20, 24):
        fr = bytearray(clean)
        # spread over both blocks; avoid the symbol-0 positions (index quirk, tested separately)
        pos = rng.choice(np.arange(58, 518), nerr, replace=False)
        for p in pos:
            fr[p] ^= int(rng.integers(1, 256))
        per_block = [sum(1 for p in pos if (p - 57 + 1) % 2 == b) for b in (0, 1)]
        ret = orc.rs41_correct(fr)
        if max(per_block) <= 12:
            assert ret == nerr and fr == clean
        else:
            assert ret == -1
    fr = bytearray(clean)
    for p in rng.choice(np.arange(58, 518), 60, replace=False):
        fr[p] ^= 0x55
    assert orc.rs41_correct(fr) == -1


def test_rs_symbol_zero_quirk(orc):
    """logtable[1] == n in the reference (rs.c:78-87), so an error at symbol 0 of a block is located at
    index n and never repaired; the count still includes it."""
    clean = bytearray(synth.rs41_frame_bytes())
    fr = bytearray(clean)
    fr[57] ^= 0x21            # data[0] = symbol 0 of block 1
    assert orc.rs41_correct(fr) == 1
    assert fr[57] == clean[57] ^ 0x21


def test_correlator_edge_cases(orc):
    m = synth.MODEMS[synth.RS41]
    sync = m.syncword.to_bytes(8, "big")
    rng = np.random.default_rng(3)
    body = bytes(rng.integers(0, 256, 518 + 8, dtype=np.uint8))
    # exact hit at a bit offset, inverted hit, and earliest-minimum tie-break
    for off in (0, 1, 77, 4143):
        bits = np.unpackbits(np.frombuffer(body, dtype=np.uint8)).copy()
        bits[off:off + 64] = np.unpackbits(np.frombuffer(sync, dtype=np.uint8))
        buf = np.packbits(bits).tobytes()
        assert orc.correlate(m.syncword, 64, buf, 518) == (off, 0)
        inv = bytes(b ^ 0xFF for b in buf)
        assert orc.correlate(m.syncword, 64, inv, 518) == (off, 1)
    bits = np.unpackbits(np.frombuffer(body, dtype=np.uint8)).copy()
    one_err = np.unpackbits(np.frombuffer(sync, dtype=np.uint8)).copy()
    one_err[5] ^= 1
    bits[100:164] = one_err
    bits[900:964] = one_err
    assert orc.correlate(m.syncword, 64, np.packbits(bits).tobytes(), 518) == (100, 0)


def test_checksums(orc):
    msg = b"123456789"
    assert orc.checksum("crc16_ccitt_false", msg) == 0x29B1
    assert orc.checksum("crc16_aug_ccitt", msg) == 0xE5CC
    assert orc.checksum("crc16_modbus", msg) == 0x4B37
    assert orc.checksum("fcs16", bytes([1, 2, 3])) == (6 << 8 | 10)
    assert orc.checksum("crc16_ccitt_false", b"") == 0xFFFF


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref not built in this checkout")
@pytest.mark.parametrize("stype", range(7))
def test_oracle_equals_compiled_reference(orc, stype):
    ref = reflib.RefLib()
    from tests.gpu_util import rec_key
    rb = (synth.MODEMS[stype].frame_bits + 7) // 8
    for ch, chunk in ((3, 1024), (4, 333), (5, 48000)):
        spec = synth.default_spec(stype, ch)
        spec.bit_errors = 6 if stype in (synth.RS41, synth.IMS100, synth.DFM09) else 0
        fm = synth.make_fm(spec, 48000 * 2)
        a, b = ref.frames_run(stype, fm, chunk), orc.frames_run(stype, fm, chunk)
        assert [rec_key(x, rb) for x in a] == [rec_key(x, rb) for x in b]
        assert np.array_equal(ref.demod_bits(stype, fm, chunk), orc.demod_bits(stype, fm, chunk))
    # the public xxx_decode() API returns PARSED exactly once per record
    sd, _ = ref.decode_run(stype, fm, 48000)
    assert len(sd) == len(a)


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref not built in this checkout")
def test_fec_random_vs_reference(orc):
    ref = reflib.RefLib()
    rng = np.random.default_rng(11)
    clean = synth.rs41_frame_bytes()
    for trial in range(60):
        fr = bytearray(clean)
        nerr = int(rng.integers(0, 40))
        for p in rng.choice(np.arange(8, 518), nerr, replace=False):
            fr[p] ^= int(rng.integers(1, 256))
        if trial % 7 == 0:
            fr[56] = 0x0F                       # non-extended frame: 132-byte chunks + zero padding
        a, b = bytearray(fr), bytearray(fr)
        assert ref.rs41_correct(a) == orc.rs41_correct(b)
        assert a == b
    msgs = load("bch_messages.json")
    for trial in range(200):
        m = bytearray(64)
        m[17:63] = bytes(msgs[trial % 6])
        for p in rng.choice(63, int(rng.integers(0, 5)), replace=False):
            m[p] ^= 1
        a, b = bytearray(m), bytearray(m)
        assert ref.bch_fix(a) == orc.bch_fix(b)
        assert a == b


def test_nco_wrap_identity():
    """The AFSK pipeline kernel wraps its NCO phases without double arithmetic (demod_pipe_afsk.cu, NC warp):
    for every float x in [2 pi, 2 pi + 4.28] the reference's (float)fmod((double)x, 2 pi) == (float)((double)x - 2 pi)
    (SD/demod/afsk.c:127-128) equals fl(fl(x - HI) + DL) with HI = the smallest float >= 2 pi and DL = fl(HI - 2 pi).
    Checked exhaustively over all 6.3 M floats of the interval."""
    two_pi = np.float64(2.0 * 3.14159265358979323846)
    hi_bits = 0x40C90FDB
    hi = np.array([hi_bits], dtype=np.uint32).view(np.float32)[0]
    assert np.float64(hi) >= two_pi > np.float64(np.array([hi_bits - 1], dtype=np.uint32).view(np.float32)[0])
    dl = np.float32(np.float64(hi) - two_pi)
    assert dl == np.float32(1.74845553e-07)
    xs = np.arange(hi_bits, hi_bits + (1 << 22) + (1 << 21), dtype=np.uint32).view(np.float32)
    assert xs[-1] > two_pi + 4.28
    ref = np.fmod(xs.astype(np.float64), two_pi).astype(np.float32)
    a = xs - hi
    assert np.all(a.astype(np.float64) == xs.astype(np.float64) - np.float64(hi))      # Sterbenz: exact
    got = (a + dl).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("stype", [synth.IMET4, synth.C50])
def test_ref_afsk_soft_mirror_reproduces_afsk_demod(stype):
    """ref_afsk_soft() (oracle/ref_harness.c) restates the per-sample statements of SD/demod/afsk.c:104-151 around the
    reference's own primitives to expose the symbol VALUES; its hard decisions must be afsk_demod()'s bits."""
    if not reflib.have_ref():
        pytest.skip("compiled reference not built")
    ref = reflib.RefLib()
    n = 48000 * 3
    fm = synth.make_fm(synth.default_spec(stype, 7), n)
    for chunk in (1024, 48000):
        soft, _ = ref.afsk_soft(stype, fm, chunk)
        bits = ref.demod_bits(stype, fm, chunk)
        assert soft.size == bits.size and soft.size > 1000
        assert np.array_equal((soft > 0).astype(np.uint8), bits), (stype, chunk)


def test_discriminator_drift_vs_libm_atan2f():
    """"Parity unpinned" stage (SDR++'s dsp::demod::FM is not under /root/reference): the report SURVEY.md §8c asks for.
    The polynomial discriminator of this repo vs a glibc atan2f one on the same IQ: per-sample difference of the FM
    stream, and what it does downstream in the UNMODIFIED reference chain (soft symbols, frames)."""
    if not (reflib.have_ref() and reflib.have_oracle()):
        pytest.skip("needs oracle/_ref and oracle/_build")
    ref, orc = reflib.RefLib(), reflib.OracleLib()
    n = 48000 * 4
    iq = synth.make_iq(synth.default_spec(synth.RS41, 3), n)
    fm_poly, fm_libm = orc.discriminate(iq), orc.discriminate(iq, libm=True)
    d = np.abs(fm_poly.astype(np.float64) - fm_libm)
    rms = float(np.sqrt(np.mean(fm_libm.astype(np.float64) ** 2)))
    print(f"discriminator: max |poly - atan2f| = {d.max():.3e} ({d.max() / rms:.2e} of the FM rms), "
          f"{np.count_nonzero(fm_poly != fm_libm)} of {n} samples differ in the last bits")
    assert d.max() < 1e-6 * max(1.0, rms * 10)            # 2.4 ulp polynomial vs libm: ~1e-7 of full scale
    sp, _ = ref.gfsk_soft(4800, fm_poly, 48000)
    sl, _ = ref.gfsk_soft(4800, fm_libm, 48000)
    m = min(sp.size, sl.size)
    srms = float(np.sqrt(np.mean(sl[:m].astype(np.float64) ** 2)))
    rel = np.abs(sp[:m].astype(np.float64) - sl[:m]) / srms
    print(f"downstream soft symbols: {np.count_nonzero(sp[:m] != sl[:m])} of {m} differ, "
          f"{np.count_nonzero(rel > 1e-5)} by more than 1e-5 of the rms, max {rel.max():.2e}; "
          f"hard decisions differing: {np.count_nonzero((sp[:m] > 0) != (sl[:m] > 0))}")
    fp = ref.frames_run(synth.RS41, fm_poly, 48000)
    fl = ref.frames_run(synth.RS41, fm_libm, 48000)
    assert [int(r.ok) for r in fp] == [int(r.ok) for r in fl] and sum(int(r.ok) for r in fp) >= 3
    assert [bytes(r.data[:320]) for r in fp if r.ok] == [bytes(r.data[:320]) for r in fl if r.ok]


def _gf_tables(m, poly):
    """alpha[] / logtable[] exactly as rs_init_internal builds them (SD/decode/ecc/rs.c:66-89): n = 2^m - 1 entries of
    successive doublings reduced by `poly`; logtable[1] ends up = n, not 0, because alpha[n] = 1 overwrites it."""
    n = (1 << m) - 1
    ex, lg = [0] * (n + 1), [0] * (n + 1)
    a = 1
    ex[0] = 1
    for i in range(1, n + 1):
        a <<= 1
        if a > n:
            a ^= poly
        ex[i] = a
        lg[a] = i
    return n, ex, lg


def test_gf_polynomial_evaluation_by_independent_terms():
    """The framer kernel evaluates every polynomial over GF(2^m) as an XOR of independent table reads
    exp[(log c_k + k log x) mod n] where the reference uses Horner's rule with its log/antilog multiply
    (rs.c:215-224, 258-270).  Same field element, including through the logtable[1] == n quirk — checked here on
    random polynomials for both fields (GF(256)/0x11D of RS(255,231), GF(64)/0x61 of BCH(63,51)), for the
    binary-coefficient case the BCH syndromes use, and at every evaluation point."""
    rng = np.random.default_rng(11)
    for m, poly in ((8, 0x11D), (6, 0x61)):
        n, ex, lg = _gf_tables(m, poly)
        assert lg[1] == n and ex[0] == 1 and ex[n] == 1

        def mul(x, y):
            return 0 if (x == 0 or y == 0) else ex[(lg[x] + lg[y]) % n]

        def horner(coef, x):                     # coef[k] multiplies x^k
            r = 0
            for c in reversed(coef):
                r = mul(r, x) ^ c
            return r

        def terms(coef, x):
            r = 0
            for k, c in enumerate(coef):
                if c:
                    r ^= ex[(lg[c] + k * lg[x]) % n]
            return r

        for _ in range(60):
            deg = int(rng.integers(1, 25))
            coef = [int(v) for v in rng.integers(0, n + 1, deg + 1)]
            if rng.random() < 0.3:
                coef = [c if rng.random() < 0.5 else 0 for c in coef]
            for x in range(1, n + 1):
                assert horner(coef, x) == terms(coef, x), (m, coef, x)
        # binary message polynomials at the BCH roots alpha^1 .. alpha^4: XOR of alpha^(j k) over the set bits
        if m == 6:
            for _ in range(200):
                msg = int(rng.integers(0, 1 << 62)) | (int(rng.integers(0, 2)) << 62)
                bits = [(msg >> k) & 1 for k in range(63)]
                for j in range(1, 5):
                    z = 1 << j                                   # roots {2, 4, 8, 16}
                    want = horner(bits, z)
                    got = 0
                    for k in range(63):
                        if bits[k]:
                            e = j * k
                            e = (e & 63) + (e >> 6)
                            e = e - 63 if e >= 63 else e
                            got ^= ex[e]
                    assert want == got, (msg, j)


def test_oracle_stream_api_equals_whole_recording_calls():
    """orc_chan_open / push_fm / push_iq / close (the streaming form the batch-ABI stand-in of the host-logic tests is
    built on, tests/cpp/stub_sonde_b200.c) returns exactly the records of orc_frames_run / orc_frames_run_iq with the
    same buffering, record for record, for FM and IQ input and for buffers of irregular lengths (chunk index aside)."""
    import ctypes
    lib = reflib.OracleLib().lib
    lib.orc_chan_open.restype = ctypes.c_void_p
    lib.orc_chan_open.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float]
    for fn in (lib.orc_chan_push_fm, lib.orc_chan_push_iq):
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(reflib.FrameRec), ctypes.c_int]
    lib.orc_chan_close.argtypes = [ctypes.c_void_p]

    def stream(stype, x, lens, iq):
        h = lib.orc_chan_open(stype, 48000, 0.0)
        assert h
        out, pos, k = [], 0, 0
        buf = (reflib.FrameRec * 64)()
        while pos < len(x):
            n = min(lens[k % len(lens)], len(x) - pos)
            part = np.ascontiguousarray(x[pos:pos + n])
            got = (lib.orc_chan_push_iq if iq else lib.orc_chan_push_fm)(h, part.ctypes.data, n, k, buf, 64)
            assert 0 <= got <= 64
            for i in range(got):
                r = reflib.FrameRec()
                ctypes.memmove(ctypes.byref(r), ctypes.byref(buf[i]), ctypes.sizeof(r))
                assert r.chunk == k
                out.append(r)
            pos += n
            k += 1
        lib.orc_chan_close(h)
        return out

    def key(r):
        return (r.type, r.sync_offset, r.inverted, r.status, r.ok, r.aux, r.data_len, bytes(r.raw), bytes(r.data))

    orc = reflib.OracleLib()
    for stype in (synth.RS41, synth.M10, synth.IMS100, synth.IMET4, synth.C50):
        fm = synth.make_fm(synth.default_spec(stype, 5), 48000 * 3).astype(np.float32)
        want = orc.frames_run(stype, fm, 1024)
        got = stream(stype, fm, [1024], False)
        assert len(want) >= 2 and [key(r) for r in got] == [key(r) for r in want] and [r.chunk for r in got] == [r.chunk for r in want]
        ragged = stream(stype, fm, [1, 4099, 77, 20000, 513], False)
        assert [key(r)[3:] for r in ragged] == [key(r)[3:] for r in want]          # same frames whatever the buffering
    iq = synth.make_iq(synth.default_spec(synth.DFM09, 6), 48000 * 3).astype(np.complex64)
    want = orc.frames_run_iq(synth.DFM09, iq, 4096)
    got = stream(synth.DFM09, iq, [4096], True)
    assert len(want) >= 2 and [key(r) for r in got] == [key(r) for r in want]
