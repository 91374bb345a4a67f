"""Known-answer and property tests of the CPU oracle (oracle/), from public standards
(the reference's own tests -- /root/reference/python/qa_*.py:34-37,
lib/qa_liquiddsp.cc:30-36 -- are empty templates, so there are no reference vectors to pin)."""
import zlib

import numpy as np
import pytest

import lqo_py as o
import util

L = o.lib()


def _fec_roundtrip(fs, d, flips=()):
    n = len(d)
    el = L.lqo_fec_enc_len(fs, n)
    e = np.zeros(el + 8, np.uint8)
    L.lqo_fec_encode(fs, n, o._ptr(d), o._ptr(e))
    for byte, mask in flips:
        e[byte % el] ^= mask
    r = np.zeros(n + 8, np.uint8)
    L.lqo_fec_decode(fs, n, o._ptr(e), o._ptr(r))
    return e[:el], r[:n]


def test_crc32_matches_zlib_check_value():
    m = np.frombuffer(b"123456789", np.uint8).copy()
    assert L.lqo_crc_key(6, o._ptr(m), 9) == 0xCBF43926
    rng = np.random.default_rng(0)
    for n in (1, 7, 64, 1500):
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert L.lqo_crc_key(6, o._ptr(b), n) == zlib.crc32(b.tobytes())


def test_crc_widths_and_sensitivity():
    rng = np.random.default_rng(1)
    b = rng.integers(0, 256, 100, dtype=np.uint8)
    for scheme, bits in ((3, 8), (4, 16), (5, 24), (2, 8)):
        k = L.lqo_crc_key(scheme, o._ptr(b), 100)
        assert k < (1 << bits)
        c = b.copy(); c[17] ^= 0x04
        assert L.lqo_crc_key(scheme, o._ptr(c), 100) != k


def test_msequence_period_and_balance():
    st = (np.zeros(8, np.uint32))
    L.lqo_mseq_init(o._ptr(st), 7, 0x89, 1)
    bits = [L.lqo_mseq_advance(o._ptr(st)) for _ in range(254)]
    assert bits[:127] == bits[127:]            # period 2^7 - 1
    assert sum(bits[:127]) == 64               # balance property
    # no shorter period
    for p in range(1, 127):
        assert bits[:127] != bits[p:p + 127]


def test_conv_v27_impulse_response():
    # a single '1' followed by zeros reads the generator polynomials 0x6d, 0x4f out of the encoder
    d = np.array([0x80], np.uint8)
    e, _ = _fec_roundtrip(11, d)
    bits = np.unpackbits(e)[:14].reshape(7, 2)
    g0 = [(0x6d >> i) & 1 for i in range(7)]
    g1 = [(0x4f >> i) & 1 for i in range(7)]
    assert bits[:, 0].tolist() == g0 and bits[:, 1].tolist() == g1
    assert len(e) == 2 * 1 + 2


def test_conv_v29_impulse_response():
    d = np.array([0x80], np.uint8)
    e, _ = _fec_roundtrip(12, d)
    bits = np.unpackbits(e)[:18].reshape(9, 2)
    assert bits[:, 0].tolist() == [(0x1af >> i) & 1 for i in range(9)]
    assert bits[:, 1].tolist() == [(0x11d >> i) & 1 for i in range(9)]


@pytest.mark.parametrize("fs", [11, 12, 15, 16, 17, 18, 19, 20, 21, 26])
def test_viterbi_corrects_scattered_errors(fs):
    rng = np.random.default_rng(fs)
    d = rng.integers(0, 256, 120, dtype=np.uint8)
    nflip = 6 if fs in (11, 12) else 2
    flips = [(40 * (i + 1), 0x10) for i in range(nflip)]
    _, r = _fec_roundtrip(fs, d, flips)
    assert np.array_equal(r, d)


def test_rs_255_223_corrects_16_and_detects_17():
    rng = np.random.default_rng(5)
    blk = np.zeros(255, np.uint8)
    blk[:223] = rng.integers(0, 256, 223)
    par = np.zeros(32, np.uint8)
    L.lqo_rs_encode_block(o._ptr(blk), 0, o._ptr(par))
    blk[223:] = par
    clean = blk.copy()
    assert L.lqo_rs_decode_block(o._ptr(clean), 0) == 0
    for nerr in (1, 8, 16):
        b = blk.copy()
        pos = rng.choice(255, nerr, replace=False)
        b[pos] ^= rng.integers(1, 256, nerr).astype(np.uint8)
        assert L.lqo_rs_decode_block(o._ptr(b), 0) == nerr
        assert np.array_equal(b, blk)
    b = blk.copy()
    pos = rng.choice(255, 17, replace=False)
    b[pos] ^= rng.integers(1, 256, 17).astype(np.uint8)
    assert L.lqo_rs_decode_block(o._ptr(b), 0) == -1


def test_rs_shortened_blocks_and_liquid_block_split():
    # cfg-3: 3008 bytes -> 14 blocks of 215+32
    assert L.lqo_fec_enc_len(27, 3008) == 14 * 247
    rng = np.random.default_rng(6)
    d = rng.integers(0, 256, 3008, dtype=np.uint8)
    flips = [(247 * b + 3 * k, 0xff) for b in range(14) for k in range(16)]
    _, r = _fec_roundtrip(27, d, flips)
    assert np.array_equal(r, d)


def test_hamming84_table_and_single_error():
    tab = [0x00, 0xd2, 0x55, 0x87, 0x99, 0x4b, 0xcc, 0x1e, 0xe1, 0x33, 0xb4, 0x66, 0x78, 0xaa, 0x2d, 0xff]
    d = np.arange(16, dtype=np.uint8)
    e, _ = _fec_roundtrip(5, d)
    assert e[1::2].tolist() == tab
    for bit in range(8):
        _, r = _fec_roundtrip(5, d, [(5, 1 << bit)])
        assert np.array_equal(r, d)


@pytest.mark.parametrize("fs,nerr", [(4, 1), (6, 1), (7, 3), (8, 1), (9, 1), (10, 1), (2, 1), (3, 2)])
def test_block_codes_correct_their_design_errors(fs, nerr):
    rng = np.random.default_rng(100 + fs)
    for n in (1, 2, 3, 5, 8, 16, 33):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        el = L.lqo_fec_enc_len(fs, n)
        e = np.zeros(el + 8, np.uint8)
        L.lqo_fec_encode(fs, n, o._ptr(d), o._ptr(e))
        if fs in (2, 3):          # repetition: corrupt whole copies of one byte
            for k in range(nerr):
                e[k * n] ^= 0xff
        else:                     # one codeword (the first) gets nerr bit errors
            for b in rng.choice(7 if fs == 4 else 12, nerr, replace=False):
                e[b // 8] ^= 0x80 >> (b % 8)
        r = np.zeros(n + 8, np.uint8)
        L.lqo_fec_decode(fs, n, o._ptr(e), o._ptr(r))
        assert np.array_equal(r[:n], d), (fs, n)


@pytest.mark.parametrize("fs,nb,R", [(8, 2, 6), (9, 4, 7), (10, 8, 8)])
def test_secded_matrices_are_hsiao_codes_and_decode_as_liquid_does(fs, nb, R):
    """The parity matrices are restated from liquid-dsp (fec_secded2216 / 3932 / 7264).  What can be checked without
    liquid: they are Hsiao SEC-DED codes (every column distinct, of odd weight >= 3, inside R bits), so every single
    error is corrected and every double error leaves the data as received."""
    import ctypes as C
    L.lqo_secded_columns.argtypes = [C.c_uint, C.c_void_p]
    col = np.zeros(8 * nb, np.uint8)
    L.lqo_secded_columns(nb, o._ptr(col))
    w = np.array([bin(int(c)).count("1") for c in col])
    assert len(set(col.tolist())) == 8 * nb and np.all(w % 2 == 1) and np.all(w >= 3) and np.all(col < (1 << R))
    assert sorted(set(w.tolist())) == ([3, 5] if nb == 8 else [3])        # (72,64): 56 columns of weight 3, 8 of weight 5
    rng = np.random.default_rng(fs)
    d = rng.integers(0, 256, nb, dtype=np.uint8)
    e = np.zeros(nb + 1, np.uint8)
    L.lqo_fec_encode(fs, nb, o._ptr(d), o._ptr(e))
    assert e[0] == np.bitwise_xor.reduce(col[np.unpackbits(d).astype(bool)], initial=0) and np.array_equal(e[1:], d)
    r = np.zeros(nb + 8, np.uint8)
    for b in range(8 * (nb + 1)):                    # every single error, parity byte included (its unused high bits too)
        x = e.copy(); x[b // 8] ^= 0x80 >> (b % 8)
        L.lqo_fec_decode(fs, nb, o._ptr(x), o._ptr(r))
        assert np.array_equal(r[:nb], d), b
    for _ in range(200):                             # double errors inside the code word: data left as received
        b1, b2 = rng.choice(np.arange(8 - R, 8 * (nb + 1)), 2, replace=False)
        x = e.copy(); x[b1 // 8] ^= 0x80 >> (b1 % 8); x[b2 // 8] ^= 0x80 >> (b2 % 8)
        L.lqo_fec_decode(fs, nb, o._ptr(x), o._ptr(r))
        assert np.array_equal(r[:nb], x[1:])


def test_secded_detects_double_error_without_miscorrecting_other_bits():
    d = np.arange(8, dtype=np.uint8) * 17
    e = np.zeros(16, np.uint8)
    L.lqo_fec_encode(10, 8, o._ptr(d), o._ptr(e))
    e[2] ^= 0x41          # two bit errors in one data byte
    r = np.zeros(16, np.uint8)
    L.lqo_fec_decode(10, 8, o._ptr(e), o._ptr(r))
    diff = np.unpackbits(r[:8] ^ d).sum()
    assert diff == 2      # left as received, not made worse


@pytest.mark.parametrize("n", [2, 3, 8, 27, 54, 259, 1001, 3458])
def test_interleaver_is_invertible_and_scatters(n):
    rng = np.random.default_rng(n)
    d = rng.integers(0, 256, n, dtype=np.uint8)
    x = d.copy()
    L.lqo_interleave(o._ptr(x), n, 4)
    y = x.copy()
    L.lqo_deinterleave(o._ptr(y), n, 4)
    assert np.array_equal(y, d)
    if n > 8:
        assert (x != d).mean() > 0.5
    assert sorted(np.unpackbits(x).tolist()) == sorted(np.unpackbits(d).tolist())   # a bit permutation


def test_scrambler_is_an_involution():
    d = np.arange(37, dtype=np.uint8)
    x = d.copy()
    L.lqo_scramble(o._ptr(x), 37)
    assert x[:4].tolist() == [0xb4, 0x6a ^ 1, 0x8b ^ 2, 0xc5 ^ 3]
    L.lqo_scramble(o._ptr(x), 37)
    assert np.array_equal(x, d)


@pytest.mark.parametrize("f0", util.INNER)
@pytest.mark.parametrize("f1", util.OUTER)
def test_packetizer_roundtrip_all_block_api_code_pairs(f0, f1):
    rng = np.random.default_rng(f0 * 100 + f1)
    for n in (1, 17, 256):
        msg = rng.integers(0, 256, n, dtype=np.uint8)
        el = L.lqo_packetizer_enc_len(n, util.CRC24, f0, f1)
        pkt = np.zeros(el + 8, np.uint8)
        L.lqo_packetizer_encode(n, util.CRC24, f0, f1, o._ptr(msg), o._ptr(pkt))
        out = np.zeros(n + 8, np.uint8)
        assert L.lqo_packetizer_decode(n, util.CRC24, f0, f1, o._ptr(pkt), o._ptr(out)) == 1
        assert np.array_equal(out[:n], msg)
        pkt[el // 2] ^= 0xff
        pkt[el // 3] ^= 0xff
        ok = L.lqo_packetizer_decode(n, util.CRC24, f0, f1, o._ptr(pkt), o._ptr(out))
        if f0 == 1 and f1 == 1:
            assert ok == 0                       # no FEC: the CRC must catch it


def test_frame_lengths_of_baseline_configs():
    # cfg1 2690, cfg3 28282, cfg5 6630 samples
    assert len(o.tx_frame(util.PSK4, util.CRC24, 1, 1, np.zeros(256, np.uint8))) == 2690
    assert L.lqo_qpm_frame_len(1500, util.CRC24, 11, 27, util.PSK4) == 13832
    assert 2 * (64 + 231 + 13832 + 14) == 28282
    assert L.lqo_qpm_frame_len(1500, util.CRC24, 1, 1, util.QAM16) == 3006
    assert 2 * (64 + 231 + 3006 + 14) == 6630


def test_fft_matches_numpy():
    rng = np.random.default_rng(3)
    for n in (32, 512):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        y = np.zeros(n, np.complex64)
        L.lqo_fft(o._ptr(x), o._ptr(y), n, 1)
        assert np.allclose(y, np.fft.fft(x), rtol=1e-4, atol=1e-3)
        L.lqo_fft(o._ptr(x), o._ptr(y), n, -1)
        assert np.allclose(y, np.fft.ifft(x) * n, rtol=1e-4, atol=1e-3)


def test_filters_are_root_nyquist_pairs():
    h = np.zeros(32, np.float32)
    L.lqo_interp_taps(2, 7, np.float32(0.3), o._ptr(h))
    assert abs(float((h[:29] ** 2).sum()) - 2.0) < 1e-5
    c = np.convolve(h[:29], h[:29])               # matched pair: ISI at even lags is small
    centre = len(c) // 2
    isi = np.delete(c[centre % 2::2], centre // 2)
    assert np.abs(isi).max() < 0.02 * c[centre]
    banks = np.zeros((32, 28), np.float32)
    L.lqo_pfb_rnyquist(32, 2, 7, np.float32(0.3), o._ptr(banks))
    assert np.allclose((banks ** 2).sum(axis=1), 2.0, atol=0.05)


def test_nco_constrain_wraps():
    assert L.lqo_nco_constrain(0.0) == 0
    assert abs(int(L.lqo_nco_constrain(np.pi)) - 2 ** 31) < 512
    assert abs(int(L.lqo_nco_constrain(-np.pi / 2)) - 3 * 2 ** 30) < 512
    assert L.lqo_nco_constrain(np.float32(2 * np.pi)) < 1024 or L.lqo_nco_constrain(np.float32(2 * np.pi)) > 2 ** 32 - 1024
