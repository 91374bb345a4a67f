"""A/B harness: genuine liquid-dsp (dlopen) against oracle/ -- the hook that pins parity.

liquid-dsp is an un-vendored, un-pinned dependency of the reference (`-lliquid`, /root/reference/lib/CMakeLists.txt:33)
and is not installable in the build container, so `oracle/` is a restatement whose wire compatibility with real
liquid-dsp is UNPINNED.  On a machine that has liquid-dsp (>= 1.3.1, built with libfec):

    LIQUID_SO=/usr/local/lib/libliquid.so python -m pytest tests/test_liquid_ab.py -q

runs liquid's own functions -- the names the reference calls at /root/reference/lib/flex_rx_impl.cc:49,213,
/root/reference/lib/flex_tx_impl.cc:56,198-201 and /root/reference/lib/frame_detector_cc_impl.cc:47-55,77 plus the
modules underneath -- beside the oracle on the same inputs.  Each test names the docs/FRAME_FORMAT.md choice (the
items marked there as uncertain) it settles.  Without LIQUID_SO every test here is skipped; nothing in the GPU box
runs or needs it.  CPU only (no gpu marker): the CUDA path is tied to the oracle by tests/test_gpu_*.py.
"""
import ctypes as C
import os

import numpy as np
import pytest

import lqo_py as o
import util

LIQUID_SO = os.environ.get("LIQUID_SO")
pytestmark = pytest.mark.skipif(not LIQUID_SO, reason="LIQUID_SO not set: no genuine liquid-dsp to compare with (parity unpinned)")

u8p, vp, u32 = C.POINTER(C.c_uint8), C.c_void_p, C.c_uint


class Cf(C.Structure):               # float complex by value (one SSE eightbyte on x86-64, as the C ABI passes it)
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class Props(C.Structure):            # flexframegenprops_s
    _fields_ = [("check", u32), ("fec0", u32), ("fec1", u32), ("mod_scheme", u32)]


class Stats(C.Structure):            # framesyncstats_s
    _fields_ = [("evm", C.c_float), ("rssi", C.c_float), ("cfo", C.c_float), ("framesyms", vp), ("num_framesyms", u32),
                ("mod_scheme", u32), ("mod_bps", u32), ("check", u32), ("fec0", u32), ("fec1", u32)]


CB = C.CFUNCTYPE(C.c_int, u8p, C.c_int, u8p, u32, C.c_int, Stats, vp)

_liq = None


def liq():
    global _liq
    if _liq is None:
        L = C.CDLL(LIQUID_SO)
        L.msequence_create.restype = vp
        L.msequence_create.argtypes = [u32, u32, u32]
        L.msequence_advance.restype = u32
        L.msequence_advance.argtypes = [vp]
        L.msequence_destroy.argtypes = [vp]
        L.liquid_firdes_arkaiser.argtypes = [u32, u32, C.c_float, C.c_float, vp]
        L.crc_generate_key.restype = u32
        L.crc_generate_key.argtypes = [C.c_int, vp, u32]
        L.scramble_data.argtypes = [vp, u32]
        L.fec_create.restype = vp
        L.fec_create.argtypes = [C.c_int, vp]
        L.fec_destroy.argtypes = [vp]
        L.fec_get_enc_msg_length.restype = u32
        L.fec_get_enc_msg_length.argtypes = [C.c_int, u32]
        L.fec_encode.argtypes = [vp, u32, vp, vp]
        L.fec_decode.argtypes = [vp, u32, vp, vp]
        L.interleaver_create.restype = vp
        L.interleaver_create.argtypes = [u32]
        L.interleaver_encode.argtypes = [vp, vp, vp]
        L.interleaver_destroy.argtypes = [vp]
        L.packetizer_create.restype = vp
        L.packetizer_create.argtypes = [u32, C.c_int, C.c_int, C.c_int]
        L.packetizer_get_enc_msg_len.restype = u32
        L.packetizer_get_enc_msg_len.argtypes = [vp]
        L.packetizer_encode.argtypes = [vp, vp, vp]
        L.packetizer_destroy.argtypes = [vp]
        # liquid 1.3.x: modem_*; 1.4+: modemcf_* (same arguments)
        pre = "modem" if hasattr(L, "modem_create") else "modemcf"
        L.m_create, L.m_modulate, L.m_destroy = getattr(L, pre + "_create"), getattr(L, pre + "_modulate"), getattr(L, pre + "_destroy")
        L.m_create.restype = vp
        L.m_create.argtypes = [C.c_int]
        L.m_modulate.argtypes = [vp, u32, C.POINTER(Cf)]
        L.m_destroy.argtypes = [vp]
        L.flexframegenprops_init_default.argtypes = [C.POINTER(Props)]
        L.flexframegen_create.restype = vp
        L.flexframegen_create.argtypes = [C.POINTER(Props)]
        L.flexframegen_destroy.argtypes = [vp]
        L.flexframegen_assemble.argtypes = [vp, vp, vp, u32]
        L.flexframegen_getframelen.restype = u32
        L.flexframegen_getframelen.argtypes = [vp]
        L.flexframegen_write_samples.restype = C.c_int
        L.flexframegen_write_samples.argtypes = [vp, vp, u32]
        L.flexframesync_create.restype = vp
        L.flexframesync_create.argtypes = [CB, vp]
        L.flexframesync_execute.argtypes = [vp, vp, u32]
        L.flexframesync_destroy.argtypes = [vp]
        L.qdetector_cccf_create_linear.restype = vp
        L.qdetector_cccf_create_linear.argtypes = [vp, u32, C.c_int, u32, u32, C.c_float]
        L.qdetector_cccf_set_threshold.argtypes = [vp, C.c_float]
        L.qdetector_cccf_execute.restype = vp
        L.qdetector_cccf_execute.argtypes = [vp, Cf]
        L.qdetector_cccf_destroy.argtypes = [vp]
        for g in ("tau", "gamma", "dphi", "phi"):
            f = getattr(L, "qdetector_cccf_get_" + g)
            f.restype = C.c_float
            f.argtypes = [vp]
        _liq = L
    return _liq


def liquid_tx_frame(ms, check, fec0, fec1, payload, header=None):
    """The reference's own sequence: /root/reference/lib/flex_tx_impl.cc:51-59 (props, 14-byte header), :198-201."""
    L = liq()
    p = Props()
    L.flexframegenprops_init_default(C.byref(p))
    p.check, p.fec0, p.fec1, p.mod_scheme = check, fec0, fec1, ms
    fg = L.flexframegen_create(C.byref(p))
    hdr = np.zeros(14, np.uint8) if header is None else np.ascontiguousarray(header, np.uint8)
    pl = np.ascontiguousarray(payload, np.uint8)
    L.flexframegen_assemble(fg, hdr.ctypes.data, pl.ctypes.data, len(pl))
    n = L.flexframegen_getframelen(fg)
    buf = np.zeros(n, np.complex64)
    L.flexframegen_write_samples(fg, buf.ctypes.data, n)
    L.flexframegen_destroy(fg)
    return buf


def liquid_rx_capture(x):
    """The reference's receive loop: flexframesync_execute in 256-sample chunks (/root/reference/lib/flex_rx_impl.cc:212-215)
    with a callback that copies what flex_rx_impl::callback reads (:182-201)."""
    L = liq()
    out = []

    def cb(header, hv, payload, plen, pv, stats, ud):
        syms = np.zeros(0, np.complex64)
        if stats.framesyms and stats.num_framesyms:
            syms = np.frombuffer((C.c_float * (2 * stats.num_framesyms)).from_address(stats.framesyms), np.complex64).copy()
        out.append(dict(header=bytes(bytearray(header[:20])), header_valid=int(hv), payload_valid=int(pv), payload_len=int(plen),
                        payload=bytes(bytearray(payload[:plen])) if (hv and payload) else b"", framesyms=syms,
                        num_framesyms=int(stats.num_framesyms), evm=stats.evm, rssi=stats.rssi, cfo=stats.cfo,
                        mod_scheme=int(stats.mod_scheme), mod_bps=int(stats.mod_bps), check=int(stats.check),
                        fec0=int(stats.fec0), fec1=int(stats.fec1)))
        return 0
    keep = CB(cb)
    fs = L.flexframesync_create(keep, None)
    x = np.ascontiguousarray(x, np.complex64)
    for i in range(0, len(x) // 256 * 256, 256):
        L.flexframesync_execute(fs, x[i:i + 256].ctypes.data, 256)
    L.flexframesync_destroy(fs)
    return out


# ----------------------------------------------------------------------------- building blocks
def test_ab_msequence():
    """generator / state conventions of msequence_create(7, 0x0089, 1) (the preamble source)."""
    L, Lo = liq(), o.lib()
    for (m, g, a) in [(7, 0x0089, 1), (4, 0x13, 1), (6, 0x43, 1)]:
        q = L.msequence_create(m, g, a)
        st = (C.c_uint * 8)()
        Lo.lqo_mseq_init(C.byref(st), m, g, a)
        bits_l = [L.msequence_advance(q) for _ in range(2 * ((1 << m) - 1))]
        bits_o = [Lo.lqo_mseq_advance(C.byref(st)) for _ in range(2 * ((1 << m) - 1))]
        L.msequence_destroy(q)
        assert bits_l == bits_o, (m, g, a)


@pytest.mark.parametrize("beta", [0.25, 0.3])
def test_ab_arkaiser_taps(beta):
    """A.4: the ARKAISER design constants (interpolator and matched-filter prototypes)."""
    L, Lo = liq(), o.lib()
    for (k, m) in [(2, 7), (64, 7)]:
        n = 2 * k * m + 1
        hl, ho = np.zeros(n, np.float32), np.zeros(n, np.float32)
        L.liquid_firdes_arkaiser(k, m, beta, 0.0, hl.ctypes.data)
        Lo.lqo_firdes_arkaiser(k, m, np.float32(beta), np.float32(0.0), o._ptr(ho))
        assert np.allclose(hl, ho, rtol=0, atol=2e-6), float(np.abs(hl - ho).max())


@pytest.mark.parametrize("check", [2, 3, 4, 5, 6])
def test_ab_crc(check):
    """A.7: CRC polynomials / reflection / presets, and the checksum."""
    L, Lo = liq(), o.lib()
    rng = np.random.default_rng(check)
    for n in (1, 9, 20, 257):
        msg = rng.integers(0, 256, n, dtype=np.uint8)
        assert L.crc_generate_key(check, msg.ctypes.data, n) == Lo.lqo_crc_key(check, o._ptr(msg), n)


def test_ab_scrambler():
    """A.7: whitening masks b4 6a 8b c5."""
    L, Lo = liq(), o.lib()
    a = np.arange(64, dtype=np.uint8)
    b = a.copy()
    L.scramble_data(a.ctypes.data, 64)
    Lo.lqo_scramble(o._ptr(b), 64)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [27, 54, 100, 414, 3008, 3458])
def test_ab_interleaver(n):
    """FRAME_FORMAT 'Interleaver' (start column cols/3, four masked passes)."""
    L, Lo = liq(), o.lib()
    msg = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    q = L.interleaver_create(n)
    enc = np.zeros(n, np.uint8)
    L.interleaver_encode(q, msg.ctypes.data, enc.ctypes.data)
    L.interleaver_destroy(q)
    mine = msg.copy()
    Lo.lqo_interleave(o._ptr(mine), n, 4)
    assert np.array_equal(enc, mine)


# every fec_scheme the oracle implements: 1 none, 2-3 rep, 4-6 Hamming, 7 Golay, 8-10 SECDED, 11-12 conv, 15-26 punctured, 27 RS
@pytest.mark.parametrize("fs", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27])
def test_ab_fec_encode(fs):
    """Settles the block-code tables marked uncertain in FRAME_FORMAT (SECDED(22,16)/(39,32)/(72,64) -- the last one
    protects EVERY header --, Golay(24,12), Hamming(12,8)), the v29 puncturing matrices and the RS block split."""
    L, Lo = liq(), o.lib()
    rng = np.random.default_rng(100 + fs)
    for n in (1, 8, 24, 203, 1503):
        msg = rng.integers(0, 256, n, dtype=np.uint8)
        nl = L.fec_get_enc_msg_length(fs, n)
        assert nl == Lo.lqo_fec_enc_len(fs, n), (fs, n)
        q = L.fec_create(fs, None)
        el, eo = np.zeros(nl, np.uint8), np.zeros(nl, np.uint8)
        L.fec_encode(q, n, msg.ctypes.data, el.ctypes.data)
        Lo.lqo_fec_encode(fs, n, o._ptr(msg), o._ptr(eo))
        assert np.array_equal(el, eo), (fs, n)
        # decode each other's codewords with a few bit errors injected
        bad = el.copy()
        if fs != 1 and nl >= 8:
            bad[nl // 2] ^= 0x10
        dl, do_ = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        L.fec_decode(q, n, bad.ctypes.data, dl.ctypes.data)
        Lo.lqo_fec_decode(fs, n, o._ptr(bad), o._ptr(do_))
        L.fec_destroy(q)
        assert np.array_equal(dl, do_), (fs, n)


@pytest.mark.parametrize("f0,f1", [(1, 1), (11, 27), (15, 7), (20, 10), (12, 27), (1, 6)])
def test_ab_packetizer(f0, f1):
    """A.7: CRC big-endian -> whitening -> fec0 -> interleave -> fec1 -> interleave."""
    L, Lo = liq(), o.lib()
    for n in (1, 256, 1500):
        msg = np.random.default_rng(n + f0).integers(0, 256, n, dtype=np.uint8)
        q = L.packetizer_create(n, 5, f0, f1)
        nl = L.packetizer_get_enc_msg_len(q)
        assert nl == Lo.lqo_packetizer_enc_len(n, 5, f0, f1)
        pl, po = np.zeros(nl, np.uint8), np.zeros(nl, np.uint8)
        L.packetizer_encode(q, msg.ctypes.data, pl.ctypes.data)
        Lo.lqo_packetizer_encode(n, 5, f0, f1, o._ptr(msg), o._ptr(po))
        L.packetizer_destroy(q)
        assert np.array_equal(pl, po)


@pytest.mark.parametrize("ms", util.MODS + [30, 31, 39, 40])
def test_ab_modem_constellations(ms):
    """A.6: symbol maps (Gray coding, QAM axis split and scale) of every scheme the block API reaches + QAM128/256, BPSK, QPSK."""
    L, Lo = liq(), o.lib()
    q = L.m_create(ms)
    st = (C.c_uint8 * 65536)()
    Lo.lqo_modem_init(C.byref(st), ms)
    Lo.lqo_modem_modulate.restype = Cf
    Lo.lqo_modem_modulate.argtypes = [vp, u32]
    bps = Lo.lqo_modem_bps(ms)
    for s in range(1 << bps):
        y = Cf()
        L.m_modulate(q, s, C.byref(y))
        if 9 <= ms <= 16:                  # DPSK has memory: compare phase increments from a fresh modem each time
            L.m_destroy(q)
            q = L.m_create(ms)
            Lo.lqo_modem_init(C.byref(st), ms)
            L.m_modulate(q, s, C.byref(y))
        r = Lo.lqo_modem_modulate(C.byref(st), s)
        assert abs(y.re - r.re) <= 2e-6 and abs(y.im - r.im) <= 2e-6, (ms, s)
    L.m_destroy(q)


# ----------------------------------------------------------------------------- the three objects the reference uses
@pytest.mark.parametrize("ms,f0,f1,n", [(util.PSK4, 1, 1, 256), (util.PSK4, 11, 27, 1500), (util.QAM16, 1, 1, 1500),
                                        (util.PSK2, 15, 7, 64), (util.QAM64, 20, 10, 300), (util.DPSK4, 1, 4, 100)])
def test_ab_flexframegen(ms, f0, f1, n):
    """The whole transmit side: settles the TX interpolator beta (0.25 here), header layout and coding (protocol byte 102,
    CRC-32 + SECDED(72,64) + Hamming(8,4)), pilot sequence and spacing, frame length."""
    pl = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    a = liquid_tx_frame(ms, util.CRC24, f0, f1, pl)
    b = o.tx_frame(ms, util.CRC24, f0, f1, pl)
    assert len(a) == len(b), (len(a), len(b))
    assert np.allclose(a, b, rtol=0, atol=5e-6), float(np.abs(a - b).max())


def test_ab_qdetector():
    """qdetector_cccf with the reference's own arguments (/root/reference/lib/frame_detector_cc_impl.cc:47-55, .h:34-36):
    detection positions and tau / gamma / dphi / phi estimates; settles the RX template beta and the seek/align details."""
    L = liq()
    ms = L.msequence_create(7, 0x0089, 1)
    pn = np.zeros(64, np.complex64)
    for i in range(64):
        re = np.sqrt(0.5) if L.msequence_advance(ms) else -np.sqrt(0.5)
        im = np.sqrt(0.5) if L.msequence_advance(ms) else -np.sqrt(0.5)
        pn[i] = re + 1j * im
    L.msequence_destroy(ms)
    q = L.qdetector_cccf_create_linear(pn.ctypes.data, 64, 9, 2, 7, 0.3)          # 9 = LIQUID_FIRFILT_ARKAISER
    L.qdetector_cccf_set_threshold(q, 0.45)
    rng = np.random.default_rng(8)
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 64, dtype=np.uint8)) for _ in range(6)]
    cap = util.build_capture(frames, rng, [1500] * 6, snr_db=12.0, cfo=0.03, tau=0.3)
    mine = o.detect_capture(cap, 0.3, 0.45)
    theirs = []
    for i, v in enumerate(cap):
        if L.qdetector_cccf_execute(q, Cf(float(v.real), float(v.imag))):
            theirs.append(dict(at=i, tau=L.qdetector_cccf_get_tau(q), gamma=L.qdetector_cccf_get_gamma(q),
                               dphi=L.qdetector_cccf_get_dphi(q), phi=L.qdetector_cccf_get_phi(q)))
    L.qdetector_cccf_destroy(q)
    assert len(theirs) == len(mine) >= 6
    for t, m in zip(theirs, mine):
        for k, mk in (("tau", "tau_hat"), ("gamma", "gamma_hat"), ("dphi", "dphi_hat"), ("phi", "phi_hat")):
            assert abs(t[k] - m[mk]) <= 2e-5 + 1e-3 * abs(m[mk]), (k, t[k], m[mk])


@pytest.mark.parametrize("ms,f0,f1,n,snr", [(util.PSK4, 1, 1, 256, 30.0), (util.PSK4, 11, 27, 1500, 9.0), (util.QAM16, 1, 1, 1500, 22.0),
                                             (util.PSK8, 15, 7, 300, 16.0), (util.PSK4, 11, 27, 1500, 3.0)])
def test_ab_flexframesync(ms, f0, f1, n, snr):
    """The whole receive side on an impaired capture of liquid-generated frames: header / payload bytes and flags equal,
    EVM / RSSI / CFO within 1e-3 relative, constellation within 1e-4 (liquid's libm vs the pinned arg / sincos)."""
    rng = np.random.default_rng(int(snr * 10) + n)
    pls = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(4)]
    frames = [liquid_tx_frame(ms, util.CRC24, f0, f1, p) for p in pls]
    cap = util.build_capture(frames, rng, [1800] * 4, snr_db=snr, cfo=0.012, tau=-0.3, gain=0.8)
    a, b = liquid_rx_capture(cap), o.rx_capture(cap)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for k in ("header_valid", "payload_valid", "payload_len", "num_framesyms", "mod_scheme", "mod_bps", "check", "fec0", "fec1"):
            assert x[k] == y[k], (k, x[k], y[k])
        if y["header_valid"]:
            assert x["header"][:20] == y["header"] and x["payload"] == y["payload"]
            for k in ("evm", "rssi", "cfo"):
                assert abs(x[k] - y[k]) <= 2e-5 + 1e-3 * abs(y[k]), (k, x[k], y[k])
            assert np.allclose(x["framesyms"], y["framesyms"], rtol=0, atol=1e-4)
