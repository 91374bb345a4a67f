"""The three GNU Radio blocks (flex_tx, flex_rx, frame_detector_cc) driven through their block interface
on the test shim (gr-liquiddsp_b200/host/gr_shim), and the liquid-dsp-signature veneer
(include/lqb200_liquid.h).  Contract checked: SURVEY.md section 8(b1) -- make() arguments, stream
signatures, message ports, publish order, dict keys, index tables, error behaviour
(/root/reference/lib/flex_rx_impl.cc:44-63,204-254; lib/flex_tx_impl.cc:42-65,183-218;
lib/frame_detector_cc_impl.cc:41-56,67-97)."""
import ctypes as C
import os

import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLK = os.path.join(ROOT, "gr-liquiddsp_b200", "lib", "liblqb_blocks_test.so")


def _lib():
    capi.lib()
    L = C.CDLL(BLK)
    vp = C.c_void_p
    for f in ("blk_make_flex_tx", "blk_make_flex_rx", "blk_make_flex_rx_multi", "blk_make_frame_detector"):
        getattr(L, f).restype = vp
    L.blk_make_flex_tx.argtypes = [C.c_uint] * 3
    L.blk_make_flex_rx_multi.argtypes = [C.c_uint]
    L.blk_destroy.argtypes = [vp]
    L.blk_name.restype = C.c_char_p
    L.blk_name.argtypes = [vp]
    L.blk_error.restype = C.c_char_p
    L.blk_error.argtypes = [vp]
    L.blk_output_multiple.argtypes = [vp]
    L.blk_sig.argtypes = [vp, C.c_int, C.c_int]
    L.blk_ports.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.blk_work.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.blk_post_pdu.argtypes = [vp, C.c_char_p, vp, C.c_int]
    L.blk_post_dict.argtypes = [vp, C.c_char_p, C.c_char_p, vp, C.c_int]
    L.blk_pending.argtypes = [vp]
    L.blk_pop.argtypes = [vp, C.c_char_p, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    L.blk_rx_index.argtypes = [C.c_int, C.c_uint]
    L.blk_tx_props.argtypes = [vp, vp]
    L.blk_rx_decode_capture.restype = C.c_long
    L.blk_det_detect_capture.restype = C.c_long
    L.blk_det_detect_capture.argtypes = [vp, vp, C.c_ulong, vp, C.c_ulong, C.c_uint, C.c_uint, C.c_uint]
    L.blk_det_frames.restype = C.c_ulong
    L.blk_det_frames.argtypes = [vp]
    L.blk_rx_decode_capture.argtypes = [vp, vp, C.c_ulong, C.c_uint, C.c_uint, C.c_uint]
    return L


def ports(L, h, out):
    buf = C.create_string_buffer(256)
    L.blk_ports(h, out, buf, 256)
    return buf.value.decode().split(",") if buf.value else []


def pop_all(L, h):
    msgs = []
    port = C.create_string_buffer(64)
    data = np.zeros(1 << 20, np.uint8)
    n = C.c_int(0)
    while True:
        k = L.blk_pop(h, port, 64, data.ctypes.data, data.nbytes, C.byref(n))
        if k == 0:
            break
        if k == 1:
            msgs.append((port.value.decode(), data[:8 * n.value].view(np.complex64).copy()))
        elif k == 2:
            msgs.append((port.value.decode(), bytes(data[:n.value])))
        else:
            txt = bytes(data[:4096]).split(b"\0")[0].decode()
            msgs.append((port.value.decode(), dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in txt.split(";") if kv)))
    return msgs


def test_index_tables_match_the_reference_numbering():
    L = _lib()
    for i, ms in enumerate(util.MODS):
        assert L.blk_rx_index(0, ms) == i
    for i, fs in enumerate(util.INNER):
        assert L.blk_rx_index(1, fs) == i
    for i, fs in enumerate(util.OUTER):
        assert L.blk_rx_index(2, fs) == i
    # additive extension indices sit AFTER the reference's ranges; schemes in neither stay -1
    assert L.blk_rx_index(0, 30) == 11 and L.blk_rx_index(0, 31) == 12                                   # QAM128, QAM256
    assert L.blk_rx_index(1, 16) == 7 and L.blk_rx_index(1, 12) == 8 and L.blk_rx_index(1, 26) == 14     # v27p34, v29, v29p78
    assert L.blk_rx_index(0, 32) == -1 and L.blk_rx_index(1, 13) == -1 and L.blk_rx_index(2, 5) == -1    # APSK4, v39, Hamming84
    # config_id = mod*56 + inner*8 + outer + 1 covers 1..616 (python/cognitive_engine.py:87)
    ids = {m * 56 + i * 8 + oo + 1 for m in range(11) for i in range(7) for oo in range(8)}
    assert ids == set(range(1, 617))


@pytest.mark.gpu
def test_flex_tx_block_contract(gpu_required):
    L = _lib()
    h = L.blk_make_flex_tx(1, 0, 0)
    assert h
    assert L.blk_name(h) == b"flex_tx"
    assert [L.blk_sig(h, 0, k) for k in range(2)] == [0, 0] and [L.blk_sig(h, 1, k) for k in range(2)] == [0, 0]
    assert ports(L, h, 0) == ["pdus", "configuration"] and ports(L, h, 1) == ["pdus"]
    props = (C.c_uint * 4)()
    L.blk_tx_props(h, props)
    assert list(props) == [util.PSK4, util.CRC24, 1, 1]
    assert L.blk_work(h, None, 0, 256, None) == -1 and L.blk_error(h) == b"This is not a stream block."
    rng = np.random.default_rng(1)
    pl = rng.integers(0, 256, 256, dtype=np.uint8)
    assert L.blk_post_pdu(h, b"pdus", pl.ctypes.data, 256) == 0
    (port, frame), = pop_all(L, h)
    assert port == "pdus" and len(frame) == 2690
    assert np.array_equal(frame.view(np.uint32), o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl).view(np.uint32))
    # configuration message: optional keys, applied to the next frame
    vals = (C.c_long * 2)(8, 2)
    assert L.blk_post_dict(h, b"configuration", b"modulation,outer_code", vals, 2) == 0
    L.blk_tx_props(h, props)
    assert list(props) == [util.QAM16, util.CRC24, 1, 27]
    L.blk_post_pdu(h, b"pdus", pl.ctypes.data, 256)
    (_, frame2), = pop_all(L, h)
    assert np.array_equal(frame2.view(np.uint32), o.tx_frame(util.QAM16, util.CRC24, 1, 27, pl).view(np.uint32))
    L.blk_destroy(h)
    # out-of-range indices fall back to PSK2 / no FEC
    h = L.blk_make_flex_tx(99, 99, 99)
    L.blk_tx_props(h, props)
    assert list(props) == [util.PSK2, util.CRC24, 1, 1]
    L.blk_destroy(h)


@pytest.mark.gpu
def test_flex_tx_to_flex_rx_loopback_messages(gpu_required):
    L = _lib()
    rng = np.random.default_rng(2)
    rx = L.blk_make_flex_rx()
    assert L.blk_name(rx) == b"flex_rx" and L.blk_output_multiple(rx) == 256
    assert [L.blk_sig(rx, 0, k) for k in range(3)] == [0, 1, 8] and [L.blk_sig(rx, 1, k) for k in range(3)] == [0, 0, 0]
    assert ports(L, rx, 1) == ["constellation", "payload_data", "packet_info"] and ports(L, rx, 0) == []
    cfgs = [(1, 0, 0), (8, 1, 2), (5, 3, 1), (10, 6, 7)]
    sent, frames = [], []
    for (m, i, oo) in cfgs:
        tx = L.blk_make_flex_tx(m, i, oo)
        pl = rng.integers(0, 256, 256, dtype=np.uint8)
        L.blk_post_pdu(tx, b"pdus", pl.ctypes.data, 256)
        (_, fr), = pop_all(L, tx)
        frames.append(fr); sent.append(pl.tobytes())
        L.blk_destroy(tx)
    bad = frames[0].copy(); bad[2 * 90:2 * 250] = 0            # a frame whose header cannot be decoded
    cap = util.build_capture(frames + [bad], rng, [1024] * 5, snr_db=30.0, cfo=0.01)
    cap = np.concatenate([cap, np.zeros((-len(cap)) % 256 + 4096, np.complex64)])
    assert L.blk_work(rx, cap[:100].ctypes.data, 1, 100, None) == -1          # not a multiple of 256
    msgs = []
    for i in range(0, len(cap), 2048):
        chunk = np.ascontiguousarray(cap[i:i + 2048])
        assert L.blk_work(rx, chunk.ctypes.data, 1, len(chunk), None) == len(chunk)
        msgs += pop_all(L, rx)
    names = [m[0] for m in msgs]
    assert names == ["constellation", "payload_data", "packet_info"] * 4 + ["constellation"]
    for k, (m, i, oo) in enumerate(cfgs):
        const, pay, info = msgs[3 * k][1], msgs[3 * k + 1][1], msgs[3 * k + 2][1]
        assert pay == sent[k] and len(const) > 0
        assert info == {"header_valid": 1, "payload_valid": 1, "modulation": m, "inner_code": i, "outer_code": oo}
    assert len(msgs[-1][1]) == 0                                              # header invalid: empty constellation only
    L.blk_destroy(rx)


@pytest.mark.gpu
def test_flex_rx_decode_capture_publishes_what_work_publishes(gpu_required):
    """Additive offline entry of the block: one recorded capture, cut in time over many GPU streams underneath
    (lqb_rx_execute_sharded), publishes the messages that feeding it through work() in 256-multiples publishes."""
    L = _lib()
    rng = np.random.default_rng(21)
    frames = [o.tx_frame(util.MODS[k % 11], util.CRC24, util.INNER[k % 7], util.OUTER[k % 8], rng.integers(0, 256, 60 + 37 * (k % 9), dtype=np.uint8))
              for k in range(22)]
    cap = util.build_capture(frames, rng, [500 + 611 * (k % 5) for k in range(22)], snr_db=28.0, cfo=0.006, tau=-0.15)
    cap = np.concatenate([cap, np.zeros((-len(cap)) % 256 + 2048, np.complex64)])
    rx = L.blk_make_flex_rx()
    ref = []
    for i in range(0, len(cap), 4096):
        chunk = np.ascontiguousarray(cap[i:i + 4096])
        assert L.blk_work(rx, chunk.ctypes.data, 1, len(chunk), None) == len(chunk)
        ref += pop_all(L, rx)
    L.blk_destroy(rx)
    rx = L.blk_make_flex_rx()
    n = L.blk_rx_decode_capture(rx, cap.ctypes.data, len(cap), 24, 8192, 4096)
    got = pop_all(L, rx)
    L.blk_destroy(rx)
    assert n == 22 and [m[0] for m in ref] == ["constellation", "payload_data", "packet_info"] * 22
    assert len(got) == len(ref)
    for a, b in zip(ref, got):
        assert a[0] == b[0]
        if a[0] == "constellation":
            assert np.allclose(a[1], b[1], atol=1e-6)
        else:
            assert a[1] == b[1]


@pytest.mark.gpu
def test_flex_rx_multi_channel_extension(gpu_required):
    L = _lib()
    rng = np.random.default_rng(3)
    rx = L.blk_make_flex_rx_multi(4)
    assert [L.blk_sig(rx, 0, k) for k in range(3)] == [4, 4, 8]
    pls = [rng.integers(0, 256, 100 + c, dtype=np.uint8) for c in range(4)]
    caps = [util.impair(o.tx_frame(util.MODS[c], util.CRC24, 1, 1, pls[c]), rng, snr_db=30, pre=300 + 50 * c, post=900) for c in range(4)]
    n = (max(len(c) for c in caps) + 255) // 256 * 256
    buf = np.zeros((4, n), np.complex64)
    for c in range(4):
        buf[c, :len(caps[c])] = caps[c]
    assert L.blk_work(rx, buf.ctypes.data, 4, n, None) == n
    msgs = pop_all(L, rx)
    assert [m[1] for m in msgs if m[0] == "payload_data"] == [p.tobytes() for p in pls]
    L.blk_destroy(rx)


@pytest.mark.gpu
def test_frame_detector_block_passthrough_and_count(gpu_required):
    L = _lib()
    rng = np.random.default_rng(4)
    det = L.blk_make_frame_detector()
    assert L.blk_name(det) == b"frame_detector_cc"
    assert [L.blk_sig(det, 0, k) for k in range(3)] == [1, 1, 8] and [L.blk_sig(det, 1, k) for k in range(3)] == [1, 1, 8]
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 256, dtype=np.uint8)) for _ in range(6)]
    cap = util.build_capture(frames, rng, [1500] * 6, snr_db=15.0, cfo=0.02)
    out = np.zeros_like(cap)
    for i in range(0, len(cap), 1000):
        chunk = np.ascontiguousarray(cap[i:i + 1000])
        assert L.blk_work(det, chunk.ctypes.data, 1, len(chunk), out[i:i + 1000].ctypes.data) == len(chunk)
    assert np.array_equal(out, cap)
    ref = o.detect_capture(cap, 0.3, 0.45)
    # the block only counts: the streamed work() calls counted what the sequential oracle finds
    assert len(ref) >= 6 and L.blk_det_frames(det) == len(ref)
    # additive offline entry: the same capture in one call, cut in time over GPU streams underneath
    idx = np.zeros(64, np.int64)
    n = L.blk_det_detect_capture(det, cap.ctypes.data, len(cap), idx.ctypes.data, 64, 16, 4096, 2048)
    assert n == len(ref) and L.blk_det_frames(det) == 2 * len(ref)
    assert list(idx[:n]) == [int(np.int64(np.uint64(r["sample_index"]))) for r in ref]
    L.blk_destroy(det)


# ----------------------------------------------------------------------------- liquid-signature veneer
class Stats(C.Structure):
    _fields_ = [("evm", C.c_float), ("rssi", C.c_float), ("cfo", C.c_float), ("framesyms", C.c_void_p),
                ("num_framesyms", C.c_uint), ("mod_scheme", C.c_uint), ("mod_bps", C.c_uint), ("check", C.c_uint),
                ("fec0", C.c_uint), ("fec1", C.c_uint)]


CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.POINTER(C.c_ubyte), C.c_uint, C.c_int, Stats, C.c_void_p)


class Props(C.Structure):
    _fields_ = [("check", C.c_uint), ("fec0", C.c_uint), ("fec1", C.c_uint), ("mod_scheme", C.c_uint)]


@pytest.mark.gpu
def test_liquid_signature_veneer_loopback(gpu_required):
    L = capi.lib()
    L.flexframegen_create.restype = C.c_void_p
    L.flexframegen_create.argtypes = [C.POINTER(Props)]
    L.flexframegen_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    L.flexframegen_getframelen.argtypes = [C.c_void_p]
    L.flexframegen_write_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.flexframegen_setprops.argtypes = [C.c_void_p, C.POINTER(Props)]
    L.flexframegen_destroy.argtypes = [C.c_void_p]
    L.flexframesync_create.restype = C.c_void_p
    L.flexframesync_create.argtypes = [CB, C.c_void_p]
    L.flexframesync_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.flexframesync_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    p = Props()
    L.flexframegenprops_init_default(C.byref(p))
    assert (p.check, p.fec0, p.fec1, p.mod_scheme) == (4, 1, 1, 40)           # CRC-16, none, none, QPSK
    p.check, p.fec0, p.fec1, p.mod_scheme = util.CRC24, 11, 27, util.PSK4
    fg = L.flexframegen_create(C.byref(p))
    hdr = np.zeros(14, np.uint8)
    sent, frames = [], []
    for _ in range(3):
        pl = rng.integers(0, 256, 200, dtype=np.uint8)
        L.flexframegen_assemble(fg, hdr.ctypes.data, pl.ctypes.data, 200)
        n = L.flexframegen_getframelen(fg)
        buf = np.zeros(n, np.complex64)
        assert L.flexframegen_write_samples(fg, buf.ctypes.data, n) == 1
        assert np.array_equal(buf.view(np.uint32), o.tx_frame(util.PSK4, util.CRC24, 11, 27, pl).view(np.uint32))
        frames.append(buf); sent.append(pl.tobytes())
    L.flexframegen_destroy(fg)
    got = []

    def cb(header, hv, payload, plen, pv, stats, ud):
        got.append((hv, pv, bytes(bytearray(payload[:plen])) if hv else b"", stats.num_framesyms, stats.mod_scheme, stats.fec0, stats.fec1))
        return 0
    cbk = CB(cb)
    fs = L.flexframesync_create(cbk, None)
    cap = util.build_capture(frames, rng, [800] * 3, snr_db=20.0)
    cap = np.concatenate([cap, np.zeros(4 * 4096, np.complex64)])
    for i in range(0, len(cap) // 256 * 256, 256):                           # the reference's 256-sample chunking
        chunk = np.ascontiguousarray(cap[i:i + 256])
        before = len(got)
        L.flexframesync_execute(fs, chunk.ctypes.data, 256)
        assert len(got) - before <= 1                                        # at most one callback per call
    L.flexframesync_destroy(fs)
    assert [g[2] for g in got] == sent and all(g[0] == 1 and g[1] == 1 for g in got)
    assert all(g[4] == util.PSK4 and g[5] == 11 and g[6] == 27 for g in got)
    # one large buffer in one call: every completed frame is delivered before execute returns; the tail that does
    # not fill a batch is processed by flexframesync_flush (no trailing zero padding this time)
    L.flexframesync_flush.argtypes = [C.c_void_p]
    got.clear()
    fs = L.flexframesync_create(cbk, None)
    cap = np.ascontiguousarray(util.build_capture(frames, rng, [800] * 3, snr_db=20.0))
    cut = (len(cap) - 3000) // 256 * 256                 # the third frame ends 1700 samples before the end: inside the tail
    L.flexframesync_execute(fs, cap.ctypes.data, cut)
    assert len(got) == 2
    tail = np.ascontiguousarray(cap[cut:])               # fewer samples than one batch: they wait
    L.flexframesync_execute(fs, tail.ctypes.data, len(tail))
    assert len(got) == 2
    L.flexframesync_flush(fs)
    assert len(got) == 3
    L.flexframesync_destroy(fs)
    assert [g[2] for g in got] == sent


@pytest.mark.gpu
def test_liquid_signature_detector_veneer(gpu_required):
    L = capi.lib()
    L.msequence_create.restype = C.c_void_p
    L.msequence_create.argtypes = [C.c_uint] * 3
    L.msequence_advance.argtypes = [C.c_void_p]
    L.msequence_destroy.argtypes = [C.c_void_p]
    L.qdetector_cccf_create_linear.restype = C.c_void_p
    L.qdetector_cccf_create_linear.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_float]
    L.qdetector_cccf_set_threshold.argtypes = [C.c_void_p, C.c_float]
    L.qdetector_cccf_execute.restype = C.c_void_p
    L.qdetector_cccf_execute.argtypes = [C.c_void_p, C.c_float, C.c_float]     # float complex by value == two floats in xmm0
    L.qdetector_cccf_get_tau.restype = C.c_float
    L.qdetector_cccf_get_tau.argtypes = [C.c_void_p]
    L.qdetector_cccf_destroy.argtypes = [C.c_void_p]
    # the reference builds its preamble exactly like this (lib/frame_detector_cc_impl.cc:46-52)
    ms = L.msequence_create(7, 0x0089, 1)
    pn = np.zeros(64, np.complex64)
    for i in range(64):
        re = np.sqrt(0.5) if L.msequence_advance(ms) else -np.sqrt(0.5)
        im = np.sqrt(0.5) if L.msequence_advance(ms) else -np.sqrt(0.5)
        pn[i] = re + 1j * im
    L.msequence_destroy(ms)
    q = L.qdetector_cccf_create_linear(pn.ctypes.data, 64, 9, 2, 7, 0.3)
    assert q
    L.qdetector_cccf_set_threshold(q, 0.45)
    assert not L.qdetector_cccf_create_linear(pn.ctypes.data, 63, 9, 2, 7, 0.3)   # anything else is refused, not guessed
