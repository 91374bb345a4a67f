"""ctypes loader for the CPU ORACLE (test infrastructure only -- see oracle/lqo.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblqo.so")


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


class FrameRecord(C.Structure):
    _fields_ = [("sample_index", C.c_uint64), ("header_valid", C.c_int), ("payload_valid", C.c_int),
                ("payload_len", C.c_uint), ("num_framesyms", C.c_uint),
                ("mod_scheme", C.c_uint), ("mod_bps", C.c_uint), ("check", C.c_uint),
                ("fec0", C.c_uint), ("fec1", C.c_uint),
                ("evm", C.c_float), ("rssi", C.c_float), ("cfo", C.c_float), ("tau_hat", C.c_float),
                ("gamma_hat", C.c_float), ("dphi_hat", C.c_float), ("phi_hat", C.c_float), ("rxy", C.c_float),
                ("header", C.c_uint8 * 20), ("payload_off", C.c_uint64), ("syms_off", C.c_uint64)]


class Detection(C.Structure):
    _fields_ = [("sample_index", C.c_uint64), ("tau_hat", C.c_float), ("gamma_hat", C.c_float),
                ("dphi_hat", C.c_float), ("phi_hat", C.c_float), ("rxy", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, f32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.c_void_p
        L.lqo_crc_key.restype = C.c_uint
        L.lqo_crc_key.argtypes = [C.c_int, vp, C.c_uint]
        L.lqo_fec_enc_len.restype = C.c_uint
        L.lqo_fec_enc_len.argtypes = [C.c_int, C.c_uint]
        L.lqo_fec_encode.argtypes = [C.c_int, C.c_uint, vp, vp]
        L.lqo_fec_decode.argtypes = [C.c_int, C.c_uint, vp, vp]
        L.lqo_interleave.argtypes = [vp, C.c_uint, C.c_int]
        L.lqo_deinterleave.argtypes = [vp, C.c_uint, C.c_int]
        L.lqo_scramble.argtypes = [vp, C.c_uint]
        L.lqo_packetizer_enc_len.restype = C.c_uint
        L.lqo_packetizer_enc_len.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int]
        L.lqo_packetizer_encode.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, vp, vp]
        L.lqo_packetizer_decode.restype = C.c_int
        L.lqo_packetizer_decode.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, vp, vp]
        L.lqo_rs_decode_block.restype = C.c_int
        L.lqo_rs_decode_block.argtypes = [vp, C.c_uint]
        L.lqo_rs_encode_block.argtypes = [vp, C.c_uint, vp]
        L.lqo_qpm_frame_len.restype = C.c_uint
        L.lqo_qpm_frame_len.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
        L.lqo_qpm_encode.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.lqo_qpm_decode.restype = C.c_int
        L.lqo_qpm_decode.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.lqo_firdes_arkaiser.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, vp]
        L.lqo_interp_taps.argtypes = [C.c_uint, C.c_uint, C.c_float, vp]
        L.lqo_pfb_rnyquist.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_float, vp]
        L.lqo_fft.argtypes = [vp, vp, C.c_uint, C.c_int]
        L.lqo_nco_sintab.restype = f32p
        L.lqo_nco_constrain.restype = C.c_uint32
        L.lqo_nco_constrain.argtypes = [C.c_float]
        L.lqo_tx_frame.restype = C.c_uint
        L.lqo_tx_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_uint, vp, C.c_uint]
        L.lqo_rx_capture.restype = C.c_uint
        L.lqo_rx_capture.argtypes = [vp, C.c_uint64, C.c_uint, vp, C.c_uint, vp, C.c_uint64, vp, C.c_uint64]
        L.lqo_rx_capture_soft.restype = C.c_uint
        L.lqo_rx_capture_soft.argtypes = L.lqo_rx_capture.argtypes
        L.lqo_rx_many.restype = C.c_uint64
        L.lqo_rx_many.argtypes = [vp, C.c_uint, C.c_uint64, C.c_uint64, C.c_uint, C.POINTER(C.c_uint64)]
        L.lqo_detect_capture.restype = C.c_uint
        L.lqo_detect_capture.argtypes = [vp, C.c_uint64, C.c_float, C.c_float, vp, C.c_uint]
        L.lqo_qpilotgen.argtypes = [C.c_uint, C.c_uint, vp, vp]
        L.lqo_qpilotsync.argtypes = [C.c_uint, C.c_uint, vp, vp, f32p, f32p, f32p]
        L.lqo_mseq_init.argtypes = [vp, C.c_uint, C.c_uint, C.c_uint]
        L.lqo_mseq_advance.restype = C.c_uint
        L.lqo_mseq_advance.argtypes = [vp]
        L.lqo_modem_init.argtypes = [vp, C.c_int]
        L.lqo_modem_bps.restype = C.c_uint
        L.lqo_modem_bps.argtypes = [C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def tx_frame(ms, check, fec0, fec1, payload, header=None):
    """flexframegen assemble + write_samples -> complex64 array."""
    L = lib()
    payload = np.ascontiguousarray(payload, dtype=np.uint8)
    hdr = np.zeros(14, np.uint8) if header is None else np.ascontiguousarray(header, dtype=np.uint8)
    n = L.lqo_tx_frame(ms, check, fec0, fec1, _ptr(hdr), _ptr(payload), len(payload), None, 0)
    out = np.zeros(n, np.complex64)
    L.lqo_tx_frame(ms, check, fec0, fec1, _ptr(hdr), _ptr(payload), len(payload), _ptr(out), n)
    return out


def rx_capture(x, chunk=256, max_frames=4096, soft=False):
    """Run one flexframesync over capture x; returns list of dict frames.  soft=True: soft-decision payloads (opt-in
    extension: soft demodulation + soft-input Viterbi where the stage nearest the channel is convolutional)."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.complex64)
    recs = (FrameRecord * max_frames)()
    pcap = max(1 << 16, 2 * len(x))
    scap = max(1 << 16, len(x))
    ppool = np.zeros(pcap, np.uint8)
    spool = np.zeros(scap, np.complex64)
    fn = L.lqo_rx_capture_soft if soft else L.lqo_rx_capture
    n = fn(_ptr(x), len(x), chunk, C.byref(recs), max_frames, _ptr(ppool), pcap, _ptr(spool), scap)
    out = []
    for i in range(min(n, max_frames)):
        r = recs[i]
        d = {k: getattr(r, k) for k, _ in FrameRecord._fields_ if k not in ("header", "payload_off", "syms_off")}
        d["sample_index"] = int(np.int64(np.uint64(r.sample_index)))
        d["header"] = bytes(r.header)
        d["payload"] = bytes(ppool[r.payload_off:r.payload_off + r.payload_len]) if r.header_valid else b""
        d["framesyms"] = spool[r.syms_off:r.syms_off + r.num_framesyms].copy()
        out.append(d)
    return out


def rx_many(x2d, n_threads):
    """x2d: [n_streams, n] complex64. Returns (frames, valid_payloads)."""
    L = lib()
    x2d = np.ascontiguousarray(x2d, dtype=np.complex64)
    nv = C.c_uint64(0)
    f = L.lqo_rx_many(_ptr(x2d), x2d.shape[0], x2d.shape[1], x2d.shape[1], n_threads, C.byref(nv))
    return int(f), int(nv.value)


def detect_capture(x, beta=0.3, threshold=0.45, max_out=65536):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = (Detection * max_out)()
    n = L.lqo_detect_capture(_ptr(x), len(x), beta, threshold, C.byref(out), max_out)
    return [{k: getattr(out[i], k) for k, _ in Detection._fields_} for i in range(min(n, max_out))]
