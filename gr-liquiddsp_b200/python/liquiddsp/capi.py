"""ctypes binding of include/lqb200.h (liblqb200.so).

This is the Python-side stand-in for the SWIG module `liquiddsp_swig` of the reference
(/root/reference/swig/liquiddsp_swig.i:19-26): it binds exactly the C-ABI the C++ blocks call.
It never imports anything from oracle/ and has no CPU fallback -- if the CUDA library cannot
be loaded, importing this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQB_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "liblqb200.so"))

MEM_HOST, MEM_DEVICE, MEM_HOST_SC16, MEM_DEVICE_SC16 = 0, 1, 2, 3
RX_NO_FRAMESYMS, RX_DEVICE_RESULTS, RX_SOFT = 1, 2, 4

DECLARED_SYMBOLS = [
    "lqb_last_error", "lqb_device_count", "lqb_version",
    "lqb_rx_create", "lqb_rx_destroy", "lqb_rx_reset", "lqb_rx_execute", "lqb_rx_execute_dense",
    "lqb_rx_submit", "lqb_rx_submit_dense", "lqb_rx_collect",
    "lqb_rx_poll", "lqb_rx_counts", "lqb_rx_last_timing", "lqb_rx_launch_count", "lqb_rx_last_work", "lqb_rx_last_search_bins", "lqb_rx_lane_count",
    "lqb_tx_create", "lqb_tx_destroy", "lqb_tx_props_init_default", "lqb_tx_frame_len", "lqb_tx_assemble", "lqb_tx_submit", "lqb_tx_collect", "lqb_tx_last_timing",
    "lqb_det_create", "lqb_det_destroy", "lqb_det_reset", "lqb_det_execute", "lqb_det_execute_dense",
    "lqb_det_poll", "lqb_det_last_timing", "lqb_det_last_work", "lqb_det_last_search",
    "lqb_det_execute_sharded", "lqb_det_last_shard_info", "lqb_rx_execute_sharded", "lqb_rx_last_shard_info",
    "lqb_tab_interp_taps", "lqb_tab_pfb_banks", "lqb_tab_detector_template", "lqb_tab_nco_sintab",
    "lqb_tab_packet_len", "lqb_tab_ilv_bit_perm", "lqb_tab_secded_columns",
]


class RxOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("n_streams", C.c_uint32), ("max_frame_samples", C.c_uint32),
                ("flags", C.c_uint32), ("cuda_stream", C.c_void_p), ("n_lanes", C.c_uint32)]


class FrameResult(C.Structure):
    _fields_ = [("stream", C.c_uint32), ("seq", C.c_uint32), ("sample_index", C.c_int64),
                ("header", C.c_uint8 * 20), ("header_valid", C.c_int32), ("payload_valid", C.c_int32),
                ("payload_len", C.c_uint32), ("payload", C.c_void_p), ("framesyms", C.c_void_p),
                ("num_framesyms", C.c_uint32),
                ("mod_scheme", C.c_uint32), ("mod_bps", C.c_uint32), ("check", C.c_uint32),
                ("fec0", C.c_uint32), ("fec1", C.c_uint32),
                ("evm", C.c_float), ("rssi", C.c_float), ("cfo", C.c_float),
                ("tau_hat", C.c_float), ("gamma_hat", C.c_float), ("dphi_hat", C.c_float),
                ("phi_hat", C.c_float), ("rxy", C.c_float), ("flags", C.c_uint32)]


# numpy view of lqb_frame_result (same field order; pointers as addresses)
FRAME_DTYPE = np.dtype([("stream", "<u4"), ("seq", "<u4"), ("sample_index", "<i8"), ("header", "u1", (20,)),
                        ("header_valid", "<i4"), ("payload_valid", "<i4"), ("payload_len", "<u4"),
                        ("payload", "<u8"), ("framesyms", "<u8"), ("num_framesyms", "<u4"),
                        ("mod_scheme", "<u4"), ("mod_bps", "<u4"), ("check", "<u4"), ("fec0", "<u4"), ("fec1", "<u4"),
                        ("evm", "<f4"), ("rssi", "<f4"), ("cfo", "<f4"), ("tau_hat", "<f4"), ("gamma_hat", "<f4"),
                        ("dphi_hat", "<f4"), ("phi_hat", "<f4"), ("rxy", "<f4"), ("flags", "<u4")], align=True)
assert FRAME_DTYPE.itemsize == C.sizeof(FrameResult) and all(
    FRAME_DTYPE.fields[k][1] == getattr(FrameResult, k).offset for k, _ in FrameResult._fields_)


class TxOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("flags", C.c_uint32), ("cuda_stream", C.c_void_p)]


class TxProps(C.Structure):
    _fields_ = [("check", C.c_uint32), ("fec0", C.c_uint32), ("fec1", C.c_uint32), ("mod_scheme", C.c_uint32)]


class DetOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("n_streams", C.c_uint32), ("beta", C.c_float), ("threshold", C.c_float),
                ("dphi_max", C.c_float), ("cuda_stream", C.c_void_p)]


class DetectionResult(C.Structure):
    _fields_ = [("stream", C.c_uint32), ("seq", C.c_uint32), ("sample_index", C.c_int64),
                ("tau_hat", C.c_float), ("gamma_hat", C.c_float), ("dphi_hat", C.c_float),
                ("phi_hat", C.c_float), ("rxy", C.c_float)]


_lib = None


def lib():
    """Load liblqb200.so (raises OSError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(LIB_PATH)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.lqb_last_error.restype = C.c_char_p
    L.lqb_rx_create.restype = vp
    L.lqb_rx_create.argtypes = [C.POINTER(RxOpts)]
    L.lqb_rx_destroy.argtypes = [vp]
    L.lqb_rx_reset.argtypes = [vp, C.c_int]
    L.lqb_rx_execute.argtypes = [vp, u32, vp, vp, vp, C.c_int]
    L.lqb_rx_execute_dense.argtypes = [vp, vp, u64, u64, C.c_int]
    L.lqb_rx_submit.argtypes = [vp, u32, vp, vp, vp, C.c_int]
    L.lqb_rx_submit_dense.argtypes = [vp, vp, u64, u64, C.c_int]
    L.lqb_rx_collect.argtypes = [vp]
    L.lqb_rx_poll.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.lqb_rx_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.lqb_rx_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.lqb_rx_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.lqb_rx_last_work.argtypes = [vp, C.POINTER(u64)]
    L.lqb_rx_lane_count.argtypes = [vp]
    L.lqb_rx_last_search_bins.argtypes = [vp, C.POINTER(u64)]
    L.lqb_det_last_work.argtypes = [vp, C.POINTER(u64)]
    L.lqb_det_last_search.argtypes = [vp, C.POINTER(u64)]
    if hasattr(L, "lqb_tx_create"):
        L.lqb_tx_create.restype = vp
        L.lqb_tx_create.argtypes = [C.POINTER(TxOpts)]
        L.lqb_tx_destroy.argtypes = [vp]
        L.lqb_tx_props_init_default.argtypes = [C.POINTER(TxProps)]
        L.lqb_tx_frame_len.argtypes = [C.POINTER(TxProps), u32, C.POINTER(u32)]
        L.lqb_tx_assemble.argtypes = [vp, u32, vp, vp, vp, vp, vp, C.c_int]
        L.lqb_tx_submit.argtypes = [vp, u32, vp, vp, vp, vp, vp, C.c_int]
        L.lqb_tx_collect.argtypes = [vp]
        L.lqb_tx_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.lqb_det_create.restype = vp
    L.lqb_det_create.argtypes = [C.POINTER(DetOpts)]
    L.lqb_det_destroy.argtypes = [vp]
    L.lqb_det_reset.argtypes = [vp, C.c_int]
    L.lqb_det_execute.argtypes = [vp, u32, vp, vp, vp, C.c_int]
    L.lqb_det_execute_dense.argtypes = [vp, vp, u64, u64, C.c_int]
    L.lqb_det_poll.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.lqb_det_execute_sharded.argtypes = [vp, vp, u64, C.c_int, u32, u32]
    L.lqb_rx_execute_sharded.argtypes = [vp, vp, u64, C.c_int, u32, u32]
    L.lqb_rx_last_shard_info.argtypes = [vp, C.POINTER(u64)]
    L.lqb_det_last_shard_info.argtypes = [vp, C.POINTER(u64)]
    L.lqb_det_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.lqb_tab_interp_taps.argtypes = [C.c_float, vp]
    L.lqb_tab_pfb_banks.argtypes = [C.c_float, vp]
    L.lqb_tab_detector_template.argtypes = [C.c_float, vp]
    L.lqb_tab_nco_sintab.argtypes = [vp]
    L.lqb_tab_packet_len.argtypes = [u32, u32, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    _lib = L
    return L


class LqbError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise LqbError("lqb error %d: %s" % (rc, lib().lqb_last_error().decode()))


def _frame_to_dict(r, host_results=True):
    d = {k: getattr(r, k) for k, _ in FrameResult._fields_ if k not in ("header", "payload", "framesyms")}
    d["header"] = bytes(r.header)
    if host_results and r.header_valid and r.payload:
        d["payload"] = C.string_at(r.payload, r.payload_len)
    else:
        d["payload"] = b""
    if host_results and r.framesyms and r.num_framesyms:
        buf = (C.c_float * (2 * r.num_framesyms)).from_address(r.framesyms)
        d["framesyms"] = np.frombuffer(buf, dtype=np.complex64).copy()
    else:
        d["framesyms"] = np.zeros(0, np.complex64)
    return d


class Rx:
    """Batch flexframesync: n_streams independent channels on one GPU."""

    def __init__(self, n_streams=1, device=0, max_frame_samples=0, flags=0, cuda_stream=None, lanes=0):
        o = RxOpts(device, n_streams, max_frame_samples, flags, cuda_stream, lanes)
        self._L = lib()
        self._h = self._L.lqb_rx_create(C.byref(o))
        if not self._h:
            raise LqbError("lqb_rx_create failed: " + self._L.lqb_last_error().decode())
        self.n_streams = n_streams
        self.flags = flags

    def close(self):
        if getattr(self, "_h", None):
            self._L.lqb_rx_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=-1):
        _check(self._L.lqb_rx_reset(self._h, stream))

    def execute(self, chunks, stream_ids=None):
        """chunks: list of complex64 numpy arrays (host memory), one per listed stream."""
        n = len(chunks)
        arrs = [np.ascontiguousarray(c, dtype=np.complex64) for c in chunks]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[len(a) for a in arrs])
        ids = None if stream_ids is None else (C.c_uint32 * n)(*stream_ids)
        _check(self._L.lqb_rx_execute(self._h, n, ids, ptrs, lens, MEM_HOST))

    def execute_dense_ptr(self, ptr, stride, n_samples, mem):
        _check(self._L.lqb_rx_execute_dense(self._h, C.c_void_p(ptr), stride, n_samples, mem))

    def execute_sharded(self, x, seg_len=1 << 20, preroll=1 << 16):
        """One complex64 capture (host numpy array) decoded as ONE sequential receiver would decode it, cut in time over
        the handle's streams; the frames do not depend on seg_len / preroll (see lqb_rx_execute_sharded)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        _check(self._L.lqb_rx_execute_sharded(self._h, C.c_void_p(x.ctypes.data), len(x), MEM_HOST, seg_len, preroll))

    def execute_sharded_ptr(self, ptr, n_samples, mem, seg_len=1 << 20, preroll=1 << 16):
        _check(self._L.lqb_rx_execute_sharded(self._h, C.c_void_p(ptr), n_samples, mem, seg_len, preroll))

    def shard_info(self):
        w = (C.c_uint64 * 4)()
        _check(self._L.lqb_rx_last_shard_info(self._h, w))
        return dict(segments=int(w[0]), runs=int(w[1]), rounds=int(w[2]), executes=int(w[3]))

    def execute_sc16(self, chunks, stream_ids=None):
        """chunks: list of int16 arrays of shape [n, 2] (re, im); a sample is value / 32768 (LQB_MEM_HOST_SC16)."""
        n = len(chunks)
        arrs = [np.ascontiguousarray(c, dtype=np.int16).reshape(-1, 2) for c in chunks]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[len(a) for a in arrs])
        ids = None if stream_ids is None else (C.c_uint32 * n)(*stream_ids)
        _check(self._L.lqb_rx_execute(self._h, n, ids, ptrs, lens, MEM_HOST_SC16))

    # pipelined form: submit() returns when the search is done, collect() makes the oldest submitted call current
    def submit(self, chunks, stream_ids=None):
        n = len(chunks)
        arrs = [np.ascontiguousarray(c, dtype=np.complex64) for c in chunks]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[len(a) for a in arrs])
        ids = None if stream_ids is None else (C.c_uint32 * n)(*stream_ids)
        _check(self._L.lqb_rx_submit(self._h, n, ids, ptrs, lens, MEM_HOST))

    def submit_dense_ptr(self, ptr, stride, n_samples, mem):
        _check(self._L.lqb_rx_submit_dense(self._h, C.c_void_p(ptr), stride, n_samples, mem))

    def collect(self):
        _check(self._L.lqb_rx_collect(self._h))

    def execute_dense(self, x2d):
        x2d = np.ascontiguousarray(x2d, dtype=np.complex64)
        assert x2d.shape[0] == self.n_streams
        self.execute_dense_ptr(x2d.ctypes.data, x2d.shape[1], x2d.shape[1], MEM_HOST)

    def counts(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.lqb_rx_counts(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def poll(self, raw=False):
        nf, _ = self.counts()
        arr = (FrameResult * max(nf, 1))()
        n = C.c_uint32(0)
        _check(self._L.lqb_rx_poll(self._h, C.byref(arr), nf, C.byref(n)))
        if raw:
            return arr, nf
        host = not (self.flags & RX_DEVICE_RESULTS)
        return [_frame_to_dict(arr[i], host) for i in range(nf)]

    def poll_array(self):
        """All frame records of the last execute / collect as one numpy structured array (fields of lqb_frame_result;
        `payload` and `framesyms` are raw addresses valid until the next submit / execute).  No per-frame Python objects:
        this is the form for callers that handle tens of thousands of frames per call."""
        nf, _ = self.counts()
        arr = (FrameResult * max(nf, 1))()
        n = C.c_uint32(0)
        _check(self._L.lqb_rx_poll(self._h, C.byref(arr), nf, C.byref(n)))
        return np.frombuffer(arr, dtype=FRAME_DTYPE)[:nf]          # (the view keeps the ctypes array alive)

    def timing(self):
        ms = (C.c_float * 6)()
        _check(self._L.lqb_rx_last_timing(self._h, ms))
        return list(ms)

    def lanes(self):
        return int(self._L.lqb_rx_lane_count(self._h))

    def launches(self):
        v = C.c_uint64(0)
        _check(self._L.lqb_rx_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def work(self):
        w = (C.c_uint64 * 6)()
        _check(self._L.lqb_rx_last_work(self._h, w))
        b = C.c_uint64(0)
        _check(self._L.lqb_rx_last_search_bins(self._h, C.byref(b)))
        return dict(windows=int(w[0]), aligns=int(w[1]), symbols=int(w[2]), samples=int(w[3]),
                    exact_windows=int(w[4]), coarse_tiles=int(w[5]), exact_bins=int(b.value))


class Det:
    """Batch qdetector_cccf with frame_detector_cc's parameters."""

    def __init__(self, n_streams=1, device=0, beta=0.0, threshold=0.0, dphi_max=0.0, cuda_stream=None):
        o = DetOpts(device, n_streams, beta, threshold, dphi_max, cuda_stream)
        self._L = lib()
        self._h = self._L.lqb_det_create(C.byref(o))
        if not self._h:
            raise LqbError("lqb_det_create failed: " + self._L.lqb_last_error().decode())
        self.n_streams = n_streams

    def close(self):
        if getattr(self, "_h", None):
            self._L.lqb_det_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=-1):
        _check(self._L.lqb_det_reset(self._h, stream))

    def execute(self, chunks, stream_ids=None):
        n = len(chunks)
        arrs = [np.ascontiguousarray(c, dtype=np.complex64) for c in chunks]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[len(a) for a in arrs])
        ids = None if stream_ids is None else (C.c_uint32 * n)(*stream_ids)
        _check(self._L.lqb_det_execute(self._h, n, ids, ptrs, lens, MEM_HOST))

    def execute_dense_ptr(self, ptr, stride, n_samples, mem):
        _check(self._L.lqb_det_execute_dense(self._h, C.c_void_p(ptr), stride, n_samples, mem))

    def execute_sharded(self, x, seg_len=1 << 18, preroll=1 << 14):
        """One complex64 capture (host numpy array) searched as the sequential detector would, cut in time over the
        handle's streams; the detection list does not depend on seg_len / preroll (see lqb_det_execute_sharded)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        _check(self._L.lqb_det_execute_sharded(self._h, C.c_void_p(x.ctypes.data), len(x), MEM_HOST, seg_len, preroll))

    def execute_sharded_ptr(self, ptr, n_samples, mem, seg_len=1 << 18, preroll=1 << 14):
        _check(self._L.lqb_det_execute_sharded(self._h, C.c_void_p(ptr), n_samples, mem, seg_len, preroll))

    def shard_info(self):
        w = (C.c_uint64 * 4)()
        _check(self._L.lqb_det_last_shard_info(self._h, w))
        return dict(segments=int(w[0]), runs=int(w[1]), rounds=int(w[2]), launches=int(w[3]))

    def poll(self):
        n = C.c_uint32(0)
        _check(self._L.lqb_det_poll(self._h, None, 0, C.byref(n)))
        arr = (DetectionResult * max(n.value, 1))()
        _check(self._L.lqb_det_poll(self._h, C.byref(arr), n.value, C.byref(n)))
        return [{k: getattr(arr[i], k) for k, _ in DetectionResult._fields_} for i in range(n.value)]

    def timing(self):
        ms = C.c_float(0)
        _check(self._L.lqb_det_last_timing(self._h, C.byref(ms)))
        return float(ms.value)

    def windows(self):
        w = C.c_uint64(0)
        _check(self._L.lqb_det_last_work(self._h, C.byref(w)))
        return int(w.value)

    def search(self):
        w = (C.c_uint64 * 4)()
        _check(self._L.lqb_det_last_search(self._h, w))
        return dict(windows=int(w[0]), aligns=int(w[1]), exact_windows=int(w[2]), exact_bins=int(w[3]))


def tab_interp_taps(beta):
    h = np.zeros(30, np.float32)
    lib().lqb_tab_interp_taps(beta, h.ctypes.data)
    return h


def tab_pfb_banks(beta):
    b = np.zeros((32, 28), np.float32)
    lib().lqb_tab_pfb_banks(beta, b.ctypes.data)
    return b


def tab_detector_template(beta):
    s = np.zeros(156, np.complex64)
    lib().lqb_tab_detector_template(beta, s.ctypes.data)
    return s


def tab_nco_sintab():
    t = np.zeros(1024, np.float32)
    lib().lqb_tab_nco_sintab(t.ctypes.data)
    return t


def tab_secded_columns(data_bytes):
    c = np.zeros(8 * data_bytes, np.uint8)
    lib().lqb_tab_secded_columns.argtypes = [C.c_uint32, C.c_void_p]
    _check(lib().lqb_tab_secded_columns(data_bytes, c.ctypes.data))
    return c


def tab_ilv_bit_perm(n):
    p = np.zeros(8 * n, np.uint32)
    lib().lqb_tab_ilv_bit_perm.argtypes = [C.c_uint32, C.c_void_p]
    _check(lib().lqb_tab_ilv_bit_perm(n, p.ctypes.data))
    return p


def tab_packet_len(n, check, fec0, fec1, ms):
    a, b = C.c_uint32(0), C.c_uint32(0)
    _check(lib().lqb_tab_packet_len(n, check, fec0, fec1, ms, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


class Tx:
    """Batch flexframegen: assemble + write many frames per launch."""

    def __init__(self, device=0, cuda_stream=None):
        o = TxOpts(device, 0, cuda_stream)
        self._L = lib()
        self._h = self._L.lqb_tx_create(C.byref(o))
        if not self._h:
            raise LqbError("lqb_tx_create failed: " + self._L.lqb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.lqb_tx_destroy(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def frame_len(mod_scheme, check, fec0, fec1, payload_len):
        p = TxProps(check, fec0, fec1, mod_scheme)
        n = C.c_uint32(0)
        _check(lib().lqb_tx_frame_len(C.byref(p), payload_len, C.byref(n)))
        return int(n.value)

    def assemble(self, props, payloads, headers=None):
        """props: list of (mod_scheme, check, fec0, fec1); payloads: list of uint8 arrays.
        Returns a list of complex64 arrays (host)."""
        n = len(payloads)
        P = (TxProps * n)(*[TxProps(c, f0, f1, ms) for (ms, c, f0, f1) in props])
        pls = [np.ascontiguousarray(p, dtype=np.uint8) for p in payloads]
        lens = (C.c_uint32 * n)(*[len(p) for p in pls])
        pp = (C.c_void_p * n)(*[p.ctypes.data if len(p) else None for p in pls])
        hp = None
        if headers is not None:
            hs = [np.ascontiguousarray(h, dtype=np.uint8) for h in headers]
            hp = (C.c_void_p * n)(*[h.ctypes.data for h in hs])
        outs = [np.zeros(self.frame_len(ms, c, f0, f1, len(p)), np.complex64) for (ms, c, f0, f1), p in zip(props, pls)]
        op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        _check(self._L.lqb_tx_assemble(self._h, n, P, hp, pp, lens, op, MEM_HOST))
        return outs

    def kernel_ms(self):
        v = C.c_float(0)
        _check(self._L.lqb_tx_last_timing(self._h, C.byref(v)))
        return float(v.value)

    def collect(self):
        _check(self._L.lqb_tx_collect(self._h))

    def assemble_device_arrays(self, props4, payload_ptrs, payload_lens, out_ptrs, wait=True):
        """Device-resident variant for large batches: numpy arrays in (no per-frame Python objects).
        props4: uint32 [n, 4] rows (check, fec0, fec1, mod_scheme) = lqb_tx_props; payload_ptrs / out_ptrs: uint64 [n]
        device addresses; payload_lens: uint32 [n].  Frames are written in place at out_ptrs."""
        props4 = np.ascontiguousarray(props4, dtype=np.uint32)
        pp = np.ascontiguousarray(payload_ptrs, dtype=np.uint64)
        op = np.ascontiguousarray(out_ptrs, dtype=np.uint64)
        ln = np.ascontiguousarray(payload_lens, dtype=np.uint32)
        n = len(pp)
        assert props4.shape == (n, 4) and len(op) == n and len(ln) == n
        fn = self._L.lqb_tx_assemble if wait else self._L.lqb_tx_submit      # wait=False: complete in stream order (lqb_tx_submit)
        _check(fn(self._h, n, props4.ctypes.data, None, pp.ctypes.data, ln.ctypes.data, op.ctypes.data, MEM_DEVICE))

    def assemble_device(self, props, payload_ptrs, payload_lens, out_ptrs, header_ptrs=None):
        """Device-resident variant: raw device pointers in, frames written to out_ptrs."""
        n = len(payload_ptrs)
        P = (TxProps * n)(*[TxProps(c, f0, f1, ms) for (ms, c, f0, f1) in props])
        lens = (C.c_uint32 * n)(*payload_lens)
        pp = (C.c_void_p * n)(*payload_ptrs)
        op = (C.c_void_p * n)(*out_ptrs)
        hp = None if header_ptrs is None else (C.c_void_p * n)(*header_ptrs)
        _check(self._L.lqb_tx_assemble(self._h, n, P, hp, pp, lens, op, MEM_DEVICE))
