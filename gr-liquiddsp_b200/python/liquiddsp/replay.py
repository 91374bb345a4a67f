"""Capture replay: the step either side of the receiver in the authors' loopback -- a GNU Radio
file_source of gr_complex items (`python/.idea/workspace.xml:79-86` in the reference) feeding flex_rx, and PDUs
coming out (`lib/flex_rx_impl.cc:218-247`).

A capture file is raw interleaved float32 I/Q (GNU Radio's complex64 file_sink format).  One file is one channel; several
files (or one file de-interleaved channel-major with `n_channels`) are replayed as independent streams of one batch
receiver.  Chunks are multiples of 256 samples, as `flex_rx` feeds them (`set_output_multiple(256)`,
lib/flex_rx_impl.cc:50); the receiver carries partial frames across chunks, so the chunk size never changes a result."""
import numpy as np

from . import capi


def open_capture(path, n_channels=1):
    """Memory-map a complex64 capture; returns an array [n_channels, n_samples] (channel-major when n_channels > 1)."""
    x = np.memmap(path, dtype=np.complex64, mode="r")
    n = (len(x) // n_channels) * n_channels
    return x[:n].reshape(n_channels, n // n_channels)


def chunk_bounds(n_samples, chunk):
    """[(start, stop)] covering [0, n_samples) in chunks of `chunk` samples (a multiple of 256); the tail may be short."""
    if chunk <= 0 or chunk % 256:
        raise ValueError("chunk must be a positive multiple of 256 samples")
    return [(s, min(s + chunk, n_samples)) for s in range(0, n_samples, chunk)]


def replay(captures, chunk=1 << 18, device=0, flags=0, rx=None):
    """Feed `captures` (list of 1-D complex64 arrays / memmaps, or a 2-D array, one row per channel) through one batch
    receiver chunk by chunk and yield every frame (dict, as `capi.Rx.poll`) in (call, stream, time) order.
    Uses the pipelined submit / collect form: chunk k+1 travels and is searched while chunk k's payloads decode."""
    rows = [np.asarray(c).reshape(-1) for c in captures]
    own = rx is None
    if own:
        rx = capi.Rx(len(rows), device=device, flags=flags)
    n_max = max((len(r) for r in rows), default=0)
    pending = 0
    try:
        for s, e in chunk_bounds(n_max, chunk):
            ids = [i for i, r in enumerate(rows) if s < len(r)]
            parts = [np.ascontiguousarray(rows[i][s:min(e, len(rows[i]))], dtype=np.complex64) for i in ids]
            rx.submit(parts, ids)
            pending += 1
            if pending == 2:
                rx.collect()
                pending -= 1
                yield from rx.poll()
        while pending:
            rx.collect()
            pending -= 1
            yield from rx.poll()
    finally:
        if own:
            rx.close()


def replay_sharded(capture, workers=512, seg_len=1 << 20, preroll=1 << 16, device=0, flags=0, rx=None):
    """ONE capture (a single channel: what one flex_rx block instance sees) decoded at batch speed: the capture is cut in
    time over `workers` streams of a batch receiver inside the library (lqb_rx_execute_sharded) and the frames that come
    back are those of one sequential receiver, whatever the cut.  Returns the list of frames (dicts, as `capi.Rx.poll`,
    stream 0, in time order).  `preroll` should span the longest frame plus a gap; seg_len * workers samples are searched
    per pass."""
    x = np.ascontiguousarray(np.asarray(capture).reshape(-1), dtype=np.complex64)
    own = rx is None
    if own:
        n_seg = max(1, (len(x) + seg_len - 1) // seg_len)
        rx = capi.Rx(max(1, min(workers, n_seg)), device=device, flags=flags)
    try:
        rx.execute_sharded(x, seg_len=seg_len, preroll=preroll)
        return rx.poll()
    finally:
        if own:
            rx.close()


def to_pdus(frame):
    """The messages `flex_rx` publishes for a frame, in publish order -- ('constellation', (None, complex64[])), then for a
    valid header ('payload_data', (None, bytes)) and ('packet_info', dict) -- lib/flex_rx_impl.cc:218-247."""
    from .blocks import MODULATION, INNER_CODE, OUTER_CODE, _index        # the block's own index tables
    if frame.get("flags", 0) & 1:         # dropped (longer than the receive buffer): the block publishes nothing
        return []
    syms = frame.get("framesyms")
    msgs = [("constellation", (None, np.asarray(syms if syms is not None else [], dtype=np.complex64)))]
    if frame["header_valid"]:
        msgs.append(("payload_data", (None, frame["payload"])))
        msgs.append(("packet_info", {
            "header_valid": 1, "payload_valid": int(bool(frame["payload_valid"])),
            "modulation": _index(MODULATION, frame["mod_scheme"], "Modulation Defaulting to BPSK"),
            "inner_code": _index(INNER_CODE, frame["fec0"], "FEC Defaulting to none"),
            "outer_code": _index(OUTER_CODE, frame["fec1"], "FEC defaulting to none")}))
    return msgs
