"""One capture decoded by the time-sharded receiver (lqb_rx_execute_sharded) against ONE sequential oracle receiver:
the single-stream case of the reference's block (/root/reference/lib/flex_rx_impl.cc:213, one flexframesync per
block) spread over many CTAs.  Frames, flags, payload bytes and estimates must not depend on the cut."""
import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def _capture(rng):
    """Mixed frames: short and long (longer than the small pre-rolls), several schemes, weak ones whose header fails,
    back-to-back frames, long gaps, noise."""
    specs = [(util.PSK4, 11, 27, 1500), (util.QAM16, 1, 1, 300), (util.PSK8, 15, 7, 64), (util.PSK2, 11, 1, 700),
             (util.QAM64, 11, 1, 1200), (util.DPSK4, 1, 6, 33), (util.PSK4, 1, 1, 9)]
    parts = [np.zeros(500, np.complex64)]
    sent = []
    for k in range(34):
        ms, f0, f1, n = specs[k % len(specs)]
        pl = rng.integers(0, 256, n, dtype=np.uint8)
        fr = o.tx_frame(ms, util.CRC24, f0, f1, pl)
        g = 10.0 ** ([18.0, 25.0, 0.5, 14.0, 30.0, -1.0][k % 6] / 20.0)       # some preambles / headers are hopeless
        parts.append((g * fr).astype(np.complex64))
        parts.append(np.zeros([40, 900, 5000, 300, 16000][k % 5], np.complex64))
        sent.append(pl.tobytes())
    x = np.concatenate(parts)
    n = np.arange(len(x))
    x = x * np.exp(1j * (0.011 * n + 1.0))
    x = x + (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) / np.sqrt(2.0)
    return x.astype(np.complex64), sent


@pytest.mark.parametrize("seg_len,preroll,workers", [(8192, 0, 9), (16384, 4096, 32), (65536, 32768, 6), (1 << 18, 1 << 16, 3), (1 << 20, 1 << 16, 2)])
def test_time_sharded_receiver_is_the_sequential_receiver(seg_len, preroll, workers):
    rng = np.random.default_rng(99)
    cap, sent = _capture(rng)
    cap = cap[:len(cap) - 9000]                      # the capture ends inside the last frame
    ref = o.rx_capture(cap)
    rx = capi.Rx(workers, max_frame_samples=32768)
    rx.execute_sharded(cap, seg_len=seg_len, preroll=preroll)
    got = rx.poll()
    info = rx.shard_info()
    assert len(ref) >= 25 and sum(r["payload_valid"] for r in ref) >= 15
    assert [g["sample_index"] for g in got] == [r["sample_index"] for r in ref], info
    for r, g in zip(ref, got):
        assert g["stream"] == 0
        for k in ("header_valid", "payload_valid", "payload_len", "mod_scheme", "check", "fec0", "fec1", "num_framesyms"):
            assert r[k] == g[k], (k, r[k], g[k])
        assert r["header"] == g["header"]
        if r["header_valid"]:
            assert r["payload"] == g["payload"]
            assert np.all(np.abs(g["framesyms"] - r["framesyms"]) <= 1e-5 + 1e-5 * np.abs(r["framesyms"]))
        for k in ("rssi", "cfo", "tau_hat", "gamma_hat", "rxy") + (("evm",) if r["header_valid"] else ()):
            assert abs(g[k] - r[k]) <= 2e-5 + 1e-3 * abs(r[k]), (k, r[k], g[k])
    assert [g["seq"] for g in got] == list(range(len(got)))
    assert rx.counts() == (len(ref), sum(1 for r in ref if r["header_valid"] and r["payload_valid"]))
    assert info["segments"] == (len(cap) + seg_len - 1) // seg_len
    if seg_len <= 16384:
        assert info["runs"] > info["segments"]       # frames longer than the pre-roll: seams were re-run
    # the handle goes back to ordinary streaming after a reset
    rx.reset()
    rx.execute([cap], [0])
    assert [g["sample_index"] for g in rx.poll()] == [r["sample_index"] for r in ref]


def test_time_sharded_receiver_soft_and_empty():
    rx = capi.Rx(4, flags=capi.RX_SOFT)
    rx.execute_sharded(np.zeros(300, np.complex64))
    assert rx.poll() == [] and rx.counts() == (0, 0)
    rng = np.random.default_rng(3)
    pls = [rng.integers(0, 256, 200, dtype=np.uint8) for _ in range(12)]
    cap = util.build_capture([o.tx_frame(util.QAM16, util.CRC24, 11, 1, p) for p in pls], rng, [700] * 12, snr_db=8.0, cfo=0.01)
    ref = o.rx_capture(cap, soft=True)
    rx.execute_sharded(cap, seg_len=8192, preroll=4096)
    got = rx.poll()
    assert [(g["sample_index"], g["payload_valid"], g["payload"]) for g in got] == [(r["sample_index"], r["payload_valid"], r["payload"]) for r in ref]
