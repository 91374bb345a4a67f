"""TX oracle -> channel -> RX oracle loopback (BASELINE.json configs[0] and its siblings):
the loopback the reference names (qa_flex_tx/qa_flex_rx) but never implemented
(/root/reference/python/qa_flex_rx.py:34-37)."""
import numpy as np
import pytest

import lqo_py as o
import util


def _one(ms, f0, f1, n, rng, **chan):
    pl = rng.integers(0, 256, n, dtype=np.uint8)
    x = o.tx_frame(ms, util.CRC24, f0, f1, pl)
    fr = o.rx_capture(util.impair(x, rng, pre=900, post=900, **chan))
    return pl, fr


def test_cfg1_loopback_many_frames():
    rng = np.random.default_rng(11)
    pls = [rng.integers(0, 256, 256, dtype=np.uint8) for _ in range(20)]
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, p) for p in pls]
    cap = util.build_capture(frames, rng, [1024] * 20, snr_db=30.0)
    got = o.rx_capture(cap)
    assert len(got) == 20
    for p, f in zip(pls, got):
        assert f["header_valid"] and f["payload_valid"] and f["payload"] == p.tobytes()
        assert f["num_framesyms"] == 1036 and f["mod_scheme"] == util.PSK4 and f["check"] == util.CRC24
    starts = [f["sample_index"] for f in got]
    assert starts == [700 + i * (2690 + 1024) for i in range(20)]


@pytest.mark.parametrize("ms", util.MODS)
def test_every_block_api_modulation(ms):
    rng = np.random.default_rng(ms)
    pl, fr = _one(ms, 1, 1, 200, rng, snr_db=40.0, cfo=0.02, tau=0.3, gain=0.5, phi=1.0)
    assert len(fr) == 1 and fr[0]["payload_valid"] and fr[0]["payload"] == pl.tobytes()
    assert abs(fr[0]["rssi"] - 20 * np.log10(0.5)) < 0.5
    assert abs(fr[0]["cfo"] - 0.02) < 2e-3


@pytest.mark.parametrize("f0", util.INNER)
@pytest.mark.parametrize("f1", util.OUTER)
def test_every_block_api_code_pair(f0, f1):
    rng = np.random.default_rng(1000 + 10 * f0 + f1)
    pl, fr = _one(util.PSK4, f0, f1, 100, rng, snr_db=12.0, cfo=-0.03, tau=-0.2)
    assert len(fr) == 1 and fr[0]["payload_valid"] and fr[0]["payload"] == pl.tobytes()
    assert fr[0]["fec0"] == f0 and fr[0]["fec1"] == f1


def test_noise_only_finds_nothing():
    rng = np.random.default_rng(12)
    x = ((rng.standard_normal(60000) + 1j * rng.standard_normal(60000)) / np.sqrt(2)).astype(np.complex64)
    assert o.rx_capture(x) == []


def test_chunking_does_not_change_results():
    rng = np.random.default_rng(13)
    pl = rng.integers(0, 256, 64, dtype=np.uint8)
    x = util.impair(o.tx_frame(util.QAM16, util.CRC24, 11, 1, pl), rng, snr_db=25, cfo=0.01, tau=0.1, pre=333, post=800)
    a = o.rx_capture(x, chunk=256)
    b = o.rx_capture(x, chunk=1)
    c = o.rx_capture(x, chunk=100000)
    assert len(a) == len(b) == len(c) == 1
    for k in ("sample_index", "evm", "tau_hat", "payload"):
        assert a[0][k] == b[0][k] == c[0][k]


def test_detector_matches_frame_positions():
    rng = np.random.default_rng(14)
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 256, dtype=np.uint8)) for _ in range(5)]
    cap = util.build_capture(frames, rng, [2000] * 5, snr_db=20.0, cfo=0.03)
    det = o.detect_capture(cap, 0.3, 0.45)
    starts = [d["sample_index"] for d in det]
    for i in range(5):
        assert any(abs(s - (700 + i * 4690)) <= 1 for s in starts)
