"""GPU parity for the frame generator (flex_tx path) and the bare detector (frame_detector_cc path)."""
import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


@pytest.mark.parametrize("ms", util.MODS + [31, 39, 40])
def test_tx_modulations_match_oracle(ms):
    rng = np.random.default_rng(ms)
    pl = rng.integers(0, 256, 211, dtype=np.uint8)
    hdr = rng.integers(0, 256, 14, dtype=np.uint8)
    tx = capi.Tx()
    got = tx.assemble([(ms, util.CRC24, 1, 1)], [pl], [hdr])[0]
    ref = o.tx_frame(ms, util.CRC24, 1, 1, pl, hdr)
    assert len(got) == len(ref)
    if 9 <= ms <= 16:      # DPSK symbols come from device sincosf: within float rounding
        assert np.allclose(got, ref, rtol=0, atol=2e-6)
    else:
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("f0", util.INNER + [12, 16, 23])
@pytest.mark.parametrize("f1", util.OUTER + [2, 3, 5])
def test_tx_code_pairs_bit_exact(f0, f1):
    rng = np.random.default_rng(100 * f0 + f1)
    tx = capi.Tx()
    for n in (1, 64, 300):
        pl = rng.integers(0, 256, n, dtype=np.uint8)
        got = tx.assemble([(util.QAM16, util.CRC24, f0, f1)], [pl])[0]
        ref = o.tx_frame(util.QAM16, util.CRC24, f0, f1, pl)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (f0, f1, n)


def test_tx_batch_and_reference_defaults():
    # flex_tx defaults: CRC-24, 14 zero header bytes (lib/flex_tx_impl.cc:52,58-59)
    rng = np.random.default_rng(9)
    tx = capi.Tx()
    props, pls = [], []
    for k in range(50):
        props.append((util.MODS[k % 11], util.CRC24, util.INNER[k % 7], util.OUTER[k % 8]))
        pls.append(rng.integers(0, 256, 1 + 31 * k, dtype=np.uint8))
    outs = tx.assemble(props, pls)
    for (ms, c, f0, f1), p, g in zip(props, pls, outs):
        ref = o.tx_frame(ms, c, f0, f1, p)
        assert np.allclose(g, ref, rtol=0, atol=2e-6)
    assert capi.Tx.frame_len(util.PSK4, util.CRC24, 1, 1, 256) == 2690
    assert capi.Tx.frame_len(util.PSK4, util.CRC24, 11, 27, 1500) == 28282
    assert capi.Tx.frame_len(util.QAM16, util.CRC24, 1, 1, 1500) == 6630


def test_tx_gpu_to_rx_gpu_loopback():
    rng = np.random.default_rng(10)
    tx = capi.Tx()
    pls = [rng.integers(0, 256, 1500, dtype=np.uint8) for _ in range(8)]
    frames = tx.assemble([(util.QAM16, util.CRC24, 1, 1)] * 8, pls)
    cap = util.build_capture(frames, rng, [900] * 8, snr_db=22.0, cfo=0.01, tau=0.2)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert [g["payload"] for g in got] == [p.tobytes() for p in pls]


def test_detector_matches_oracle_qdetector():
    rng = np.random.default_rng(11)
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 256, dtype=np.uint8)) for _ in range(12)]
    cap = util.build_capture(frames, rng, [800 + 301 * k for k in range(12)], snr_db=12.0, cfo=0.04, tau=0.3, gain=0.6)
    ref = o.detect_capture(cap, 0.3, 0.45)
    det = capi.Det(1)
    det.execute([cap])
    got = det.poll()
    assert len(ref) >= 12
    assert [g["sample_index"] for g in got] == [int(np.int64(np.uint64(r["sample_index"]))) for r in ref]
    for r, g in zip(ref, got):
        for k in ("tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy"):
            assert abs(g[k] - r[k]) <= 2e-5 + 1e-3 * abs(r[k]), (k, r[k], g[k])
    # streamed in 256-sample work() calls like the GR block (lib/frame_detector_cc_impl.cc:76)
    det2 = capi.Det(1)
    acc = []
    for i in range(0, len(cap), 256):
        det2.execute([cap[i:i + 256]])
        acc += det2.poll()
    assert [g["sample_index"] for g in acc] == [g["sample_index"] for g in got]


def test_detector_many_streams_noise_and_signal():
    rng = np.random.default_rng(12)
    caps = []
    for s in range(16):
        if s % 2:
            x = ((rng.standard_normal(30000) + 1j * rng.standard_normal(30000)) / np.sqrt(2)).astype(np.complex64)
        else:
            fr = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 100, dtype=np.uint8)) for _ in range(5)]
            x = util.build_capture(fr, rng, [3000] * 5, snr_db=5.0 + s, cfo=0.05 * (s / 16 - 0.5))
        caps.append(x)
    det = capi.Det(16)
    det.execute(caps)
    got = det.poll()
    for s in range(16):
        ref = o.detect_capture(caps[s], 0.3, 0.45)
        mine = [g for g in got if g["stream"] == s]
        assert [g["sample_index"] for g in mine] == [int(np.int64(np.uint64(r["sample_index"]))) for r in ref]


def test_bulk_time_sharded_detection_matches_sequential_oracle():
    # configs[1] flavour: one long capture, frames at jittered 8192-sample spacing, CFO U(+-0.05), SNR sweep
    from liquiddsp import bulk
    rng = np.random.default_rng(13)
    base = o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 256, dtype=np.uint8))
    n_frames, spacing = 60, 8192
    x = np.zeros(n_frames * spacing + 4096, np.complex128)
    starts = []
    for k in range(n_frames):
        s0 = k * spacing + int(rng.integers(0, spacing - len(base) - 64))
        cfo, ph = rng.uniform(-0.05, 0.05), rng.uniform(0, 2 * np.pi)
        x[s0:s0 + len(base)] += base * np.exp(1j * (cfo * np.arange(len(base)) + ph))
        starts.append(s0)
    snr_db = np.repeat(np.linspace(4.0, 20.0, n_frames), spacing)[:len(x)]
    snr_db = np.concatenate([snr_db, np.full(len(x) - len(snr_db), 20.0)])
    x = x + 10 ** (-snr_db / 20) * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) / np.sqrt(2)
    x = x.astype(np.complex64)
    det = bulk.BulkDetector(8)
    got = det.run_host(x, seg_len=1 << 16)
    ref = bulk.dedup([dict(d, sample_index=int(np.int64(np.uint64(d["sample_index"])))) for d in o.detect_capture(x, 0.3, 0.45)])
    gi = np.array([g["sample_index"] for g in got])
    ri = np.array([r["sample_index"] for r in ref])
    # every transmitted frame is found by both, at the same sample (+-1), seams included
    for s0 in starts:
        assert np.abs(gi - s0).min() <= 1 and np.abs(ri - s0).min() <= 1
    # and the two detection lists agree except for hop-grid effects on spurious re-triggers
    common = sum(np.abs(ri - g).min() <= 1 for g in gi)
    assert common >= 0.95 * len(gi) and abs(len(gi) - len(ri)) <= 0.05 * len(ri) + 2
