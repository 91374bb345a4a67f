"""Soft-decision payload decoding (opt-in extension LQB_RX_SOFT).

The reference's blocks use liquid's default hard decisions (/root/reference/lib/flex_rx_impl.cc:49 creates the
synchroniser and never asks for soft decoding), so this path is additive: its definition is the oracle's
(oracle/lqo.h: lqo_modem_demodulate_soft, lqo_fec_decode_soft, lqo_qpm_decode_soft) and the CUDA kernels are held to it
bit for bit.  CPU tests pin the definition's properties; the GPU tests compare through the C-ABI."""
import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi

V27, V29, V27P23, V27P34, V27P78, V29P23, RS8 = 11, 12, 15, 16, 20, 21, 27


def _frames_at(ms, f0, f1, n, snr_db, count, seed, **chan):
    rng = np.random.default_rng(seed)
    pls = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(count)]
    frames = [o.tx_frame(ms, util.CRC24, f0, f1, p) for p in pls]
    cap = util.build_capture(frames, rng, [900] * count, snr_db=snr_db, **chan)
    return pls, cap


# ------------------------------------------------------------------ the definition (CPU, oracle only)
def test_soft_equals_hard_on_a_clean_channel():
    for ms, f0, f1 in [(util.PSK4, V27, 1), (util.QAM16, 1, V27), (util.PSK8, V27P34, 1), (util.QAM64, V29, 1)]:
        pls, cap = _frames_at(ms, f0, f1, 120, 35.0, 2, 100 + ms)
        hard, soft = o.rx_capture(cap), o.rx_capture(cap, soft=True)
        assert len(hard) == len(soft) == 2
        for h, s, p in zip(hard, soft, pls):
            assert h["payload_valid"] and s["payload_valid"] and h["payload"] == s["payload"] == p.tobytes()
            for k in ("sample_index", "evm", "rssi", "cfo"):
                assert h[k] == s[k]


def test_soft_falls_back_to_hard_where_it_does_not_apply():
    # DPSK has no soft demodulator here; v27 under rs8 is not the stage nearest the channel; block codes are hard
    for ms, f0, f1 in [(util.DPSK4, V27, 1), (util.PSK4, V27, RS8), (util.PSK4, 7, 1)]:
        pls, cap = _frames_at(ms, f0, f1, 90, 9.0, 3, 200 + ms + f1)
        hard, soft = o.rx_capture(cap), o.rx_capture(cap, soft=True)
        assert [(f["payload_valid"], f["payload"]) for f in hard] == [(f["payload_valid"], f["payload"]) for f in soft]


def test_soft_decisions_buy_packets_near_the_knee():
    # QAM16 + v27, 200 bytes at 8 dB (the QPSK header is safe there): hard decisions lose a third of the frames, soft none
    pls, cap = _frames_at(util.QAM16, V27, 1, 200, 8.0, 40, 7)
    hard, soft = o.rx_capture(cap), o.rx_capture(cap, soft=True)
    nh = sum(f["payload_valid"] for f in hard)
    ns = sum(f["payload_valid"] for f in soft)
    assert len(hard) == len(soft)
    assert ns >= nh + 8, (nh, ns)
    for f in soft:
        if f["payload_valid"]:
            assert any(f["payload"] == p.tobytes() for p in pls)


# ------------------------------------------------------------------ the CUDA path (through the C-ABI)
SOFT_CASES = [
    (util.PSK2, V27, 1, 5.0), (util.PSK4, V27, 1, 4.5), (util.PSK8, V27, 1, 7.0), (util.PSK16, V27P23, 1, 13.0),
    (util.ASK4, V27, 1, 9.0), (util.QAM16, V27, 1, 8.0), (util.QAM32, V27P78, 1, 15.0), (util.QAM64, V27, 1, 12.5),
    (util.QAM16, 1, V27, 8.0), (util.PSK4, 7, V27, 4.5), (util.PSK8, RS8, V27P34, 9.0),
    (util.PSK8, V29, 1, 6.5), (util.QAM16, 1, V29P23, 9.0), (31, V27, 1, 18.0),
    (util.DPSK4, V27, 1, 8.0), (util.PSK4, V27, RS8, 4.5),           # fall-backs: same as hard
]


@pytest.mark.gpu
@pytest.mark.usefixtures("gpu_required")
@pytest.mark.parametrize("ms,f0,f1,snr", SOFT_CASES)
def test_soft_payloads_match_the_oracle_bit_for_bit(ms, f0, f1, snr):
    pls, cap = _frames_at(ms, f0, f1, 333, snr, 6, 4000 + 64 * ms + 8 * f0 + f1, cfo=0.013, tau=-0.2, gain=0.9)
    ref = o.rx_capture(cap, soft=True)
    rx = capi.Rx(1, flags=capi.RX_SOFT)
    rx.execute([cap])
    got = rx.poll()
    assert len(got) == len(ref) >= 5
    for r, g in zip(ref, got):
        assert r["sample_index"] == g["sample_index"] and r["header_valid"] == g["header_valid"]
        assert r["payload_valid"] == g["payload_valid"], (r["sample_index"],)
        if r["header_valid"]:
            assert r["payload"] == g["payload"]                 # the decoded bytes, valid or not
            assert np.all(np.abs(g["framesyms"] - r["framesyms"]) <= 1e-5 + 1e-5 * np.abs(r["framesyms"]))
    # the cases sit near each scheme's knee: the soft path must have had errors to correct (or to fail on)
    hard = o.rx_capture(cap)
    assert sum(f["payload_valid"] for f in ref) >= sum(f["payload_valid"] for f in hard)


@pytest.mark.gpu
@pytest.mark.usefixtures("gpu_required")
def test_soft_flag_off_is_the_hard_path_and_mixed_batches_work():
    # one call, four streams: soft-eligible and not, long and short frames, several chunks
    rng = np.random.default_rng(77)
    specs = [(util.PSK4, V27, 1, 1500, 4.0), (util.QAM16, V27, RS8, 700, 12.0), (util.PSK8, 1, V27P34, 50, 12.0), (util.DPSK2, V27, 1, 9, 9.0)]
    caps = []
    for ms, f0, f1, n, snr in specs:
        pls = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(4)]
        caps.append(util.build_capture([o.tx_frame(ms, util.CRC24, f0, f1, p) for p in pls], rng, [800] * 4, snr_db=snr, cfo=-0.01))
    for flags, soft in ((capi.RX_SOFT, True), (0, False)):
        rx = capi.Rx(len(caps), flags=flags)
        refs = [o.rx_capture(c, soft=soft) for c in caps]
        got = [[] for _ in caps]
        n = max(len(c) for c in caps)
        for a in range(0, n, 20000):
            rx.execute([c[a:a + 20000] for c in caps])
            for f in rx.poll():
                got[f["stream"]].append(f)
        for r, g in zip(refs, got):
            assert [(f["sample_index"], f["payload_valid"], f["payload"]) for f in r] == \
                   [(f["sample_index"], f["payload_valid"], f["payload"]) for f in g]


@pytest.mark.gpu
@pytest.mark.usefixtures("gpu_required")
def test_soft_decisions_lower_the_packet_error_rate_on_the_gpu():
    pls, cap = _frames_at(util.QAM16, V27, 1, 200, 8.0, 60, 8)
    n_ok = []
    for flags in (0, capi.RX_SOFT):
        rx = capi.Rx(1, flags=flags)
        rx.execute([cap])
        n_ok.append(sum(f["payload_valid"] for f in rx.poll()))
    assert n_ok[1] >= n_ok[0] + 10, n_ok
