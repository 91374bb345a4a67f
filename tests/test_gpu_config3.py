"""GPU parity at the size of the benched configuration (BASELINE.json configs[2]): PSK4, inner v27 (fec0),
outer RS(255,223) (fec1), 1500-byte payloads -- 14 RS blocks and a 12 030-step trellis per frame, 28 282 samples --
across the bench's SNR sweep, against the CPU oracle on the same captures.  Replaces the receive path behind
/root/reference/lib/flex_rx_impl.cc:49,213 (flexframesync_create / _execute); frames as /root/reference/lib/
flex_tx_impl.cc:56,198-201 generates them.

Bars (north_star): bytes and flags bit-exact, estimates within 1e-3 relative, PER near the decode threshold within
the 95 % binomial interval of the oracle's."""
import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi
from test_gpu_parity import assert_frames_match

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]

V27, RS8 = 11, 27
PAYLOAD = 1500


def _frames(rng, n):
    pls = [rng.integers(0, 256, PAYLOAD, dtype=np.uint8) for _ in range(n)]
    return pls, [o.tx_frame(util.PSK4, util.CRC24, V27, RS8, p) for p in pls]


def test_cfg3_frame_geometry():
    # 1503 -> conv 3008 B -> 14 RS blocks of 215 + 32 -> 3458 B -> 13 832 symbols -> 28 282 samples
    assert capi.tab_packet_len(PAYLOAD, util.CRC24, V27, RS8, util.PSK4) == (3458, 13832)
    assert capi.Tx.frame_len(util.PSK4, util.CRC24, V27, RS8, PAYLOAD) == 28282


def test_cfg3_size_snr_sweep_matches_oracle():
    """15 streams, one per SNR point -2 .. +12 dB, three 1500-byte v27 + RS8 frames each, CFO within +-0.02 rad/sample,
    timing offset within +-0.5 sample, gain 0.5 .. 1.5: every frame record (position, header, flags, payload bytes,
    estimates, constellation) equals the oracle's -- decodable, CRC-failed and header-failed frames alike."""
    rng = np.random.default_rng(306)                 # (a seed whose sweep holds all three outcomes)
    caps, refs = [], []
    for k in range(15):
        snr = -2.0 + k
        _, frames = _frames(rng, 3)
        cfo = (-0.02, 0.013, 0.02, -0.007, 0.0)[k % 5]
        tau = (-0.5, 0.31, 0.5, -0.12, 0.0)[(k + 2) % 5]
        cap = util.build_capture(frames, rng, [3000 + 400 * (k % 7)] * 3, snr_db=snr, cfo=cfo, tau=tau, gain=0.5 + k / 14.0)
        caps.append(cap)
        refs.append(o.rx_capture(cap))
    rx = capi.Rx(len(caps), max_frame_samples=65536)
    rx.execute(caps)
    got = rx.poll()
    n_valid = n_hdr_fail = n_crc_fail = 0
    for s, ref in enumerate(refs):
        mine = [g for g in got if g["stream"] == s]
        assert_frames_match(ref, mine)
        n_valid += sum(1 for r in ref if r["payload_valid"])
        n_hdr_fail += sum(1 for r in ref if not r["header_valid"])
        n_crc_fail += sum(1 for r in ref if r["header_valid"] and not r["payload_valid"])
    # the sweep really covers the three outcomes (else the test would say nothing about failed frames)
    assert n_valid >= 15 and n_hdr_fail >= 3 and n_crc_fail >= 1, (n_valid, n_hdr_fail, n_crc_fail)
    # at and above 7 dB (>= 3 dB over the knee near 3 dB, where the header gives out first) everything sent is decoded
    for s in range(9, 15):
        assert sum(1 for r in refs[s] if r["payload_valid"]) == 3


def test_cfg3_streamed_in_chunks_equals_one_shot():
    """The same kind of capture fed 4096 samples at a time through submit / collect (frames straddle many calls and
    both pipeline generations): same frame records as the oracle's single pass."""
    rng = np.random.default_rng(304)
    _, frames = _frames(rng, 2)
    cap = util.build_capture(frames, rng, [2500, 2500], snr_db=9.0, cfo=0.011, tau=-0.4, gain=1.2)
    ref = o.rx_capture(cap)
    assert len(ref) == 2 and all(r["payload_valid"] for r in ref)
    rx = capi.Rx(1, max_frame_samples=65536)
    got, pending = [], 0
    n = len(cap) // 4096 * 4096
    for i in range(0, n, 4096):
        rx.submit([np.ascontiguousarray(cap[i:i + 4096])])
        pending += 1
        if pending == 2:
            rx.collect(); got += rx.poll(); pending -= 1
    rx.submit([np.ascontiguousarray(cap[n:])])
    pending += 1
    while pending:
        rx.collect(); got += rx.poll(); pending -= 1
    assert_frames_match(ref, got)


def test_cfg3_per_at_the_knee_within_the_oracles_interval():
    """PER of the benched scheme around its decode threshold (2 .. 3.5 dB; the QPSK header gives out a little before the
    v27 + RS8 payload): GPU and oracle on the same captures must agree within the oracle's 95 % binomial interval at
    every point; the sweep must straddle the knee."""
    rng = np.random.default_rng(305)
    n_per_point, per_stream = 32, 8
    points = (2.0, 2.75, 3.5)
    caps, refs = [], []
    for snr in points:
        for _ in range(n_per_point // per_stream):
            _, frames = _frames(rng, per_stream)
            cap = util.build_capture(frames, rng, [2200] * per_stream, snr_db=snr, cfo=0.004, tau=0.2)
            caps.append(cap)
            refs.append(o.rx_capture(cap))
    rx = capi.Rx(len(caps), max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS)
    rx.execute(caps)
    got = rx.poll()
    spp = n_per_point // per_stream
    pers = []
    for k, snr in enumerate(points):
        ref = [r for rr in refs[k * spp:(k + 1) * spp] for r in rr]
        mine = [g for g in got if k * spp <= g["stream"] < (k + 1) * spp]
        assert abs(len(mine) - len(ref)) <= 1
        per_ref = 1.0 - sum(r["payload_valid"] for r in ref) / n_per_point          # frames not found count as errors
        per_got = 1.0 - sum(g["payload_valid"] for g in mine) / n_per_point
        half = 1.96 * np.sqrt(max(per_ref * (1.0 - per_ref), 1e-3) / n_per_point)
        assert abs(per_got - per_ref) <= half, (snr, per_ref, per_got)
        pers.append(per_ref)
    assert max(pers) > 0.1 and min(pers) < 0.9, pers       # genuinely around the threshold


def test_receivers_on_two_devices_in_one_process():
    """Function attributes (the search kernel's dynamic shared memory opt-in) are per device: a second handle on
    another GPU of the same process must work.  Needs two GPUs; one-GPU boxes skip."""
    if capi.lib().lqb_device_count() < 2:
        pytest.skip("one GPU visible")
    rng = np.random.default_rng(306)
    pl = rng.integers(0, 256, 300, dtype=np.uint8)
    cap = util.build_capture([o.tx_frame(util.PSK4, util.CRC24, V27, RS8, pl)], rng, [900], snr_db=15.0, cfo=0.01)
    ref = o.rx_capture(cap)
    for dev in (0, 1):
        rx = capi.Rx(1, device=dev)
        rx.execute([cap])
        assert_frames_match(ref, rx.poll())
        rx.close()
        det = capi.Det(1, device=dev)
        det.execute([cap])
        assert len(det.poll()) >= 1
        det.close()


def test_sc16_input_equals_the_widened_floats():
    """LQB_MEM_HOST_SC16 (interleaved int16 pairs, value / 32768 -- additive to the reference's complex64): the receiver
    and the detector give exactly what they give for the widened floats, which is what the oracle gives for them."""
    rng = np.random.default_rng(307)
    pls = [rng.integers(0, 256, 400, dtype=np.uint8) for _ in range(3)]
    frames = [o.tx_frame(util.PSK4, util.CRC24, V27, RS8, p) for p in pls]
    cap = util.build_capture(frames, rng, [1500] * 3, snr_db=12.0, cfo=0.015, tau=0.3, gain=0.9)
    q = np.clip(np.round(np.stack([cap.real, cap.imag], axis=1) * 6000.0), -32768, 32767).astype(np.int16)
    wide = (q[:, 0].astype(np.float32) / 32768.0 + 1j * (q[:, 1].astype(np.float32) / 32768.0)).astype(np.complex64)
    ref = o.rx_capture(wide)
    assert len(ref) == 3 and all(r["payload_valid"] for r in ref)
    rx = capi.Rx(2)
    rx.execute_sc16([q, q[:4001]])                       # second stream: a ragged, odd-length piece of the same capture
    got = rx.poll()
    assert_frames_match(ref, [g for g in got if g["stream"] == 0])
    rx2 = capi.Rx(1)
    rx2.execute([wide])
    assert_frames_match(ref, rx2.poll())
    # streamed in odd-sized chunks through the same format
    rx3 = capi.Rx(1)
    got3 = []
    for i in range(0, len(q), 3001):
        rx3.execute_sc16([q[i:i + 3001]])
        got3 += rx3.poll()
    assert_frames_match(ref, got3)


@pytest.mark.parametrize("fec", [12, 21, 22, 23, 24, 25, 26])
def test_k9_codes_sixteen_lane_viterbi_matches_oracle(fec):
    """v29 and its punctured rates on the sixteen-lane K = 9 decoder (k_viterbi29x16): ragged lengths (two codewords share a
    warp and run to the longer one), a 1500-byte frame, a one-byte frame, as inner code alone and under RS(255,223), near
    the knee so that decisions carry errors."""
    rng = np.random.default_rng(900 + fec)
    lens = [1500, 1, 333, 64, 257, 9, 800]
    pls = [rng.integers(0, 256, n, dtype=np.uint8) for n in lens]
    frames = [o.tx_frame(util.PSK4, util.CRC24, fec, 27 if k % 2 else 1, p) for k, p in enumerate(pls)]
    cap = util.build_capture(frames, rng, [700] * len(frames), snr_db=5.5, cfo=0.011, tau=0.2)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert len(ref) == len(got) == len(lens)
    for r, g in zip(ref, got):
        assert r["sample_index"] == g["sample_index"] and r["header_valid"] == g["header_valid"] == 1
        assert r["payload_valid"] == g["payload_valid"] and r["payload"] == g["payload"]      # the decoded bytes, right or wrong
    assert sum(r["payload_valid"] for r in ref) >= 3
