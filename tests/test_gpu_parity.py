"""GPU parity: the CUDA path behind the C-ABI against the CPU oracle on identical captures.
Bars (BASELINE.json north_star): header/payload bytes and CRC flags bit-exact at SNR >= threshold
+ 3 dB; EVM / RSSI / CFO / timing within 1e-3 relative (the tolerance asserted below);
PER near threshold within a 95 % binomial interval."""
import ctypes as C
import os

import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]

STAT_KEYS = ("evm", "rssi", "cfo", "tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy")
RTOL = 1e-3


def close(a, b, rtol=RTOL, atol=2e-5):
    return abs(a - b) <= atol + rtol * abs(b)


def assert_frames_match(ref, got, syms=True):
    assert len(got) == len(ref), (len(got), len(ref))
    for r, g in zip(ref, got):
        for k in ("sample_index", "header_valid", "payload_valid", "payload_len", "num_framesyms",
                  "mod_scheme", "mod_bps", "check", "fec0", "fec1"):
            assert r[k] == g[k], (k, r[k], g[k])
        assert r["header"] == g["header"]
        if r["header_valid"]:
            assert r["payload"] == g["payload"]
            keys = STAT_KEYS
        else:
            keys = ("rssi", "cfo", "tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy")
        for k in keys:
            assert close(g[k], r[k]), (k, r[k], g[k])
        if syms and r["header_valid"]:
            # constellation points: every one within float rounding of the oracle's.  (Until arg() / exp(j t) of the
            # signal path were pinned to one operation sequence on both sides, a last-bit difference in a phase
            # estimate could move points by one step of the 1024-entry NCO table, 2 pi / 1024 rad.)
            err = np.abs(g["framesyms"] - r["framesyms"])
            mag = np.abs(r["framesyms"])
            assert np.all(err <= 1e-5 + 1e-5 * mag)


def test_fft512_warp_kernel_is_bit_exact_with_the_oracle_fft():
    L = capi.lib()
    L.lqb_dbg_fft512.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(1)
    for d in (1, -1):
        x = (rng.standard_normal(512) + 1j * rng.standard_normal(512)).astype(np.complex64)
        y = np.zeros(512, np.complex64)
        yo = np.zeros(512, np.complex64)
        assert L.lqb_dbg_fft512(x.ctypes.data, y.ctypes.data, d) == 0
        o.lib().lqo_fft(x.ctypes.data, yo.ctypes.data, 512, d)
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32))


def test_pinned_arg_and_sincos_match_the_oracle_bit_for_bit():
    # arg() / exp(j t) inside the PSK / DPSK symbol loops are a fixed sequence of IEEE operations on both sides
    L = capi.lib()
    fp = C.c_void_p
    L.lqb_dbg_pm.argtypes = [fp, fp, fp, fp, fp, C.c_uint]
    O = o.lib()
    O.lqo_pm_atan2f.restype = C.c_float
    O.lqo_pm_atan2f.argtypes = [C.c_float, C.c_float]
    O.lqo_pm_sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    rng = np.random.default_rng(2)
    n = 20000
    y = np.concatenate([rng.uniform(-7.0, 7.0, n - 8), [0.0, -0.0, 1.0, -1.0, 1e-30, 3.14159274, -3.14159274, 6.2831855]]).astype(np.float32)
    x = np.concatenate([rng.standard_normal(n - 8), [0.0, 1.0, 0.0, -0.0, -1.0, 1.0, -1.0, 1e-20]]).astype(np.float32)
    at, sn, cs = (np.zeros(n, np.float32) for _ in range(3))
    assert L.lqb_dbg_pm(y.ctypes.data, x.ctypes.data, at.ctypes.data, sn.ctypes.data, cs.ctypes.data, n) == 0
    ro = np.zeros((3, n), np.float32)
    s_, c_ = C.c_float(), C.c_float()
    for i in range(n):
        ro[0, i] = O.lqo_pm_atan2f(float(y[i]), float(x[i]))
        O.lqo_pm_sincosf(float(y[i]), C.byref(s_), C.byref(c_))
        ro[1, i], ro[2, i] = s_.value, c_.value
    assert np.array_equal(at.view(np.uint32), ro[0].view(np.uint32))
    assert np.array_equal(sn.view(np.uint32), ro[1].view(np.uint32))
    assert np.array_equal(cs.view(np.uint32), ro[2].view(np.uint32))
    # and they are the functions they claim to be
    assert np.max(np.abs(at - np.arctan2(y.astype(np.float64), x.astype(np.float64)))) < 5e-7
    assert np.max(np.abs(sn - np.sin(y.astype(np.float64)))) < 2e-7 and np.max(np.abs(cs - np.cos(y.astype(np.float64)))) < 2e-7


def test_golden_capture_decodes_to_golden_bytes():
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loopback_v1.npz"))
    rx = capi.Rx(1)
    rx.execute([G["capture"]])
    got = rx.poll()
    assert [f["sample_index"] for f in got] == G["sample_index"].tolist()
    for f, pl in zip(got, (G["payload1"], G["payload3"], G["payload5"])):
        assert f["header_valid"] and f["payload_valid"] and f["payload"] == pl.tobytes()
    stats = np.array([[f[k] for k in STAT_KEYS] for f in got], np.float32)
    assert np.allclose(stats, G["stats"], rtol=RTOL, atol=2e-5)
    assert np.allclose(got[0]["framesyms"][:64], G["syms0_head"], rtol=RTOL, atol=1e-4)


def test_cfg1_loopback_bit_exact():
    rng = np.random.default_rng(21)
    pls = [rng.integers(0, 256, 256, dtype=np.uint8) for _ in range(40)]
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, p) for p in pls]
    cap = util.build_capture(frames, rng, [1024] * 40, snr_db=30.0)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert len(ref) == 40
    assert_frames_match(ref, got)
    assert [g["payload"] for g in got] == [p.tobytes() for p in pls]


@pytest.mark.parametrize("ms", util.MODS + [30, 31, 39, 40, 17, 5])
def test_modulations_match_oracle(ms):
    rng = np.random.default_rng(300 + ms)
    pl = rng.integers(0, 256, 333, dtype=np.uint8)
    x = o.tx_frame(ms, util.CRC24, 1, 1, pl)
    cap = util.impair(x, rng, snr_db=45.0, cfo=0.017, tau=0.27, gain=0.8, phi=0.4, pre=811, post=900)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert len(ref) == 1 and ref[0]["payload_valid"]
    assert_frames_match(ref, got)


@pytest.mark.parametrize("f0", util.INNER + [12, 16, 22])
@pytest.mark.parametrize("f1", util.OUTER + [2, 3, 5])
def test_code_pairs_match_oracle(f0, f1):
    rng = np.random.default_rng(5000 + 40 * f0 + f1)
    pl = rng.integers(0, 256, 257, dtype=np.uint8)
    x = o.tx_frame(util.PSK4, util.CRC24, f0, f1, pl)
    cap = util.impair(x, rng, snr_db=11.0, cfo=-0.021, tau=-0.31, gain=1.3, pre=650, post=900)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert len(ref) == 1
    assert_frames_match(ref, got)


@pytest.mark.parametrize("check", [1, 2, 3, 4, 5, 6])
def test_crc_schemes_match_oracle(check):
    rng = np.random.default_rng(70 + check)
    pl = rng.integers(0, 256, 100, dtype=np.uint8)
    cap = util.impair(o.tx_frame(util.QAM16, check, 11, 7, pl), rng, snr_db=25.0, pre=700, post=900)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    assert_frames_match(ref, rx.poll())


def test_many_ragged_streams_and_streaming_chunks():
    rng = np.random.default_rng(31)
    n_streams = 12
    caps, refs = [], []
    for s in range(n_streams):
        ms = util.MODS[s % len(util.MODS)]
        f0 = util.INNER[s % len(util.INNER)]
        f1 = util.OUTER[(3 * s) % len(util.OUTER)]
        frames = [o.tx_frame(ms, util.CRC24, f0, f1, rng.integers(0, 256, 40 + 37 * k + s, dtype=np.uint8)) for k in range(4)]
        cap = util.build_capture(frames, rng, [300 + 211 * k for k in range(4)], snr_db=28.0,
                                 cfo=0.03 * (s / n_streams - 0.5), tau=0.45 * (s % 3 - 1), gain=0.5 + 0.1 * s,
                                 lead=100 + 97 * s, tail=800 + 13 * s)
        caps.append(cap)
        refs.append(o.rx_capture(cap))
    # one shot, ragged lengths
    rx = capi.Rx(n_streams)
    rx.execute(caps)
    got = rx.poll()
    for s in range(n_streams):
        assert_frames_match(refs[s], [g for g in got if g["stream"] == s])
    # streamed in uneven chunks, different subset of streams per call
    rx2 = capi.Rx(n_streams)
    pos = [0] * n_streams
    acc = [[] for _ in range(n_streams)]
    step = 0
    while any(pos[s] < len(caps[s]) for s in range(n_streams)):
        ids, chunks = [], []
        for s in range(n_streams):
            if pos[s] >= len(caps[s]) or (step + s) % 3 == 0:
                continue
            sz = [256, 1000, 4096, 77, 2560][(step + s) % 5]
            ids.append(s)
            chunks.append(caps[s][pos[s]:pos[s] + sz])
            pos[s] += sz
        if ids:
            rx2.execute(chunks, ids)
            for g in rx2.poll():
                acc[g["stream"]].append(g)
        step += 1
    for s in range(n_streams):
        assert_frames_match(refs[s], acc[s])
        assert [g["seq"] for g in acc[s]] == list(range(len(acc[s])))


@pytest.mark.parametrize("lanes", [1, 2, 5, 12])
def test_lane_count_does_not_change_results(lanes):
    """Streams are split over independent pipeline lanes (stream s -> lane s % L); the frames reported,
    their order and every byte/estimate must not depend on L (host and device inputs, chunked feeding)."""
    rng = np.random.default_rng(77)
    n_streams = 12
    caps, refs = [], []
    for s in range(n_streams):
        ms = util.MODS[(2 * s + 1) % len(util.MODS)]
        f0 = util.INNER[(s + 1) % len(util.INNER)]
        f1 = util.OUTER[(5 * s) % len(util.OUTER)]
        frames = [o.tx_frame(ms, util.CRC24, f0, f1, rng.integers(0, 256, 60 + 29 * k + s, dtype=np.uint8)) for k in range(3)]
        cap = util.build_capture(frames, rng, [400 + 101 * k for k in range(3)], snr_db=27.0,
                                 cfo=0.02 * (s / n_streams - 0.5), tau=0.4 * (s % 3 - 1), gain=0.6 + 0.1 * s,
                                 lead=64 + 131 * s, tail=900)
        caps.append(cap)
        refs.append(o.rx_capture(cap))
    rx = capi.Rx(n_streams, lanes=lanes)
    assert rx.lanes() == lanes
    # first half of every capture, then the rest for a shuffled subset order
    half = [len(c) // 2 for c in caps]
    rx.execute([c[:h] for c, h in zip(caps, half)])
    got = rx.poll()
    ids = [7, 0, 11, 3, 4, 1, 2, 10, 9, 8, 6, 5]
    rx.execute([caps[s][half[s]:] for s in ids], ids)
    second = rx.poll()
    assert [(g["stream"], g["seq"]) for g in second] == sorted((g["stream"], g["seq"]) for g in second)
    got += second
    for s in range(n_streams):
        assert_frames_match(refs[s], [g for g in got if g["stream"] == s])
    rx.reset(3)
    rx.execute([caps[3]], [3])
    assert_frames_match(refs[3], rx.poll())
    rx.close()


@pytest.mark.parametrize("lanes", [1, 3])
def test_pipelined_submit_collect_equals_execute(lanes):
    """submit()/collect() keep two calls in flight (the payload work of call k runs under the search of call
    k+1); the frames must be exactly those execute() reports, chunk by chunk, including frames that straddle
    chunk boundaries (carried samples) and a reset in between."""
    rng = np.random.default_rng(91)
    n_streams = 6
    caps, refs = [], []
    for s in range(n_streams):
        ms = util.MODS[(3 * s + 2) % len(util.MODS)]
        f0 = util.INNER[(2 * s + 1) % len(util.INNER)]
        f1 = util.OUTER[(3 * s + 1) % len(util.OUTER)]
        frames = [o.tx_frame(ms, util.CRC24, f0, f1, rng.integers(0, 256, 90 + 41 * k + 3 * s, dtype=np.uint8)) for k in range(5)]
        cap = util.build_capture(frames, rng, [350 + 77 * k for k in range(5)], snr_db=26.0,
                                 cfo=0.015 * (s / n_streams - 0.5), tau=0.3 * (s % 3 - 1), gain=0.7 + 0.1 * s,
                                 lead=200 + 53 * s, tail=1200)
        caps.append(cap)
        refs.append(o.rx_capture(cap))
    n_chunks = 7
    edges = [[(len(c) * k) // n_chunks for k in range(n_chunks + 1)] for c in caps]
    rx = capi.Rx(n_streams, lanes=lanes)
    acc = [[] for _ in range(n_streams)]

    def take():
        rx.collect()
        got = rx.poll()
        assert [(g["stream"], g["seq"]) for g in got] == sorted((g["stream"], g["seq"]) for g in got)
        for g in got:
            acc[g["stream"]].append(g)

    for k in range(n_chunks):
        rx.submit([caps[s][edges[s][k]:edges[s][k + 1]] for s in range(n_streams)])
        if k >= 1:
            take()                       # results of chunk k-1 while chunk k is still being finished
    take()
    with pytest.raises(capi.LqbError):
        rx.collect()                     # nothing left
    for s in range(n_streams):
        assert_frames_match(refs[s], acc[s])
    # three submits without a collect: the third is refused, nothing is lost
    rx.reset()
    rx.submit([c[:1000] for c in caps])
    rx.submit([c[1000:2000] for c in caps])
    with pytest.raises(capi.LqbError):
        rx.submit([c[2000:3000] for c in caps])
    rx.collect(); first = rx.poll()
    rx.collect(); second = rx.poll()
    rx.execute([c[2000:] for c in caps])
    third = rx.poll()
    for s in range(n_streams):
        assert_frames_match(refs[s], [g for g in first + second + third if g["stream"] == s])
    rx.close()


def test_gr_block_chunking_256_multiples():
    rng = np.random.default_rng(32)
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, rng.integers(0, 256, 256, dtype=np.uint8)) for _ in range(3)]
    cap = util.build_capture(frames, rng, [1024] * 3, snr_db=30.0)
    cap = cap[:len(cap) // 256 * 256]
    ref = o.rx_capture(cap, chunk=256)
    rx = capi.Rx(1)
    acc = []
    for i in range(0, len(cap), 1024):
        rx.execute([cap[i:i + 1024]])
        acc += rx.poll()
    assert_frames_match(ref, acc)


def test_noise_only_and_empty_inputs():
    rng = np.random.default_rng(33)
    x = ((rng.standard_normal(50000) + 1j * rng.standard_normal(50000)) / np.sqrt(2)).astype(np.complex64)
    rx = capi.Rx(2)
    rx.execute([x, np.zeros(0, np.complex64)])
    assert rx.poll() == []
    rx.execute([np.zeros(1000, np.complex64)], [1])
    assert rx.poll() == []
    rx.execute([], [])
    assert rx.poll() == []


def test_corrupted_header_reports_invalid_like_the_oracle():
    rng = np.random.default_rng(34)
    pl = rng.integers(0, 256, 256, dtype=np.uint8)
    x = o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl).copy()
    x[2 * 90:2 * 250] = 0          # wipe most of the header symbols
    good = o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl)
    cap = util.build_capture([x, good], rng, [1200, 900], snr_db=30.0)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    assert any(not r["header_valid"] for r in ref)
    assert_frames_match(ref, got)


def test_reset_clears_stream_state():
    rng = np.random.default_rng(35)
    pl = rng.integers(0, 256, 64, dtype=np.uint8)
    cap = util.impair(o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl), rng, snr_db=30.0, pre=500, post=700)
    rx = capi.Rx(1)
    rx.execute([cap[:1500]])       # stops mid-frame
    assert rx.poll() == []
    rx.reset()
    rx.execute([cap])
    got = rx.poll()
    assert len(got) == 1 and got[0]["payload"] == pl.tobytes() and got[0]["sample_index"] == 500


def test_per_near_threshold_agrees_with_oracle():
    # uncoded QPSK around the CRC failure knee: the two decoders must see the same frames fail
    rng = np.random.default_rng(36)
    n = 120
    pls = [rng.integers(0, 256, 128, dtype=np.uint8) for _ in range(n)]
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 1, p) for p in pls]
    cap = util.build_capture(frames, rng, [700] * n, snr_db=7.5)
    ref = o.rx_capture(cap)
    rx = capi.Rx(1)
    rx.execute([cap])
    got = rx.poll()
    ref_ok = sum(r["payload_valid"] for r in ref)
    got_ok = sum(g["payload_valid"] for g in got)
    assert 0 < ref_ok < len(ref)                     # genuinely near threshold
    per_ref = 1 - ref_ok / max(len(ref), 1)
    half = 1.96 * np.sqrt(max(per_ref * (1 - per_ref), 1e-3) / max(len(ref), 1))
    per_got = 1 - got_ok / max(len(got), 1)
    assert abs(per_got - per_ref) <= half
    assert abs(len(got) - len(ref)) <= 2


def test_viterbi_parallel_traceback_rewalk_is_exact(monkeypatch):
    # k_viterbi27x4 traces a codeword back on four lanes that start speculatively and are verified top down.
    # With the warm-up shortened to 0..23 steps and noisy codewords the speculation fails often, so the re-walk
    # path runs; bytes must still equal the oracle's serial traceback (valid and CRC-failed frames alike).
    rng = np.random.default_rng(38)
    n = 24
    pls = [rng.integers(0, 256, 400, dtype=np.uint8) for _ in range(n)]
    caps = []
    for i, (f0, f1) in enumerate([(11, 1), (11, 27), (17, 1)]):
        frames = [o.tx_frame(util.PSK4, util.CRC24, f0, f1, p) for p in pls[8 * i:8 * i + 8]]
        caps.append(util.build_capture(frames, rng, [800] * 8, snr_db=2.0 + 1.5 * i))
    ref = [o.rx_capture(c) for c in caps]
    assert any(not r["payload_valid"] for rr in ref for r in rr) and any(r["payload_valid"] for rr in ref for r in rr)
    for warm in ("0", None):
        if warm is None:
            monkeypatch.delenv("LQB_V4_WARM", raising=False)
        else:
            monkeypatch.setenv("LQB_V4_WARM", warm)
        rx = capi.Rx(len(caps))
        rx.execute(caps)
        got = rx.poll()
        for s, rr in enumerate(ref):
            assert_frames_match(rr, [g for g in got if g["stream"] == s], syms=False)


@pytest.mark.parametrize("f0,f1", [(11, 1), (11, 27), (15, 7), (1, 27), (20, 6)])
def test_tiny_payloads_match_oracle(f0, f1):
    # 1 .. 9 byte payloads with and without a check: codewords shorter than the four-lane traceback's quarters,
    # matched-filter tiles and PLL chunks with a handful of symbols, RS blocks of a few bytes
    rng = np.random.default_rng(44)
    frames, n = [], 0
    for plen in (1, 2, 3, 4, 5, 9):
        for check in (1, util.CRC24):
            frames.append(o.tx_frame(util.PSK4 if n % 2 else util.QAM16, check, f0, f1, rng.integers(0, 256, plen, dtype=np.uint8)))
            n += 1
    cap = util.build_capture(frames, rng, [640] * len(frames), snr_db=30.0, cfo=0.008, tau=0.25)
    ref = o.rx_capture(cap)
    assert len(ref) == len(frames) and all(r["payload_valid"] for r in ref)
    rx = capi.Rx(1)
    rx.execute([cap])
    assert_frames_match(ref, rx.poll())


def test_fused_pll_kernel_equals_the_two_pass_form(monkeypatch):
    # LQB_PLL_FUSED=1: tracker warp and emitter warps in one kernel (the emitters follow the tracker's progress words).
    # Every modulation class, multi-group frames of ragged lengths: bit-identical to the default two kernels in every
    # field (constellation points included), and the oracle's bytes.
    rng = np.random.default_rng(43)
    caps, refs = [], []
    for s_, ms in enumerate(util.MODS):
        frames = [o.tx_frame(ms, util.CRC24, 1, 1, rng.integers(0, 256, 300 + 411 * k + 13 * s_, dtype=np.uint8)) for k in range(3)]
        cap = util.build_capture(frames, rng, [700] * 3, snr_db=30.0, cfo=0.01 * (s_ % 3 - 1), tau=0.2)
        caps.append(cap)
        refs.append(o.rx_capture(cap))
    monkeypatch.delenv("LQB_PLL_FUSED", raising=False)
    rx = capi.Rx(len(caps))
    rx.execute(caps)
    split = rx.poll()
    monkeypatch.setenv("LQB_PLL_FUSED", "1")
    rx = capi.Rx(len(caps))
    rx.execute(caps)
    fused = rx.poll()
    assert len(split) == len(fused) == 3 * len(caps)
    for a, b in zip(split, fused):
        for k in a:
            if k == "framesyms":
                assert np.array_equal(a[k], b[k])
            else:
                assert a[k] == b[k] or (isinstance(a[k], float) and np.isnan(a[k]) and np.isnan(b[k])), k
    for s_, ref in enumerate(refs):
        assert_frames_match(ref, [g for g in fused if g["stream"] == s_], syms=False)


@pytest.mark.parametrize("snr", [5.0, 6.0])
def test_reed_solomon_corrects_and_gives_up_like_the_oracle(snr):
    # RS(255,223) alone over uncoded QPSK: at 6 dB every block has a few byte errors and all are corrected, at 5 dB
    # some blocks exceed 16 errors and are left as they are -- Berlekamp-Massey, Chien and Forney all run, and the
    # delivered bytes (valid or not) must be the oracle's
    rng = np.random.default_rng(41)
    pls = [rng.integers(0, 256, 700, dtype=np.uint8) for _ in range(6)]
    frames = [o.tx_frame(util.PSK4, util.CRC24, 1, 27, p) for p in pls]
    cap = util.build_capture(frames, rng, [800] * 6, snr_db=snr, cfo=0.01, tau=0.3)
    ref = o.rx_capture(cap)
    ok = sum(r["payload_valid"] for r in ref)
    assert ok == 6 if snr >= 6.0 else 0 < ok < 6
    rx = capi.Rx(1)
    rx.execute([cap])
    assert_frames_match(ref, rx.poll(), syms=False)


def test_large_batch_roundtrip_property():
    # size-independent property at scale: every transmitted payload comes back, per stream, in order
    rng = np.random.default_rng(37)
    n_streams, per = 64, 6
    pl = [[rng.integers(0, 256, 512, dtype=np.uint8) for _ in range(per)] for _ in range(n_streams)]
    base = {}
    caps = []
    for s in range(n_streams):
        frames = [o.tx_frame(util.PSK4, util.CRC24, 11, 27, p) for p in pl[s]]
        caps.append(util.build_capture(frames, rng, [600 + 50 * (s % 7)] * per, snr_db=9.0, cfo=0.02 * ((s % 9) / 4 - 1), tau=0.1 * (s % 5 - 2)))
    L = max(len(c) for c in caps)
    dense = np.zeros((n_streams, L), np.complex64)
    for s, c in enumerate(caps):
        dense[s, :len(c)] = c
    rx = capi.Rx(n_streams, flags=capi.RX_NO_FRAMESYMS)
    rx.execute_dense(dense)
    got = rx.poll()
    for s in range(n_streams):
        mine = [g for g in got if g["stream"] == s]
        assert [g["payload"] for g in mine] == [p.tobytes() for p in pl[s]]
        assert all(g["payload_valid"] for g in mine)
    assert rx.counts() == (n_streams * per, n_streams * per)
