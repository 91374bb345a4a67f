"""Time slices of the search (k_seek hands a stream from CTA to CTA inside one launch): results must not change.
Reference behaviour: flexframesync / qdetector are sample-serial per stream (liquid-dsp flexframesync_execute,
qdetector_cccf_execute); how the samples of a call are cut up between CTAs is invisible to the caller."""
import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import capi
from test_gpu_parity import assert_frames_match

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("gpu_required")]


def _streams(rng, n_streams, n_frames=3, snr_db=27.0):
    caps = []
    for s in range(n_streams):
        ms = util.MODS[(2 * s + 1) % len(util.MODS)]
        f0 = util.INNER[(s + 1) % len(util.INNER)]
        f1 = util.OUTER[(5 * s) % len(util.OUTER)]
        frames = [o.tx_frame(ms, util.CRC24, f0, f1, rng.integers(0, 256, 60 + 29 * k + s, dtype=np.uint8)) for k in range(n_frames)]
        caps.append(util.build_capture(frames, rng, [400 + 101 * k for k in range(n_frames)], snr_db=snr_db,
                                       cfo=0.02 * (s / n_streams - 0.5), tau=0.4 * (s % 3 - 1), gain=0.6 + 0.1 * s,
                                       lead=64 + 131 * s, tail=900))
    return caps


@pytest.mark.parametrize("slice_len", [256, 1024, 4096])
def test_sliced_search_matches_oracle(monkeypatch, slice_len):
    """LQB_SEEK_SLICE given explicitly applies to any number of streams: every stream is walked by a chain of CTAs
    (dozens at 256 samples per slice); frames, bytes and estimates must equal the oracle's, also across calls."""
    monkeypatch.setenv("LQB_SEEK_SLICE", str(slice_len))
    rng = np.random.default_rng(177)
    caps = _streams(rng, 10)
    refs = [o.rx_capture(c) for c in caps]
    rx = capi.Rx(len(caps), lanes=1)
    half = [len(c) // 2 for c in caps]
    rx.execute([c[:h] for c, h in zip(caps, half)])
    got = rx.poll()
    rx.execute([c[h:] for c, h in zip(caps, half)])
    got += rx.poll()
    for s in range(len(caps)):
        assert_frames_match(refs[s], [g for g in got if g["stream"] == s])
    rx.close()


def test_sliced_detector_matches_unsliced(monkeypatch):
    rng = np.random.default_rng(178)
    caps = _streams(rng, 6, n_frames=4, snr_db=12.0)
    out = {}
    for conf in ("0", "512"):
        monkeypatch.setenv("LQB_SEEK_SLICE", conf)
        det = capi.Det(len(caps))
        det.execute(caps)
        out[conf] = sorted(tuple(d.values()) for d in det.poll())
        det.close()
    assert len(out["0"]) >= 20
    assert out["0"] == out["512"]


def test_more_streams_than_cta_slots_sliced_equals_unsliced(monkeypatch):
    """More streams than the GPU holds search CTAs (3 per SM = 444): the launch is 444 CTAs drawing slices of 460 streams
    from the queue.  Sliced against unsliced (the default): identical frames in identical order, every field."""
    rng = np.random.default_rng(179)
    base = _streams(rng, 8, n_frames=6)
    n = 160 * 1024
    S = 460
    cap = np.zeros((S, n), dtype=np.complex64)
    for s in range(S):
        b = base[s % len(base)]
        off = 1000 + 37 * s
        reps = (n - off) // (len(b) + 500)
        for r in range(reps):
            a = off + r * (len(b) + 500)
            cap[s, a:a + len(b)] = b
    res = {}
    for conf in ("8192", None):
        if conf is None:
            monkeypatch.delenv("LQB_SEEK_SLICE", raising=False)
        else:
            monkeypatch.setenv("LQB_SEEK_SLICE", conf)
        rx = capi.Rx(S, lanes=1, max_frame_samples=16384)
        l0 = rx.launches()
        rx.execute([cap[s] for s in range(S)])
        res[conf] = (rx.poll(), rx.launches() - l0)
        rx.close()
    (a, la), (b, lb) = res["8192"], res[None]
    assert la == lb + 1                     # the queue set-up kernel: slicing was on with the variable and off without
    assert len(a) == len(b) and len(a) > 5 * S
    for x, y in zip(a, b):
        for k, xv in x.items():
            yv = y[k]
            assert (np.array_equal(xv, yv) if isinstance(xv, np.ndarray) else xv == yv), (k, x["stream"], x["seq"])
