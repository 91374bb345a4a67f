"""Shared helpers for the tests: synthetic channel and capture builders (numpy only)."""
import numpy as np

# liquid enums used throughout
PSK2, PSK4, PSK8, PSK16, DPSK2, DPSK4, DPSK8, ASK4, QAM16, QAM32, QAM64 = 1, 2, 3, 4, 9, 10, 11, 18, 27, 28, 29
MODS = [PSK2, PSK4, PSK8, PSK16, DPSK2, DPSK4, DPSK8, ASK4, QAM16, QAM32, QAM64]      # flex_tx index 0..10
INNER = [1, 11, 15, 17, 18, 19, 20]                                                   # fec0 by index 0..6
OUTER = [1, 7, 27, 4, 6, 8, 9, 10]                                                    # fec1 by index 0..7
CRC24 = 5


def impair(x, rng, snr_db=30.0, cfo=0.0, tau=0.0, gain=1.0, phi=0.0, pre=0, post=0):
    """Delay by tau samples (windowed sinc), rotate, scale, pad, add AWGN (unit-power signal convention)."""
    x = np.concatenate([np.zeros(pre, np.complex128), np.asarray(x, np.complex128), np.zeros(post, np.complex128)])
    if tau != 0.0:
        n = np.arange(-24, 25)
        h = np.sinc(n - tau) * np.hamming(49)
        x = np.convolve(x, h)[24:-24]
    n = np.arange(len(x))
    x = gain * x * np.exp(1j * (cfo * n + phi))
    nstd = gain * 10.0 ** (-snr_db / 20.0)
    x = x + nstd * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) / np.sqrt(2.0)
    return x.astype(np.complex64)


def build_capture(frames, rng, gaps, snr_db=30.0, cfo=0.0, tau=0.0, gain=1.0, lead=700, tail=900):
    """Concatenate clean frames with zero gaps, then impair the whole capture once."""
    parts = [np.zeros(lead, np.complex64)]
    for f, g in zip(frames, gaps):
        parts.append(np.asarray(f, np.complex64))
        parts.append(np.zeros(g, np.complex64))
    parts.append(np.zeros(tail, np.complex64))
    return impair(np.concatenate(parts), rng, snr_db=snr_db, cfo=cfo, tau=tau, gain=gain)
