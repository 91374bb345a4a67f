"""First-contact GPU script: FFT hook, then one-frame RX against the oracle (prints, no asserts)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401  (sets sys.path)
import numpy as np
import lqo_py as o
from liquiddsp import capi
import util

rng = np.random.default_rng(7)
L = capi.lib()
x = (rng.standard_normal(512) + 1j * rng.standard_normal(512)).astype(np.complex64)
for d in (+1, -1):
    y = np.zeros(512, np.complex64); yo = np.zeros(512, np.complex64)
    L.lqb_dbg_fft512.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rc = L.lqb_dbg_fft512(x.ctypes.data, y.ctypes.data, d)
    o.lib().lqo_fft(x.ctypes.data, yo.ctypes.data, 512, d)
    print("fft dir", d, "rc", rc, "bit-exact", np.array_equal(y.view(np.uint32), yo.view(np.uint32)), "maxerr", np.abs(y - yo).max())

pl = rng.integers(0, 256, 256, dtype=np.uint8)
fr = o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl)
cap = util.build_capture([fr, fr, fr], rng, [1024, 1500, 700], snr_db=30, cfo=0.01, tau=0.2, gain=0.7)
ref = o.rx_capture(cap)
rx = capi.Rx(1)
rx.execute([cap])
got = rx.poll()
print("timing", rx.timing(), "launches", rx.launches())
print("oracle frames", len(ref), "gpu frames", len(got))
keys = ["sample_index", "header_valid", "payload_valid", "payload_len", "num_framesyms", "mod_scheme", "fec0", "fec1",
        "evm", "rssi", "cfo", "tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy"]
for a, b in zip(ref, got):
    for k in keys:
        print("  %-14s %-22r %-22r" % (k, a[k], b[k]))
    print("  payload equal", a["payload"] == b["payload"], "== sent", b["payload"] == pl.tobytes(),
          "syms maxdiff", (np.abs(a["framesyms"] - b["framesyms"]).max() if len(a["framesyms"]) == len(b["framesyms"]) and len(a["framesyms"]) else None))
# streaming: feed the same capture in odd-sized chunks
rx2 = capi.Rx(1)
tot = []
pos = 0
for sz in [1000, 37, 2048, 5000, 123, 4096, 100000]:
    rx2.execute([cap[pos:pos + sz]]); pos += sz
    tot += rx2.poll()
    if pos >= len(cap): break
print("streamed frames", len(tot), [(f["sample_index"], f["payload_valid"]) for f in tot])
