"""Debug helper: per-frame share of constellation points that differ from the oracle's by more than the test tolerance."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + "/gr-liquiddsp_b200/python"); sys.path.insert(0, R + "/oracle"); sys.path.insert(0, R + "/tests")
import numpy as np, lqo_py as o, util
from liquiddsp import capi
rng = np.random.default_rng(43)
caps, refs = [], []
for s_, ms in enumerate(util.MODS):
    frames = [o.tx_frame(ms, util.CRC24, 1, 1, rng.integers(0, 256, 300 + 411 * k + 13 * s_, dtype=np.uint8)) for k in range(3)]
    caps.append(util.build_capture(frames, rng, [700] * 3, snr_db=30.0, cfo=0.01 * (s_ % 3 - 1), tau=0.2))
    refs.append(o.rx_capture(caps[-1]))
rx = capi.Rx(len(caps)); rx.execute(caps); got = rx.poll()
for s_, ref in enumerate(refs):
    for r, g in zip(ref, [x for x in got if x["stream"] == s_]):
        err = np.abs(g["framesyms"] - r["framesyms"]); mag = np.abs(r["framesyms"])
        bad = err > 1e-4 + 1e-3 * mag
        first = int(np.argmax(bad)) if bad.any() else -1
        print("ms %2d n_sym %5d  share beyond tol %.4f  max err/mag %.5f  first bad %d  evm %.4f/%.4f" % (r["mod_scheme"], len(err), bad.mean(), float((err / np.maximum(mag, 1e-9)).max()), first, g["evm"], r["evm"]))
