"""Debug: tensor-core pre-filter vs a float64 numpy reference (prints)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import numpy as np
import lqo_py as o
from liquiddsp import capi
import util

rng = np.random.default_rng(3)
L = capi.lib()
L.lqb_dbg_coarse.argtypes = [C.c_void_p, C.c_uint, C.c_float, C.c_void_p, C.c_void_p]
pl = rng.integers(0, 256, 64, dtype=np.uint8)
fr = o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl)
cap = util.build_capture([fr], rng, [600], snr_db=10, cfo=0.03, tau=0.2, gain=float(sys.argv[1]) if len(sys.argv) > 1 else 0.7, lead=300, tail=200)
n = len(cap)
nt = (n + 127) // 128
m8 = np.zeros(nt * 16, np.float32); e8 = np.zeros(nt * 16, np.float32)
rc = L.lqb_dbg_coarse(cap.ctypes.data, n, C.c_float(0.3), m8.ctypes.data, e8.ctypes.data)
print("rc", rc, "n", n, "tiles", nt)
s = capi.tab_detector_template(0.3).astype(np.complex128)
x = np.concatenate([cap.astype(np.complex128), np.zeros(512)])
nn = np.arange(156)
ref = np.zeros(nt * 128)
for b in range(49):
    t = np.conj(s * np.exp(2j * np.pi * (b - 24) * nn / 512.0))
    c = np.correlate(x, np.conj(t), mode="valid")[:nt * 128] if False else np.array([0])
# direct: C[l,b] = sum_n x[l+n] * conj(s[n] e^{j phi})
X = np.lib.stride_tricks.sliding_window_view(x, 156)[:nt * 128]          # [lags, 156]
T = np.stack([np.conj(s * np.exp(2j * np.pi * (b - 24) * nn / 512.0)) for b in range(49)], axis=1)  # [156, 49]
Cm = X @ T
ref = (np.abs(Cm) ** 2).max(axis=1).reshape(-1, 8).max(axis=1)
eref = (np.abs(x[:nt * 128]) ** 2).reshape(-1, 8).sum(axis=1)
rel = np.abs(m8 - ref) / (ref.max() + 1e-30)
print("m8 max", m8.max(), "ref max", ref.max(), "max rel err (of peak)", rel.max(), "argmax", m8.argmax(), ref.argmax())
print("per-block rel err median", np.median(np.abs(m8 - ref) / (ref + 1e-30)))
print("e8 max rel err", (np.abs(e8 - eref) / (eref + 1e-30)).max())
print(m8[:8], ref[:8])
