"""Generate the committed golden fixtures from the CPU oracle (run from the repo root:
   python tests/golden/make_golden.py).  The oracle restates liquid-dsp (parity unpinned, see
oracle/lqo.h); these fixtures pin the oracle against itself across rounds and give the GPU
tests byte-exact expectations that do not need the oracle at all."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import lqo_py as o   # noqa: E402
import util          # noqa: E402

rng = np.random.default_rng(20261018)
out = {}
# cfg-1: PSK4, 256 B, no FEC, CRC-24 (BASELINE.json configs[0])
pl1 = rng.integers(0, 256, 256, dtype=np.uint8)
tx1 = o.tx_frame(util.PSK4, util.CRC24, 1, 1, pl1)
# cfg-3 flavour, shortened: PSK4, v27 + RS(255,223), 300 B
pl3 = rng.integers(0, 256, 300, dtype=np.uint8)
tx3 = o.tx_frame(util.PSK4, util.CRC24, 11, 27, pl3)
# cfg-5 flavour, shortened: QAM16, no FEC, 200 B
pl5 = rng.integers(0, 256, 200, dtype=np.uint8)
tx5 = o.tx_frame(util.QAM16, util.CRC24, 1, 1, pl5)
cap = util.build_capture([tx1, tx3, tx5], rng, [1024, 777, 1300], snr_db=25.0, cfo=0.013, tau=-0.3, gain=0.5)
frames = o.rx_capture(cap)
assert len(frames) == 3 and all(f["payload_valid"] for f in frames)
np.savez_compressed(
    os.path.join(HERE, "loopback_v1.npz"),
    payload1=pl1, payload3=pl3, payload5=pl5,
    tx1=tx1, tx3_head=tx3[:512], tx5_head=tx5[:512],
    tx3_len=len(tx3), tx5_len=len(tx5),
    capture=cap,
    sample_index=np.array([f["sample_index"] for f in frames], np.int64),
    stats=np.array([[f[k] for k in ("evm", "rssi", "cfo", "tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy")] for f in frames], np.float32),
    syms0_head=frames[0]["framesyms"][:64],
)
print("wrote loopback_v1.npz:", os.path.getsize(os.path.join(HERE, "loopback_v1.npz")), "bytes")
