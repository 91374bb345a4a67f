"""The oracle against the committed golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np

import lqo_py as o
import util

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loopback_v1.npz"))


def test_tx_oracle_reproduces_golden_samples():
    tx1 = o.tx_frame(util.PSK4, util.CRC24, 1, 1, G["payload1"])
    assert np.array_equal(tx1.view(np.uint32), G["tx1"].view(np.uint32))
    tx3 = o.tx_frame(util.PSK4, util.CRC24, 11, 27, G["payload3"])
    assert len(tx3) == int(G["tx3_len"]) and np.array_equal(tx3[:512], G["tx3_head"])
    tx5 = o.tx_frame(util.QAM16, util.CRC24, 1, 1, G["payload5"])
    assert len(tx5) == int(G["tx5_len"]) and np.array_equal(tx5[:512], G["tx5_head"])


def test_rx_oracle_reproduces_golden_decode():
    fr = o.rx_capture(G["capture"])
    assert [f["sample_index"] for f in fr] == G["sample_index"].tolist()
    for f, pl in zip(fr, (G["payload1"], G["payload3"], G["payload5"])):
        assert f["header_valid"] and f["payload_valid"] and f["payload"] == pl.tobytes()
    keys = ("evm", "rssi", "cfo", "tau_hat", "gamma_hat", "dphi_hat", "phi_hat", "rxy")
    got = np.array([[f[k] for k in keys] for f in fr], np.float32)
    assert np.allclose(got, G["stats"], rtol=1e-5, atol=1e-6)
    assert np.allclose(fr[0]["framesyms"][:64], G["syms0_head"], rtol=1e-5, atol=1e-6)
