"""Throughput of the frame generator (k_tx) on device buffers: frames per launch x samples per frame / time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch
from liquiddsp import capi
dev = torch.device("cuda", 0)
tx = capi.Tx(device=0)
for name, props, n in (("cfg5 QAM16 1500 B no FEC", (27, 5, 1, 1), 8192), ("cfg3 PSK4 1500 B v27+rs8", (2, 5, 11, 27), 4096)):
    L = capi.Tx.frame_len(*props, 1500)
    pl = torch.randint(0, 256, (n, 1504), dtype=torch.uint8, device=dev)
    out = torch.empty((n, L), dtype=torch.complex64, device=dev)
    import ctypes as C
    ms, c, f0, f1 = props
    P = (capi.TxProps * n)(*[capi.TxProps(c, f0, f1, ms) for _ in range(n)])      # built once: the C call is what is timed
    lens = (C.c_uint32 * n)(*([1500] * n))
    pp = (C.c_void_p * n)(*[pl[i].data_ptr() for i in range(n)])
    op = (C.c_void_p * n)(*[out[i].data_ptr() for i in range(n)])
    call = lambda: capi._check(tx._L.lqb_tx_assemble(tx._h, n, P, None, pp, lens, op, capi.MEM_DEVICE))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    t = time.perf_counter()
    K = 5
    for _ in range(K):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / K
    print("%-28s %5d frames x %6d samples: %.3f ms per launch = %.1f Gsps = %.0f GB/s written" % (name, n, L, dt * 1e3, n * L / dt / 1e9, n * L * 8 / dt / 1e9))
