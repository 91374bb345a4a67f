"""Receiver throughput for other schemes of the block API (same capture builder as bench.py, equal frame length required):
   python profiles/tools/scheme_probe.py  -> Msps and valid frames for a few (mod, fec0, fec1) triples at 1024 x 256 Ki samples."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch, bench
from liquiddsp import capi
dev = torch.device("cuda", 0)
S, N = 1024, 1 << 18
cases = [("PSK4 v27 rs8 (bench)", (2, 5, 11, 27)), ("QAM16 none none (cfg5)", (27, 5, 1, 1)), ("QAM64 v27p34 none", (29, 5, 17, 1)),
         ("PSK8 v29 rs8", (3, 5, 12, 27)), ("DPSK4 none golay", (10, 5, 1, 7)), ("PSK2 none secded7264", (1, 5, 1, 10))]
for name, props in cases:
    g = torch.Generator(device="cpu").manual_seed(3)
    pay = torch.randint(0, 256, (bench.N_DISTINCT, 1500), dtype=torch.uint8, generator=g).to(dev)
    L = capi.Tx.frame_len(*props, 1500)
    tx = capi.Tx(device=0)
    frames = torch.zeros((bench.N_DISTINCT, L), dtype=torch.complex64, device=dev)
    tx.assemble_device([props] * bench.N_DISTINCT, [pay[i].data_ptr() for i in range(bench.N_DISTINCT)], [1500] * bench.N_DISTINCT,
                       [frames[i].data_ptr() for i in range(bench.N_DISTINCT)])
    cap, sent = bench.make_capture(torch, frames, S, N, 1, dev)
    rx = capi.Rx(S, device=0, max_frame_samples=131072, flags=capi.RX_NO_FRAMESYMS)
    for _ in range(2):
        rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
    torch.cuda.synchronize(); t = time.perf_counter()
    K = 4
    for i in range(K):
        rx.submit_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
        if i: rx.collect()
    rx.collect()
    dt = (time.perf_counter() - t) / K
    fr, va = rx.counts()
    print("%-26s frame %6d samples: %7.0f Msps, %6d frames (%d valid) per call, kernel ms %s" % (name, L, S * N / dt / 1e6, fr, va, [round(x, 2) for x in rx.timing()[:5]]))
    del rx, cap
