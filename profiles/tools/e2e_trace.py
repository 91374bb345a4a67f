"""Host-buffer (e2e) timeline of the pipelined receiver next to the raw PCIe copy rates (LQB_TRACE=1 prints the marks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch, bench
from liquiddsp import capi
dev = torch.device("cuda", 0)
S, N = 1024, 1 << 18
frames, _ = bench.clean_frames_ours(torch, dev, 1)
cap, _ = bench.make_capture(torch, frames, S, N, 1, dev)
host = torch.empty((S, N), dtype=torch.complex64).pin_memory(); host.copy_(cap)
d = torch.empty_like(cap)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize()
    print("H2D %.1f GB/s" % (host.numel() * 8 / (time.perf_counter() - t) / 1e9))
h2 = torch.empty((S, N // 4), dtype=torch.complex64).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); h2.copy_(cap[:, :N // 4], non_blocking=True); torch.cuda.synchronize()
    print("D2H %.1f GB/s" % (h2.numel() * 8 / (time.perf_counter() - t) / 1e9))
L = int(os.environ.get("L", 4))
rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=0, lanes=L)
for _ in range(3):
    rx.execute_dense_ptr(host.data_ptr(), N, N, capi.MEM_HOST)
torch.cuda.synchronize(); t = time.perf_counter()
K = 4
for i in range(K):
    rx.submit_dense_ptr(host.data_ptr(), N, N, capi.MEM_HOST)
    if i: rx.collect(); rx.poll(raw=True)
rx.collect(); rx.poll(raw=True)
dt = time.perf_counter() - t
print("e2e %.0f Msps, %.1f ms/step" % (K * S * N / dt / 1e6, dt / K * 1e3))
