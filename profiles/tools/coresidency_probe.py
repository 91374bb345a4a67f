"""Does a small kernel get an SM slot while k_seek holds two CTAs on every SM?  Times tiny torch kernels of different
shapes on a high-priority stream, launched right after a search has been queued and again 8 ms into it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch, bench
from liquiddsp import capi
dev = torch.device("cuda", 0)
S, N = int(os.environ.get("S", 1024)), 1 << 20
frames, _ = bench.clean_frames_ours(torch, dev, 1)
cap, _ = bench.make_capture(torch, frames, S, N, 1, dev)
rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, lanes=1)
for _ in range(2):
    rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
small = torch.zeros(256, device=dev)
big = torch.zeros(1 << 22, device=dev)
small.add_(1.0); big.add_(1.0)               # first use loads the kernel (lazy loading synchronises the device)
his = [torch.cuda.Stream(priority=-1) for _ in range(4)]      # every probe on a stream of its own, created up front
for h in his:
    with torch.cuda.stream(h): small.add_(1.0)
def probe(label, x):
    hi = his.pop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(hi):
        e0.record(); x.add_(1.0); e1.record()
    return label, e0, e1
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record()
t0 = time.perf_counter()
rx.submit_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
t1 = time.perf_counter()
pr = [probe("256 elements, right after submit", small), probe("4 Mi elements, right after submit", big)]
time.sleep(0.008)
pr += [probe("256 elements, 8 ms in", small), probe("4 Mi elements, 8 ms in", big)]
for _, _, e1 in pr: e1.synchronize()
t2 = time.perf_counter()
rx.collect(); rx.poll(raw=True)
t3 = time.perf_counter()
print("submit returned after %.2f ms; probes done at %.2f ms; collect done at %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, (t3 - t0) * 1e3))
for label, e0, e1 in pr:
    print("%-40s queued at %.2f ms, finished at %.2f ms" % (label, base.elapsed_time(e0), base.elapsed_time(e1)))
