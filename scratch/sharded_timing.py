"""Where the time of lqb_rx_execute_sharded goes on the bench capture: one plain execute of the same samples as 1024
streams beside the sharded call (one stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch
import bench
from liquiddsp import capi
dev = torch.device("cuda", 0)
S, N = 1024, 1 << 20
frames, _ = bench.clean_frames_ours(torch, dev, 1)
cap, sent = bench.make_capture(torch, frames, S, N, 1, dev)
torch.cuda.synchronize()
rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS)
for _ in range(2):
    rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
t0 = time.perf_counter(); rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE); t1 = time.perf_counter()
print("plain execute, 1024 streams: %.1f ms" % (1e3 * (t1 - t0)), rx.counts())
rx.reset()
for seg, pre in ((1 << 20, 1 << 16), (1 << 20, 1 << 17), (1 << 21, 1 << 16)):
    rx.execute_sharded_ptr(cap.data_ptr(), S * N, capi.MEM_DEVICE, seg, pre)
    t0 = time.perf_counter(); rx.execute_sharded_ptr(cap.data_ptr(), S * N, capi.MEM_DEVICE, seg, pre); t1 = time.perf_counter()
    print("sharded seg %d preroll %d: %.1f ms" % (seg, pre, 1e3 * (t1 - t0)), rx.counts(), rx.shard_info())
