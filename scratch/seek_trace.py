"""Per-CTA timeline of k_seek on the bench workload (needs liblqb200_trace.so built with -DLQB_SEEK_TRACE).
usage: LQB_LIB=gr-liquiddsp_b200/lib/liblqb200_trace.so python scratch/seek_trace.py [lanes]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gr-liquiddsp_b200", "python"))
import numpy as np
import torch
import bench
from liquiddsp import capi

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
S, N = 1024, 1 << 20
frames, payloads = bench.clean_frames_ours(torch, dev, 1)
cap, sent = bench.make_capture(torch, frames, S, N, 1, dev)
torch.cuda.synchronize(dev)
cs = torch.cuda.current_stream(dev)
rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=lanes)
lib = capi.lib() if hasattr(capi, "lib") else ctypes.CDLL(capi.LIB_PATH)
for _ in range(3):
    rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
capn = 65536
buf = np.zeros((4, capn), dtype=np.uint64)
n = ctypes.c_uint(0)
lib.lqb_dbg_seek_trace(None, ctypes.c_uint(0), ctypes.byref(n), 1)
rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
print("timing", rx.timing())
lib.lqb_dbg_seek_trace(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(capn), ctypes.byref(n), 1)
k = n.value
t0, t1, sm, tag = (buf[i, :k].astype(np.int64) for i in range(4))
base = t0.min()
t0 = (t0 - base) / 1e6; t1 = (t1 - base) / 1e6
d = t1 - t0
span = t1.max()
print("ctas", k, "span ms %.3f" % span, "sum d / 444 = %.3f ms" % (d.sum() / 444), "mean d %.3f min %.3f max %.3f" % (d.mean(), d.min(), d.max()))
# active CTAs over time
grid = np.linspace(0, span, 41)
act = [(int(((t0 <= g) & (t1 > g)).sum())) for g in grid]
print("active CTAs at 40 time points:", act)
# duration by start order (deciles)
o = np.argsort(t0)
print("duration by start order (deciles):", [round(float(d[o[i * k // 10:(i + 1) * k // 10]].mean()), 2) for i in range(10)])
win = (tag >> 32)
print("windows per CTA: mean %.0f min %d max %d" % (win.mean(), win.min(), win.max()))
# last-finishing: how long is the machine under half full at the end
half = [g for g, a in zip(grid, act) if a < 222 and g > span / 2]
print("time under half occupancy at the end: %.3f ms" % (span - half[0] if half else 0.0))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(lanes=lanes, ctas=int(k), span_ms=float(span), sum_over_slots_ms=float(d.sum() / 444), active=act,
               d_mean=float(d.mean()), d_min=float(d.min()), d_max=float(d.max())), open("gpurun_out/seek_trace_l%d.json" % lanes, "w"))

try:
    pr = (ctypes.c_ulonglong * 24)()
    lib.lqb_dbg_seek_prof(pr, 1)
    rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
    lib.lqb_dbg_seek_prof(pr, 1)
    tot = sum(pr)
    print("prof cycles share (thread 0):", {i: round(pr[i] / tot, 4) for i in range(24) if pr[i]}, "total Mcycles %.1f" % (tot / 1e6))
except Exception as e:
    print("no prof", e)

# ---- pipelined (submit / collect) steady state: occupancy of the search over the middle steps
import time
K = 8
lib.lqb_dbg_seek_trace(None, ctypes.c_uint(0), ctypes.byref(n), 1)
torch.cuda.synchronize(dev)
tw = time.perf_counter()
for i in range(K):
    rx.submit_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
    if i:
        rx.collect()
rx.collect()
torch.cuda.synchronize(dev)
tw = (time.perf_counter() - tw) * 1e3
lib.lqb_dbg_seek_trace(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(capn), ctypes.byref(n), 1)
k = n.value
t0, t1 = (buf[i, :k].astype(np.int64) for i in range(2))
base = t0.min()
t0 = (t0 - base) / 1e6; t1 = (t1 - base) / 1e6
span = t1.max()
lo, hi = span * 0.25, span * 0.75
grid = np.linspace(lo, hi, 2001)
s0 = np.sort(t0); s1 = np.sort(t1)
act = np.searchsorted(s0, grid, side="right") - np.searchsorted(s1, grid, side="right")
print("pipelined: %d steps wall %.2f ms (%.2f per step), ctas %d, search span %.2f" % (K, tw, tw / K, k, span))
print("middle half: mean active %.1f of 444; time share at 444: %.3f, >=400: %.3f, <222: %.3f, ==0: %.3f" % (
    act.mean(), (act >= 444).mean(), (act >= 400).mean(), (act < 222).mean(), (act == 0).mean()))
d = t1 - t0
print("CTA-time per step / 444 = %.2f ms" % (d.sum() / 444 / K))
