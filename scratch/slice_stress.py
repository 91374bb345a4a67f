"""Stress of the search's time slices: many streams, many hand-overs.  usage: slice_stress.py Q mode(exec|pipe) [S] [N]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gr-liquiddsp_b200", "python"))
Q, mode = sys.argv[1], sys.argv[2]
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
N = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 18
import numpy as np
import torch
import bench
from liquiddsp import capi
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
frames, payloads = bench.clean_frames_ours(torch, dev, 1)
cap, sent = bench.make_capture(torch, frames, S, N, 1, dev)
torch.cuda.synchronize(dev)
cs = torch.cuda.current_stream(dev)
def run(q):
    os.environ["LQB_SEEK_SLICE"] = q
    rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=1)
    out = []
    if mode == "exec":
        for _ in range(3):
            rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
            a = rx.poll_array()
            out.append(np.stack([a["stream"], a["seq"], a["sample_index"], a["header_valid"], a["payload_valid"]], 1).copy())
    else:
        K = 6
        for i in range(K):
            rx.submit_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
            if i:
                rx.collect(); a = rx.poll_array()
                out.append(np.stack([a["stream"], a["seq"], a["sample_index"], a["header_valid"], a["payload_valid"]], 1).copy())
        rx.collect(); a = rx.poll_array()
        out.append(np.stack([a["stream"], a["seq"], a["sample_index"], a["header_valid"], a["payload_valid"]], 1).copy())
    rx.close()
    return out
import ctypes
def stall():
    out = (ctypes.c_uint * 8)()
    try:
        capi.lib().lqb_dbg_seek_stall(out, 1)
    except Exception as e:
        return str(e)
    return list(out)
ref = run("0")
print("unsliced ok", [len(x) for x in ref], flush=True)
try:
    got = run(Q)
finally:
    print("stall record [count, ticket, head, tail, done, n_io, bound, cta]:", stall(), flush=True)
print("sliced ok", [len(x) for x in got], flush=True)
print("identical:", all(np.array_equal(a, b) for a, b in zip(ref, got)))
