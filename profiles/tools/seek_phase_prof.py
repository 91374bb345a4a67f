"""Per-phase cycle breakdown of k_seek (debug build with -DLQB_SEEK_PROF, see gr-liquiddsp_b200/build.sh).
Usage (GPU box):  LQB_OUT=gr-liquiddsp_b200/lib/liblqb200_prof.so bash gr-liquiddsp_b200/build.sh -DLQB_SEEK_PROF
                  LQB_LIB=gr-liquiddsp_b200/lib/liblqb200_prof.so python profiles/tools/seek_phase_prof.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch
import bench
from liquiddsp import capi

S = int(os.environ.get("S", 1024)); N = int(os.environ.get("N", 1 << 20))
dev = torch.device("cuda", 0)
frames, _ = bench.clean_frames_ours(torch, dev, 1)
cap, sent = bench.make_capture(torch, frames, S, N, 1, dev)
rx = capi.Rx(S, device=0, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, lanes=1)
L = capi.lib()
out = (C.c_uint64 * 24)()
for it in range(2):
    rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
    L.lqb_dbg_seek_prof(out, 1)
names = ["between blocks", "quantise + sums", "build Z", "prefetch issue", "wait MMA", "epilogue", "rowmax+decide",
         "loop misc", "exact: eval_window", "frame: descriptor + state", "prologue",
         "exact: bin_candidates", "exact: restore tables + load window", "frame: align tail (phase sum, atan2)", "frame: header decode (deinterleave, Hamming, SECDED, CRC)",
         "exact: energy + forward FFT", "exact: inverse FFTs of the candidate bins", "frame: load + forward FFT + CFO product",
         "frame: cross IFFT + CFO FFT", "frame: tau/gamma/dphi + 156 sincos", "frame: header mix-down + matched filter",
         "frame: pilot sync + QPSK slicing", "-", "-"]
tot = sum(out[i] for i in range(24))
w = rx.work(); t = rx.timing()
print("seek %.2f ms, windows %d, tiles %d, exact %d, aligns %d" % (t[0], w["windows"], w["coarse_tiles"], w["exact_windows"], w["aligns"]))
for i in range(22):
    ev = w["exact_windows"] if names[i].startswith("exact") else w["aligns"] if names[i].startswith("frame") else 0
    print("%-62s %6.2f %%  %8.0f cycles/window" % (names[i], 100.0 * out[i] / tot, out[i] / max(1, w["windows"])) + ("  %8.0f cycles/event" % (out[i] / ev) if ev else ""))
