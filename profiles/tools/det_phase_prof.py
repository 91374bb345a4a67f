"""Per-phase cycle breakdown of k_seek in detector mode (debug build with -DLQB_SEEK_PROF):
   LQB_OUT=gr-liquiddsp_b200/lib/liblqb200_prof.so bash gr-liquiddsp_b200/build.sh -DLQB_SEEK_PROF
   LQB_LIB=gr-liquiddsp_b200/lib/liblqb200_prof.so python profiles/tools/det_phase_prof.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch
from liquiddsp import capi

S, L, SP = int(os.environ.get("S", 2048)), 1 << 18, 8192
dev = torch.device("cuda", 0)
tx = capi.Tx(device=0)
g = torch.Generator(device="cpu").manual_seed(5)
pay = torch.randint(0, 256, (64, 256), dtype=torch.uint8, generator=g).to(dev)
Lf = capi.Tx.frame_len(2, 5, 1, 1, 256)
frames = torch.zeros((64, Lf), dtype=torch.complex64, device=dev)
tx.assemble_device([(2, 5, 1, 1)] * 64, [pay[i].data_ptr() for i in range(64)], [256] * 64, [frames[i].data_ptr() for i in range(64)])
flat = frames.reshape(-1)
cap = torch.empty((S, L), dtype=torch.complex64, device=dev)
gen = torch.Generator(device=dev).manual_seed(11)
n = torch.arange(L, device=dev, dtype=torch.int64)[None, :]
for s0 in range(0, S, 64):
    s1 = min(S, s0 + 64)
    sid = torch.arange(s0, s1, device=dev, dtype=torch.int64)
    k = n // SP
    h = (sid[:, None] * 1000003 + k * 7919) % 2147483647
    off = n - k * SP - h % (SP - Lf - 64)
    inside = (off >= 0) & (off < Lf)
    x = torch.where(inside, flat[((h // 7) % 64) * Lf + off.clamp(0, Lf - 1)], torch.zeros((), dtype=torch.complex64, device=dev))
    cfo = ((h // 13) % 20001).double() / 20000.0 * 0.1 - 0.05
    ph = torch.remainder(cfo * off.double(), 2.0 * torch.pi).float()
    x = x * torch.polar(torch.ones_like(ph), ph)
    snr_db = -6.0 + 0.5 * ((sid * 64) // S).double()
    nstd = torch.pow(10.0, -snr_db / 20.0).float()[:, None] / (2.0 ** 0.5)
    cap[s0:s1] = x + nstd * torch.view_as_complex(torch.randn((s1 - s0, L, 2), generator=gen, device=dev, dtype=torch.float32))
det = capi.Det(S, device=0)
Lb = capi.lib()
out = (C.c_uint64 * 16)()
for it in range(2):
    det.reset()
    det.execute_dense_ptr(cap.data_ptr(), L, L, capi.MEM_DEVICE)
    Lb.lqb_dbg_seek_prof(out, 1)
names = ["between blocks", "stage samples", "build Z", "issue MMA+prefetch", "wait MMA", "epilogue", "rowmax+decide",
         "loop misc", "exact window", "align+header", "prologue", "-"]
tot = sum(out[i] for i in range(12))
wins = det.windows()
print("k_seek(detector) %.2f ms, windows %d, detections %d" % (det.timing(), wins, len(det.poll())))
for i in range(11):
    print("%-20s %6.2f %%  %8.0f cycles/window" % (names[i], 100.0 * out[i] / tot, out[i] / max(1, wins)))
