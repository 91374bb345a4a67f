"""A soft-decision workload for profiling: 256 streams of QAM16 + v27 frames (1500 bytes) at 9 dB through a receiver
created with LQB_RX_SOFT (device-resident capture, two steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
import torch
from liquiddsp import capi
dev = torch.device("cuda", 0)
S, N = 256, 1 << 18
g = torch.Generator().manual_seed(3)
tx = capi.Tx(device=0)
pl = [torch.randint(0, 256, (1500,), dtype=torch.uint8, generator=g).numpy() for _ in range(8)]
frames = [torch.from_numpy(f) for f in tx.assemble([(27, 5, 11, 1)] * 8, pl)]          # QAM16, CRC24, v27, none
Lf = len(frames[0])
cap = torch.zeros((S, N), dtype=torch.complex64)
pos = 500
k = 0
while pos + Lf + 600 < N:
    cap[:, pos:pos + Lf] = frames[k % 8][None, :]
    pos += Lf + 700; k += 1
nstd = 10.0 ** (-9.0 / 20.0) / 2.0 ** 0.5
cap = cap.to(dev)
cap += nstd * torch.view_as_complex(torch.randn((S, N, 2), device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
torch.cuda.synchronize()
for flags in (capi.RX_NO_FRAMESYMS, capi.RX_NO_FRAMESYMS | capi.RX_SOFT):
    rx = capi.Rx(S, device=0, max_frame_samples=32768, flags=flags, lanes=1)
    for _ in range(2):
        rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
    print("flags", flags, "frames / valid", rx.counts(), "timing ms", [round(t, 3) for t in rx.timing()])
    rx.close()
