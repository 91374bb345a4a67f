"""Bare pinned host -> device copy rate on 1 / 2 / 4 / 8 GPUs at once: the ceiling the host-buffer (e2e) leg of
bench.py can reach on this box, measured without the library.  One process, one stream per GPU, the same 1 GiB
pinned buffer per GPU copied `reps` times; per-GPU rate from CUDA events, aggregate from the host clock.

Usage (GPU box):  python profiles/tools/h2d_probe.py [--gb 1] [--reps 8] > gpurun_out/h2d_probe.json"""
import argparse
import json
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    n_dev = torch.cuda.device_count()
    nbytes = int(a.gb * (1 << 30))
    out = {"gpus_visible": n_dev, "bytes_per_copy": nbytes, "reps": a.reps, "host_cpus": os.cpu_count(), "runs": []}
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
    dst = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % d) for d in range(n_dev)]
    streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]
    for g in (1, 2, 4, 8):
        if g > n_dev:
            break
        for d in range(g):                                   # warm-up
            with torch.cuda.stream(streams[d]):
                dst[d].copy_(host[d], non_blocking=True)
        for d in range(g):
            torch.cuda.synchronize(d)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(g)]
        t0 = time.perf_counter()
        for d in range(g):
            with torch.cuda.device(d), torch.cuda.stream(streams[d]):
                ev[d][0].record(streams[d])
                for _ in range(a.reps):
                    dst[d].copy_(host[d], non_blocking=True)
                ev[d][1].record(streams[d])
        for d in range(g):
            torch.cuda.synchronize(d)
        wall = time.perf_counter() - t0
        per = [nbytes * a.reps / (ev[d][0].elapsed_time(ev[d][1]) / 1e3) / 1e9 for d in range(g)]
        out["runs"].append({"gpus": g, "aggregate_gbs": g * nbytes * a.reps / wall / 1e9, "per_gpu_gbs": [round(x, 2) for x in per],
                            "complex64_msps_equivalent": g * nbytes * a.reps / wall / 8 / 1e6})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
