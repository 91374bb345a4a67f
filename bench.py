#!/usr/bin/env python
"""bench.py -- flex_rx throughput on synthetic multi-channel IQ (BASELINE.json metric).

Workload (config.workload = "flex_rx_1024ch_qpsk_v27_rs8_1500B", BASELINE.json configs[2], the
configuration the metric "flex_rx Msps & decoded frames/s" is quoted on and the largest that is
a single-GPU flex_rx case): per GPU 1024 independent channel streams x 1,048,576 samples per
step; PSK4 + inner v27 + outer RS(255,223), 1500-byte payloads, CRC-24; ~85 % duty; per-stream
SNR sweep -2..+12 dB in 1 dB steps, CFO U(+-0.02) rad/sample, timing offset U(+-0.5) sample,
gain U(0.5,1.5).  A step = one pass of the whole receive path over that batch.

  value  : complex input samples consumed per second (Msps), whole job, inputs resident in HBM
  e2e    : same metric through the reference-facing C-ABI call with HOST buffers (pinned input
           copied H2D inside the timed region, frame results + payload bytes + constellation
           points copied D2H inside the timed region)
  --impl reference : the CPU restatement of liquid-dsp's flexframesync (oracle/) on the host cores

Multi-GPU: channels shard across ranks (weak scaling, 1024 channels per GPU), no data-path
collective; torch.distributed is used only for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))
# keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line (NCCL carries no data here, only barriers).
# NCCL prints the banner at the VERSION and at the WARN level; with the variable unset it prints nothing.
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]


def bind_near_gpu(local):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned staging buffers of
    the host-buffer (e2e) leg are first-touched on that NUMA node.  Returns the CPU count it bound to (0 = left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        allowed = os.sched_getaffinity(0)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0

PSK4, CRC24, V27, RS8 = 2, 5, 11, 27
PAYLOAD = 1500
N_DISTINCT = 64
N_TAU = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="channels per GPU")
    ap.add_argument("--samples", type=int, default=1 << 20, help="samples per channel per step")
    ap.add_argument("--e2e-samples", type=int, default=0, help="samples per channel per e2e step (0 = the full step size)")
    ap.add_argument("--workload", default="flex_rx", choices=["flex_rx", "detector", "tx", "mixed_mod", "tx_rx_per"],
                    help="flex_rx = configs[2] (default, the headline metric; short runs of the other configurations ride along "
                         "under \"workloads\"); detector = configs[1] bulk frame_detector_cc; mixed_mod = configs[3] (4096 channels, "
                         "per-frame schemes drawn by the policy stand-in, closed loop); tx_rx_per = configs[4] (8192 channels, frames "
                         "generated on the GPU, PER vs SNR beside the oracle); tx = the flexframegen batch alone")
    ap.add_argument("--no-workloads", action="store_true", help="flex_rx only: skip the short runs of configs 2, 4 and 5")
    ap.add_argument("--lanes", type=int, default=0, help="pipeline lanes per receiver handle (0 = library default)")
    ap.add_argument("--e2e-lanes", type=int, default=0, help="lanes of the host-buffer (e2e) receiver (0 = library default)")
    ap.add_argument("--no-pipeline", action="store_true", help="use lqb_rx_execute per step instead of submit/collect")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def flex_rx_config(args):
    """The workload both arms declare (BASELINE.json configs[2]).  The reference arm times a bounded SAMPLE of it and
    says which under cpu_baseline.sample / "sample" -- the configuration itself is the same dict in both lines."""
    return {"workload": "flex_rx_1024ch_qpsk_v27_rs8_1500B", "streams_per_gpu": args.streams, "samples_per_stream_per_step": args.samples,
            "payload_bytes": PAYLOAD, "mod": "PSK4", "fec0": "v27", "fec1": "rs8", "check": "crc24",
            "snr_db": "-2..+12 per stream", "l2": "inputs (%.1f GB per step) larger than L2" % (args.streams * args.samples * 8 / 1e9)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------- synthetic capture
def clean_frames_ours(torch, dev, seed):
    """64 distinct cfg-3 frames from the product's GPU frame generator; returns (frames[64, L], payloads)."""
    from liquiddsp import capi
    g = torch.Generator(device="cpu").manual_seed(seed)
    payloads = torch.randint(0, 256, (N_DISTINCT, PAYLOAD), dtype=torch.uint8, generator=g)
    L = capi.Tx.frame_len(PSK4, CRC24, V27, RS8, PAYLOAD)
    tx = capi.Tx(device=dev.index, cuda_stream=torch.cuda.current_stream(dev).cuda_stream)
    d_pay = payloads.to(dev)
    frames = torch.zeros((N_DISTINCT, L), dtype=torch.complex64, device=dev)
    tx.assemble_device([(PSK4, CRC24, V27, RS8)] * N_DISTINCT,
                       [d_pay[i].data_ptr() for i in range(N_DISTINCT)], [PAYLOAD] * N_DISTINCT,
                       [frames[i].data_ptr() for i in range(N_DISTINCT)])
    torch.cuda.synchronize(dev)
    tx.close()
    return frames, payloads


def clean_frames_oracle(torch, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lqo_py as o
    g = torch.Generator(device="cpu").manual_seed(seed)
    payloads = torch.randint(0, 256, (N_DISTINCT, PAYLOAD), dtype=torch.uint8, generator=g)
    fr = [torch.from_numpy(o.tx_frame(PSK4, CRC24, V27, RS8, payloads[i].numpy())) for i in range(N_DISTINCT)]
    return torch.stack(fr), payloads


def make_capture(torch, frames, S, N, seed, dev, stream_offset=0):
    """[S, N] complex64 capture on `dev`: frames at a per-stream period, delayed, rotated, scaled, noisy."""
    L = frames.shape[1]
    Lp = L + 64
    # fractional delays: N_TAU classes, applied to every distinct frame in the frequency domain
    pad = torch.zeros((N_DISTINCT, Lp), dtype=torch.complex64, device=dev)
    pad[:, 16:16 + L] = frames.to(dev)
    F = torch.fft.fft(pad, dim=1)
    f = torch.fft.fftfreq(Lp, device=dev)
    taus = torch.linspace(-0.5, 0.5, N_TAU, device=dev)
    variants = torch.fft.ifft(F[None, :, :] * torch.exp(-2j * torch.pi * f[None, None, :] * taus[:, None, None]), dim=2)
    variants = variants.to(torch.complex64).reshape(-1)            # [N_TAU * N_DISTINCT * Lp]
    g = torch.Generator(device=dev).manual_seed(seed + 1000 * stream_offset + 17)
    out = torch.empty((S, N), dtype=torch.complex64, device=dev)
    sent = 0
    CH = 32
    n = torch.arange(N, device=dev, dtype=torch.int64)[None, :]
    for s0 in range(0, S, CH):
        s1 = min(S, s0 + CH)
        sid = torch.arange(s0, s1, device=dev, dtype=torch.int64) + stream_offset
        gap = 3000 + (sid * 7919) % 4000                            # 3000..6999, mean ~5000 -> ~85 % duty
        period = Lp + gap
        lead = (sid * 104729) % period
        nfr = (N - lead) // period                                  # whole frames only; the tail is noise
        k = (n - lead[:, None]) // period[:, None]
        off = (n - lead[:, None]) - k * period[:, None]
        inside = (n >= lead[:, None]) & (off < Lp) & (k < nfr[:, None])
        tau_c = (sid % N_TAU)[:, None]
        which = ((sid[:, None] * 31 + k * 7) % N_DISTINCT)
        idx = ((tau_c * N_DISTINCT + which) * Lp + off).clamp_(0, variants.numel() - 1)
        x = torch.where(inside, variants[idx], torch.zeros((), dtype=torch.complex64, device=dev))
        del k, off, idx, inside
        cfo = (((sid * 2654435761) % 10007).double() / 10007.0 - 0.5) * 0.04          # U(-0.02, 0.02)
        gain = 0.5 + ((sid * 40503) % 1009).double() / 1009.0
        snr_db = -2.0 + (sid % 15).double()
        ph = torch.remainder(cfo[:, None] * n.double(), 2.0 * torch.pi).float()
        x = x * torch.polar(gain.float()[:, None].expand_as(ph).contiguous(), ph)
        nstd = (gain * torch.pow(10.0, -snr_db / 20.0)).float()[:, None] / (2.0 ** 0.5)
        noise = torch.view_as_complex(torch.randn((s1 - s0, N, 2), generator=g, device=dev, dtype=torch.float32))
        out[s0:s1] = x + nstd * noise
        sent += int(nfr.sum().item())
        del x, ph, noise
    return out, sent


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [q.strip() for q in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU legs (oracle = the checker, never shipped)
def cpu_rx(capture_host, threads):
    """Times the oracle flexframesync (oracle/) over [S, N] on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lqo_py as o
    o.lib()
    t0 = time.perf_counter()
    frames, valid = o.rx_many(capture_host, threads)
    dt = time.perf_counter() - t0
    return dt, frames, valid


def oracle_frames(capture_host, threads):
    """Per-frame records of the oracle for every row of capture_host (each row a fresh flexframesync), rows in parallel
    on `threads` host threads (ctypes releases the GIL).  Returns a list (one per row) of lists of dicts."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lqo_py as o
    from concurrent.futures import ThreadPoolExecutor
    o.lib()
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(lambda r: o.rx_capture(capture_host[r]), range(capture_host.shape[0])))


def reference_arm(args, rank, world):
    """--impl reference: liquid-dsp's algorithm (the oracle port) on all host cores, a bounded sample of the same workload."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    frames, _ = clean_frames_oracle(torch, 1)
    S = min(args.streams, 4 * cores)
    N = min(args.samples, 1 << 18)
    cap, sent = make_capture(torch, frames, S, N, 1, torch.device("cpu"))
    x = cap.numpy()
    times, fr, va = [], 0, 0
    for i in range(args.warmup + args.steps):
        dt, f, v = cpu_rx(x, cores)
        if i >= args.warmup:
            times.append(dt); fr += f; va += v
    T = sum(times)
    msps = args.steps * S * N / T / 1e6
    sample = "%d streams x %d samples per step on %d threads (oracle/ port of liquid-dsp flexframesync, own radix-2 FFT)" % (S, N, cores)
    print(json.dumps({
        "impl": "reference", "metric": "flex_rx_msps", "value": msps, "unit": "Msps", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": flex_rx_config(args), "sample": sample,
        "decoded_frames_per_s": va / T, "frames_per_s": fr / T,
        "cpu_baseline": {"value": msps, "unit": "Msps", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": "Msps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ----------------------------------------------------------------------------- configs[1]: bulk frame_detector_cc
def detector_arm(args, rank, local, world, steps=None, warmup=None, emit=True):
    """frame_detector_cc over a long synthetic capture sharded in time: per step S segments x L samples
    (default 4096 x 262144 = 1.07 Gsample, i.e. a 10 Gsample capture is ten steps), cfg-1 frames at jittered
    8192-sample spacing, per-frame CFO U(+-0.05) rad/sample, 64 SNR points -6..+25.5 dB (one per 1/64 of the
    segments), detector beta 0.3 / threshold 0.45.  Reports Msps, detections/s, Pd per SNR and the oracle's
    agreement on a subsample."""
    import torch
    import torch.distributed as dist
    from liquiddsp import capi
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    standalone = emit
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    S, L = 4096, 1 << 18
    if args.streams != 1024:
        S = args.streams
    cs = torch.cuda.current_stream(dev)
    # one clean cfg-1 frame set from the GPU frame generator
    g = torch.Generator(device="cpu").manual_seed(5)
    payloads = torch.randint(0, 256, (N_DISTINCT, 256), dtype=torch.uint8, generator=g)
    Lf = capi.Tx.frame_len(PSK4, CRC24, 1, 1, 256)
    tx = capi.Tx(device=local, cuda_stream=cs.cuda_stream)
    d_pay = payloads.to(dev)
    frames = torch.zeros((N_DISTINCT, Lf), dtype=torch.complex64, device=dev)
    tx.assemble_device([(PSK4, CRC24, 1, 1)] * N_DISTINCT, [d_pay[i].data_ptr() for i in range(N_DISTINCT)], [256] * N_DISTINCT,
                       [frames[i].data_ptr() for i in range(N_DISTINCT)])
    torch.cuda.synchronize(dev)
    flat = frames.reshape(-1)
    SP = 8192
    cap = torch.empty((S, L), dtype=torch.complex64, device=dev)
    gen = torch.Generator(device=dev).manual_seed(11 + rank)
    n = torch.arange(L, device=dev, dtype=torch.int64)[None, :]
    starts_all = []
    for s0 in range(0, S, 64):
        s1 = min(S, s0 + 64)
        sid = torch.arange(s0, s1, device=dev, dtype=torch.int64) + rank * S
        k = n // SP
        h = (sid[:, None] * 1000003 + k * 7919) % 2147483647
        jitter = h % (SP - Lf - 64)
        off = n - k * SP - jitter
        inside = (off >= 0) & (off < Lf)
        which = (h // 7) % N_DISTINCT
        cfo = ((h // 13) % 20001).double() / 20000.0 * 0.1 - 0.05
        ph0 = ((h // 17) % 6283).double() / 1000.0
        x = torch.where(inside, flat[(which * Lf + off.clamp(0, Lf - 1))], torch.zeros((), dtype=torch.complex64, device=dev))
        ph = torch.remainder(cfo * off.double() + ph0, 2.0 * torch.pi).float()
        x = x * torch.polar(torch.ones_like(ph), ph)
        snr_db = -6.0 + 0.5 * ((sid * 64) // (S * world)).double()
        nstd = torch.pow(10.0, -snr_db / 20.0).float()[:, None] / (2.0 ** 0.5)
        cap[s0:s1] = x + nstd * torch.view_as_complex(torch.randn((s1 - s0, L, 2), generator=gen, device=dev, dtype=torch.float32))
        kk = torch.arange(L // SP, device=dev, dtype=torch.int64)[None, :]
        hh = (sid[:, None] * 1000003 + kk * 7919) % 2147483647
        starts_all.append((kk * SP + hh % (SP - Lf - 64)).cpu())
        del x, ph, off, inside, k, h
    starts = torch.cat(starts_all).numpy()                       # [S, L/SP] true frame starts
    torch.cuda.synchronize(dev)
    det = capi.Det(S, device=local, cuda_stream=cs.cuda_stream)

    def step():
        det.reset()
        det.execute_dense_ptr(cap.data_ptr(), L, L, capi.MEM_DEVICE)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms, nd, wins = 0.0, 0, 0
    e0.record(cs)
    for _ in range(steps):
        step()
        kms += det.timing()
        wins += det.windows()
        srch = det.search()
    e1.record(cs)
    torch.cuda.synchronize(dev)
    found = det.poll()
    nd = len(found)
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    # ---- the same samples as ONE capture (row after row), searched by the time-sharded detector: the sequential
    # qdetector's list, whatever the cut (lqb_det_execute_sharded); checked below against the oracle on a prefix
    one = None
    if not os.environ.get("LQB_BENCH_NO_ONE_CAPTURE"):
        seg_len = int(os.environ.get("LQB_BENCH_SEG_LEN", 1 << 18))
        preroll = int(os.environ.get("LQB_BENCH_PREROLL", 1 << 14))
        n_seg = (S * L + seg_len - 1) // seg_len
        det1 = capi.Det(min(n_seg, 4096), device=local, cuda_stream=cs.cuda_stream)
        det1.execute_sharded_ptr(cap.data_ptr(), S * L, capi.MEM_DEVICE, seg_len, preroll)      # warm-up (buffers)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        det1.execute_sharded_ptr(cap.data_ptr(), S * L, capi.MEM_DEVICE, seg_len, preroll)
        torch.cuda.synchronize(dev)
        one_s = time.perf_counter() - t0
        one_found = det1.poll()
        one = {"samples": S * L, "seg_len": seg_len, "preroll": preroll, "ms": one_s * 1e3, "msps": S * L / one_s / 1e6,
               "gpu_ms": det1.timing(), "detections": len(one_found), "windows": det1.windows()}
        one.update(det1.shard_info())
        det1.close()
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    det.close()
    cap_sub = None
    if rank == 0 and not args.no_cpu_baseline:
        sub = list(range(0, S, max(1, S // 16)))[:16]
        cap_sub = {sidx: cap[sidx].cpu().numpy() for sidx in sub}
        if one is not None:
            n_pref = min(S * L, 3 << 20)
            cap_prefix = cap.reshape(-1)[:n_pref].cpu().numpy()
    del cap
    torch.cuda.empty_cache()
    if rank != 0:
        if world > 1 and standalone:
            dist.destroy_process_group()
        return None
    import numpy as np
    secs = float(tt[0]) / 1e3
    value = world * S * L * steps / secs / 1e6
    # detection probability per SNR point (a true start matched within +-2 samples)
    by_stream = {}
    for d in found:
        by_stream.setdefault(d["stream"], []).append(d["sample_index"])
    hit = np.zeros(64); tot = np.zeros(64); extra = 0
    for sidx in range(S):
        p = (sidx * 64) // (S * world)
        got = np.array(sorted(by_stream.get(sidx, [])), dtype=np.int64)
        tr = starts[sidx]
        tr = tr[tr + Lf + 512 <= L]
        tot[p] += len(tr)
        if len(got):
            dmin = np.abs(got[None, :] - tr[:, None]).min(axis=1) if len(tr) else np.zeros(0)
            hit[p] += int((dmin <= 2).sum())
            extra += int((np.abs(got[:, None] - starts[sidx][None, :]).min(axis=1) > 2).sum())
    pd = [round(float(h / t), 4) if t else None for h, t in zip(hit, tot)]
    # oracle agreement on a subsample of segments
    agree = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lqo_py as o
        same = cnt = 0
        t0 = time.perf_counter()
        sub = sorted(cap_sub)
        for sidx in sub:
            ref = [int(np.int64(np.uint64(r["sample_index"]))) for r in o.detect_capture(cap_sub[sidx], 0.3, 0.45)]
            mine = sorted(by_stream.get(sidx, []))
            cnt += 1
            same += int(ref == mine)
        cpu_s = time.perf_counter() - t0
        agree = {"segments": cnt, "identical_detection_lists": same, "cpu_msps_1_thread": len(sub) * L / cpu_s / 1e6}
        if one is not None:
            # the sequential oracle over the first samples of the one-capture run: same list up to where its input ends
            ref = [int(np.int64(np.uint64(r["sample_index"]))) for r in o.detect_capture(cap_prefix, 0.3, 0.45)]
            lim = len(cap_prefix) - 2048
            mine = [d["sample_index"] for d in one_found if d["sample_index"] < lim]
            ref = [r for r in ref if r < lim]
            one["oracle_prefix"] = {"samples": len(cap_prefix), "detections": len(ref), "identical": bool(ref == mine)}
    peaks = load_peaks()
    tc_peak = 2.0 * float(peaks.get("bf16_tflops_sustained", 1400.0))   # fp8 operands: twice the measured bf16 rate (see the flex_rx arm)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    t_k = kms / 1e3
    tiles = 2.0 * wins                                            # two 128-lag tiles per hop in steady state
    tflops = tiles * 128 * 156 * 49 * 8 / t_k / 1e12 if t_k else None  # algorithmic (SURVEY.md 8d dense form)
    out = {
        "metric": "frame_detector_msps", "value": value, "unit": "Msps", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": float(tt[0]) / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "e4m3 x e4m3 -> f32 pre-filter, f32 exact",
        "data": "synthetic",
        "config": {"workload": "frame_detector_cc_bulk", "segments_per_gpu": S, "samples_per_segment": L, "overlap": 1024,
                   "frame_spacing": SP, "cfo": "+-0.05 rad/sample per frame", "snr_points": 64, "l2": "inputs (%.1f GB) larger than L2" % (S * L * 8 / 1e9)},
        "detections_per_s": nd * world * steps / secs if secs else None, "detections_per_step": nd,
        "spurious_detections_per_step": extra, "search_work_last_step": srch, "pd_by_snr_point": pd, "snr_db_points": [-6.0 + 0.5 * i for i in range(64)],
        "clocks": clk, "gpu_launches": 2 * steps,
        "roofline": {"bound": "tensor", "kernel": "k_seek(detector)", "achieved": tflops, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": tflops / tc_peak if tflops else None, "traffic": None,
                     "hbm_gbs": 8.0 * S * L * steps / t_k / 1e9 if t_k else None,
                     "hbm_frac": 8.0 * S * L * steps / t_k / 1e9 / hbm_peak if t_k else None},
        "oracle_agreement": agree,
        "one_capture_time_sharded": one,
    }
    if emit:
        print(json.dumps(out))
    if world > 1 and standalone:
        dist.destroy_process_group()
    return out


def tx_arm(args, rank, local, world):
    """flex_tx / flexframegen batch (lqb_tx_assemble on device buffers): per step and GPU 8192 frames of BASELINE
    config 5's format (QAM16, 1500 B, no FEC, CRC-24: 6630 samples each).  Reports generated Msps, frames/s, the
    HBM-write roofline of k_tx and the oracle's frame generator on one host core beside it."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from liquiddsp import capi
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    QAM16 = 27
    n = 8192 if args.streams == 1024 else args.streams
    props = (QAM16, CRC24, 1, 1)
    L = capi.Tx.frame_len(*props, PAYLOAD)
    cs = torch.cuda.current_stream(dev)
    tx = capi.Tx(device=local, cuda_stream=cs.cuda_stream)
    g = torch.Generator(device="cpu").manual_seed(21 + rank)
    pay = torch.randint(0, 256, (n, 1504), dtype=torch.uint8, generator=g).to(dev)
    out = torch.empty((n, L), dtype=torch.complex64, device=dev)
    P = (capi.TxProps * n)(*[capi.TxProps(CRC24, 1, 1, QAM16) for _ in range(n)])
    lens = (C.c_uint32 * n)(*([PAYLOAD] * n))
    pp = (C.c_void_p * n)(*[pay[i].data_ptr() for i in range(n)])
    op = (C.c_void_p * n)(*[out[i].data_ptr() for i in range(n)])

    def step():
        capi._check(tx._L.lqb_tx_assemble(tx._h, n, P, None, pp, lens, op, capi.MEM_DEVICE))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    k_ms = 0.0
    for _ in range(args.steps):
        step()
        k_ms += tx.kernel_ms()
    e1.record(cs)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    # kernel-only time: the same launches again with events straight around them (assemble is synchronous, so the
    # events bracket plan + H2D of the frame table + kernel; the kernel share is what ncu reports in profiles/)
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    secs = float(tt[0]) / 1e3
    value = world * n * L * args.steps / secs / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import lqo_py as o
        hp = pay[:64].cpu().numpy()
        t0 = time.perf_counter()
        k = 0
        while time.perf_counter() - t0 < 5.0:
            ref = o.tx_frame(QAM16, CRC24, 1, 1, hp[k % 64][:PAYLOAD])
            if k == 0:
                assert np.allclose(out[0].cpu().numpy(), ref, atol=2e-6), "k_tx differs from the oracle frame generator"
            k += 1
        cpu_s = time.perf_counter() - t0
        cpu = {"value": k * L / cpu_s / 1e6, "unit": "Msps", "cores": 1, "kind": "port", "sample": "%d frames on one thread, %.1f s" % (k, cpu_s)}
    gbs = 8.0 * n * L * args.steps / (k_ms / 1e3) / 1e9 if k_ms else 0.0          # this rank's k_tx alone
    print(json.dumps({
        "metric": "flex_tx_msps", "value": value, "unit": "Msps", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": float(tt[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "flex_tx_batch_qam16_1500B", "frames_per_gpu_per_step": n, "samples_per_frame": L, "mod": "QAM16", "fec0": "none",
                   "fec1": "none", "check": "crc24", "l2": "outputs (%.2f GB per step) larger than L2" % (n * L * 8 / 1e9)},
        "frames_per_s": world * n * args.steps / secs, "clocks": clk, "gpu_launches": args.steps,
        "api": "lqb_tx_assemble (= lqb_tx_submit + lqb_tx_collect: host plan + frame table H2D + k_tx + stream sync inside the timed region)",
        "kernel_ms_per_step": k_ms / args.steps,
        "roofline": {"bound": "hbm", "kernel": "k_tx", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": None, "note": "algorithmic bytes = 16 B written per symbol (2 samples x 8 B) over the kernel's own device time "
                                              "(CUDA events around k_tx, lqb_tx_last_timing); value / ms_per_step are the whole call"},
        "cpu_baseline": cpu,
    }))
    if world > 1:
        dist.destroy_process_group()

# ----------------------------------------------------------------------------- configs[3]: mixed modulation, closed loop
def _parity_rows(torch, cap, rows, got_rec, cores):
    """Oracle (fresh flexframesync per row) against the GPU records of the same rows of the same capture: position,
    flags, scheme fields and payload bytes.  got_rec: structured array from Rx.poll_array() with host results."""
    import ctypes as C
    import numpy as np
    host = cap[rows].cpu().numpy()
    ref = oracle_frames(host, cores)
    frames = mism = 0
    detail = []
    for k, r in enumerate(rows):
        mine = got_rec[got_rec["stream"] == r]
        mine = mine[np.argsort(mine["seq"], kind="stable")]
        frames += len(ref[k])
        if len(mine) != len(ref[k]):
            mism += abs(len(mine) - len(ref[k])) + 1
            detail.append({"row": int(r), "oracle_frames": [(f["sample_index"], f["header_valid"], f["payload_valid"], f["mod_scheme"], f["fec0"], f["fec1"]) for f in ref[k]],
                           "gpu_frames": [(int(b["sample_index"]), int(b["header_valid"]), int(b["payload_valid"]), int(b["mod_scheme"]), int(b["fec0"]), int(b["fec1"]), int(b["flags"])) for b in mine]})
            continue
        for a, b in zip(ref[k], mine):
            why = [f for f in ("sample_index", "header_valid", "payload_valid") if a[f] != int(b[f])]
            if a["header"] != bytes(b["header"]):
                why.append("header")
            if not why and a["header_valid"]:
                why = [f for f in ("mod_scheme", "fec0", "fec1", "payload_len") if a[f] != int(b[f])]
                if not why and int(b["payload"]):
                    pb = C.string_at(int(b["payload"]), int(b["payload_len"]))
                    if a["payload"] != pb:
                        why.append("payload bytes (%d of %d differ)" % (sum(x != y for x, y in zip(a["payload"], pb)), len(pb)))
            if why:
                mism += 1
                detail.append({"row": int(r), "sample_index": a["sample_index"], "scheme": (a["mod_scheme"], a["fec0"], a["fec1"]), "differs": why,
                               "oracle": (a["header_valid"], a["payload_valid"], round(a["evm"], 3)), "gpu": (int(b["header_valid"]), int(b["payload_valid"]), round(float(b["evm"]), 3))})
    out = {"streams": len(rows), "frames": frames, "mismatches": mism}
    if detail:
        out["detail"] = detail[:8]
    return out


def mixed_mod_arm(args, rank, local, world, steps=None, warmup=None, emit=True):
    """BASELINE.json configs[3]: 4096 channels in total, channel c on rank c mod G (strong scaling: the channel count is
    fixed), every channel's scheme re-chosen each step by the vectorised stand-in for cognitive_engine.py over its 616
    configurations plus the flagged extension configurations (QAM128/256, v27p34, K = 9 codes), closed loop on the GPU:
        policy.choose -> flex_tx batch (k_tx writes every frame in place into the capture) -> AWGN at the channel's SNR
        -> flex_rx batch -> packet_info records -> policy.update.
    `value` is the receiver's rate over the step's samples (the metric of the hot path); the step breakdown says what
    the loop around it costs.  Parity: the oracle on a subsample of the channels of the last step (rx reset before it)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from liquiddsp import capi, policy
    from liquiddsp.blocks import MODULATION, INNER_CODE, OUTER_CODE, N_MOD_REF, N_INNER_REF
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    standalone = emit
    steps = max(2, args.steps if steps is None else steps)
    warmup = args.warmup if warmup is None else warmup
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    C_TOTAL, NM, PM, LEAD, GAP = 4096, 1 << 17, 256, 700, 900
    chans = np.arange(rank, C_TOTAL, world, dtype=np.int64)            # global ids of my channels (liquiddsp/sharding.py)
    S = len(chans)
    # a stream of our own: torch's default stream has handle 0, which the C-ABI reads as "no caller stream" (no ordering)
    cs = torch.cuda.Stream(dev)
    torch.cuda.set_stream(cs)
    tx = capi.Tx(device=local, cuda_stream=cs.cuda_stream)
    rx = capi.Rx(S, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream)
    pol = policy.EpsilonGreedy(S, epsilon=0.25, seed=100 + rank, extended=True)
    cap = torch.zeros((S, NM), dtype=torch.complex64, device=dev)
    g = torch.Generator(device=dev).manual_seed(31 + rank)
    pool = torch.randint(0, 256, (1 << 22,), dtype=torch.uint8, device=dev, generator=g)      # payload bytes
    snr_db = 4.0 + (chans * 7 % 23).astype(np.float64)                 # 4 .. 26 dB, fixed per channel
    nstd = torch.tensor(10.0 ** (-snr_db / 20.0) / np.sqrt(2.0), dtype=torch.float32, device=dev)[:, None]
    mod_lut, in_lut, out_lut = np.full(64, -1), np.full(64, -1), np.full(64, -1)
    mod_lut[MODULATION] = np.arange(len(MODULATION)); in_lut[INNER_CODE] = np.arange(len(INNER_CODE)); out_lut[OUTER_CODE] = np.arange(len(OUTER_CODE))
    flen = {}

    def frame_len(key):
        if key not in flen:
            flen[key] = capi.Tx.frame_len(key[0], CRC24, key[1], key[2], PM)
        return flen[key]

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = dict(tx=0.0, noise=0.0, rx=0.0, host=0.0, frames=0, valid=0, sent=0, ext=0, loop=0.0)
    cfg_seen = set()
    last = None
    for it in range(warmup + steps):
        timed = it >= warmup
        if it == warmup + steps - 1:
            rx.reset()                       # the oracle below starts fresh too
        t0 = time.perf_counter()
        m, i, o_ = pol.choose_arrays()
        ms_a, f0_a, f1_a = np.asarray(MODULATION)[m], np.asarray(INNER_CODE)[i], np.asarray(OUTER_CODE)[o_]
        key = ms_a.astype(np.int64) * 4096 + f0_a.astype(np.int64) * 64 + f1_a.astype(np.int64)
        uk, inv = np.unique(key, return_inverse=True)         # frame lengths per distinct configuration, not per channel
        L = np.array([frame_len((int(k >> 12), int((k >> 6) & 63), int(k & 63))) for k in uk], np.int64)[inv]
        nfr = np.maximum(0, (NM - LEAD - 700) // (L + GAP))           # whole frames only; 700 samples of tail for the last one
        tot = int(nfr.sum())
        ch_of = np.repeat(np.arange(S), nfr)
        k_of = np.arange(tot) - np.repeat(np.cumsum(nfr) - nfr, nfr)
        out_ptr = (cap.data_ptr() + 8 * (ch_of * NM + LEAD + k_of * (L[ch_of] + GAP))).astype(np.uint64)
        pay_ptr = (pool.data_ptr() + ((ch_of * 977 + k_of * 131 + it * 7919) * PM) % (pool.numel() - PM)).astype(np.uint64)
        props = np.stack([np.full(tot, CRC24), f0_a[ch_of], f1_a[ch_of], ms_a[ch_of]], axis=1).astype(np.uint32)
        t1 = time.perf_counter()
        ev[0].record(cs)
        cap.zero_()
        t1b = time.perf_counter()
        tx.assemble_device_arrays(props, pay_ptr, np.full(tot, PM, np.uint32), out_ptr, wait=False)
        t1c = time.perf_counter()
        ev[1].record(cs)
        for r0 in range(0, S, 512):          # AWGN in row blocks (bounded temporary)
            r1 = min(S, r0 + 512)
            cap[r0:r1] += nstd[r0:r1] * torch.view_as_complex(torch.randn((r1 - r0, NM, 2), generator=g, device=dev, dtype=torch.float32))
        ev[2].record(cs)
        rx.execute_dense_ptr(cap.data_ptr(), NM, NM, capi.MEM_DEVICE)
        ev[3].record(cs)
        rec = rx.poll_array()
        if timed:
            kt_ = rx.timing()
            acc["k"] = [a + b for a, b in zip(acc.get("k", [0.0] * 6), kt_)]
        t2 = time.perf_counter()
        hv = rec["header_valid"] != 0
        pol.update_arrays(rec["stream"][hv], mod_lut[rec["mod_scheme"][hv] % 64], in_lut[rec["fec0"][hv] % 64], out_lut[rec["fec1"][hv] % 64],
                          rec["payload_valid"][hv])
        torch.cuda.synchronize(dev)
        tx.collect()
        t3 = time.perf_counter()
        if timed:
            acc["txk"] = acc.get("txk", 0.0) + tx.kernel_ms(); acc["txcall"] = acc.get("txcall", 0.0) + 1e3 * (t1c - t1b)
            acc["tx"] += ev[0].elapsed_time(ev[1]); acc["noise"] += ev[1].elapsed_time(ev[2]); acc["rx"] += ev[2].elapsed_time(ev[3])
            acc["host"] += 1e3 * ((t1 - t0) + (t3 - t2)); acc["loop"] += 1e3 * (t3 - t0)
            acc["frames"] += len(rec); acc["valid"] += int((rec["payload_valid"] != 0).sum()); acc["sent"] += tot
            acc["ext"] += int(((m >= N_MOD_REF) | (i >= N_INNER_REF))[ch_of].sum())
            cfg_seen.update(zip(m.tolist(), i.tolist(), o_.tolist()))
        last = rec
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        rows = list(range(0, S, max(1, S // 16)))[:16]
        parity = _parity_rows(torch, cap, rows, last, os.cpu_count() or 1)
    tt = torch.tensor([acc["rx"], acc["loop"], acc["tx"], acc["noise"], acc["host"], acc.get("txk", 0.0), acc.get("txcall", 0.0)], dtype=torch.float64, device=dev)
    sm = torch.tensor([acc["frames"], acc["valid"], acc["sent"], acc["ext"], float(S)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    rx.close(); tx.close()
    torch.cuda.synchronize(dev)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    del cap, pool
    torch.cuda.empty_cache()
    if rank != 0:
        if world > 1 and standalone:
            dist.destroy_process_group()
        return None
    failed = "GPU frames differ from the oracle's on the sampled channels" if parity and parity["mismatches"] else None
    rx_s = float(tt[0]) / 1e3
    chans_all = float(sm[4])
    out = {
        "metric": "flex_rx_msps", "value": chans_all * NM * steps / rx_s / 1e6, "unit": "Msps", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": float(tt[0]) / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "flex_rx_mixed_mod_4096ch_policy_loop", "channels_total": C_TOTAL, "channels_per_gpu": S,
                   "samples_per_channel_per_step": NM, "payload_bytes": PM, "schemes": "616 reference configurations + %d extension (non-reference) per channel and step" % (pol.n_cfg - 616),
                   "snr_db": "4..26 per channel", "l2": "inputs (%.1f GB per step) larger than L2" % (S * NM * 8 / 1e9)},
        "frames_per_s": float(sm[0]) / rx_s, "decoded_frames_per_s": float(sm[1]) / rx_s,
        "frames_sent_per_step": float(sm[2]) / steps, "frames_found_per_step": float(sm[0]) / steps, "frames_valid_per_step": float(sm[1]) / steps,
        "extension_frames_per_step": float(sm[3]) / steps, "distinct_configs_rank0": len(cfg_seen),
        "loop_ms_per_step": {"flex_tx_batch": float(tt[2]) / steps, "k_tx": float(tt[5]) / steps, "lqb_tx_submit_host": float(tt[6]) / steps,
                             "awgn": float(tt[3]) / steps, "flex_rx_batch": float(tt[0]) / steps,
                             "policy_and_lists_host": float(tt[4]) / steps, "whole_loop": float(tt[1]) / steps},
        "rx_kernel_ms_per_step_rank0": dict(zip(["seek_align_header", "matched_filter", "pll_demod", "fec_crc"], [k / steps for k in acc.get("k", [0.0] * 6)[:4]])),
        "closed_loop_msps": chans_all * NM * steps / (float(tt[1]) / 1e3) / 1e6,
        "parity_sample": parity, "failed": failed,
    }
    if emit:
        print(json.dumps(out))
    if world > 1 and standalone:
        dist.destroy_process_group()
    if emit and failed:
        raise SystemExit("mixed_mod: " + failed)
    return out


# ----------------------------------------------------------------------------- configs[4]: TX -> AWGN -> RX, PER vs SNR
def tx_rx_per_arm(args, rank, local, world, steps=None, warmup=None, emit=True):
    """BASELINE.json configs[4]: 8192 channels in total (channel c on rank c mod G), QAM16 / no FEC / CRC-24 / 1500-byte
    frames GENERATED ON THE GPU by the flex_tx batch each step (fresh payloads), AWGN at the channel's SNR 6 .. 24 dB in
    1 dB steps, received by the flex_rx batch; PER per SNR point from payload_valid, beside the oracle's PER on a
    64-channel subsample of the same captures (95 % Clopper-Pearson intervals must overlap; flags are in fact equal)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from liquiddsp import capi
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    standalone = emit
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    QAM16 = 27
    C_TOTAL, NT, LEAD, GAP, NF = 8192, 1 << 14, 600, 1000, 2
    chans = np.arange(rank, C_TOTAL, world, dtype=np.int64)
    S = len(chans)
    L = capi.Tx.frame_len(QAM16, CRC24, 1, 1, PAYLOAD)
    assert LEAD + NF * (L + GAP) <= NT
    cs = torch.cuda.Stream(dev)               # see mixed_mod_arm: handle 0 would mean "no caller stream"
    torch.cuda.set_stream(cs)
    tx = capi.Tx(device=local, cuda_stream=cs.cuda_stream)
    rx = capi.Rx(S, device=local, max_frame_samples=16384, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream)
    cap = torch.zeros((S, NT), dtype=torch.complex64, device=dev)
    pay = torch.zeros((S * NF, 1504), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(41 + rank)
    point = (chans % 19).astype(np.int64)                                # SNR point of each channel: 6 + point dB
    nstd = torch.tensor(10.0 ** (-(6.0 + point) / 20.0) / np.sqrt(2.0), dtype=torch.float32, device=dev)[:, None]
    ch_of = np.repeat(np.arange(S), NF)
    k_of = np.tile(np.arange(NF), S)
    out_ptr = (cap.data_ptr() + 8 * (ch_of * NT + LEAD + k_of * (L + GAP))).astype(np.uint64)
    pay_ptr = (pay.data_ptr() + 1504 * np.arange(S * NF)).astype(np.uint64)
    props = np.tile(np.array([CRC24, 1, 1, QAM16], np.uint32), (S * NF, 1))
    lens = np.full(S * NF, PAYLOAD, np.uint32)
    sub = list(range(0, S, max(1, S // max(1, 64 // world))))[:max(1, 64 // world)] if rank == 0 else []
    sent_pt, ok_pt = np.zeros(19), np.zeros(19)
    o_sent, o_ok, o_same, o_frames = np.zeros(19), np.zeros(19), 0, 0
    tx_ms = rx_ms = txk_ms = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    subs = []
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            e0.record(cs)
        pay.random_(0, 256, generator=g)
        cap.zero_()
        ev[0].record(cs)
        tx.assemble_device_arrays(props, pay_ptr, lens, out_ptr, wait=False)      # lqb_tx_submit: complete in stream order
        ev[1].record(cs)
        cap += nstd * torch.view_as_complex(torch.randn((S, NT, 2), generator=g, device=dev, dtype=torch.float32))
        ev[2].record(cs)
        rx.reset()                                # every step is a fresh capture (frames never straddle steps)
        rx.execute_dense_ptr(cap.data_ptr(), NT, NT, capi.MEM_DEVICE)
        ev[3].record(cs)
        rec = rx.poll_array()
        if it >= warmup:
            okc = np.bincount(rec["stream"][rec["payload_valid"] != 0], minlength=S)
            np.add.at(ok_pt, point, okc)
            np.add.at(sent_pt, point, NF)
            if sub and not args.no_cpu_baseline:
                subs.append((cap[sub].cpu().numpy(), rec[np.isin(rec["stream"], sub)].copy()))
            torch.cuda.synchronize(dev)
            tx.collect()
            tx_ms += ev[0].elapsed_time(ev[1]); rx_ms += ev[2].elapsed_time(ev[3]); txk_ms += tx.kernel_ms()
    e1.record(cs)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if subs:
        cores = os.cpu_count() or 1
        for host, mine in subs:
            ref = oracle_frames(host, cores)
            for k, r in enumerate(sub):
                mr = mine[mine["stream"] == r]
                o_sent[point[r]] += NF
                o_ok[point[r]] += sum(1 for f in ref[k] if f["payload_valid"])
                o_frames += len(ref[k])
                o_same += int(len(mr) == len(ref[k]) and all(a["sample_index"] == int(b["sample_index"]) and a["payload_valid"] == int(b["payload_valid"])
                                                             for a, b in zip(ref[k], mr[np.argsort(mr["seq"], kind="stable")])))
    tt = torch.tensor([ms, tx_ms, rx_ms, txk_ms], dtype=torch.float64, device=dev)
    sm = torch.tensor(np.concatenate([sent_pt, ok_pt, [float(S)]]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    rx.close(); tx.close()
    torch.cuda.synchronize(dev)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    del cap, pay
    torch.cuda.empty_cache()
    if rank != 0:
        if world > 1 and standalone:
            dist.destroy_process_group()
        return None
    sent_all, ok_all, chans_all = sm[:19].cpu().numpy(), sm[19:38].cpu().numpy(), float(sm[38])
    per = [round(float(1.0 - b / a), 5) if a else None for a, b in zip(sent_all, ok_all)]
    oracle = failed = None
    if subs:
        from scipy.stats import beta as _beta

        def cp(k, n):          # 95 % Clopper-Pearson interval of k errors in n frames
            lo = 0.0 if k == 0 else float(_beta.ppf(0.025, k, n - k + 1))
            hi = 1.0 if k == n else float(_beta.ppf(0.975, k + 1, n - k))
            return lo, hi
        overlap, o_per = 0, []
        for p_ in range(19):
            if not o_sent[p_]:
                o_per.append(None); continue
            lo1, hi1 = cp(int(o_sent[p_] - o_ok[p_]), int(o_sent[p_]))
            lo2, hi2 = cp(int(sent_all[p_] - ok_all[p_]), int(sent_all[p_]))
            overlap += int(lo1 <= hi2 and lo2 <= hi1)
            o_per.append(round(float(1.0 - o_ok[p_] / o_sent[p_]), 5))
        oracle = {"channels": len(sub), "frames": int(o_frames), "per_by_snr_point": o_per, "points_with_overlapping_95pct_intervals": overlap,
                  "points": int((o_sent > 0).sum()), "channel_steps_with_identical_flags": int(o_same), "channel_steps": len(sub) * len(subs)}
        if overlap != oracle["points"] or o_same != oracle["channel_steps"]:
            failed = "GPU PER / flags differ from the oracle's on the sampled channels"
    secs = float(tt[0]) / 1e3
    hbm_peak = float(load_peaks().get("hbm_gbs", 6650.0))
    tx_gbs = 8.0 * chans_all / world * NF * L * steps / (float(tt[3]) / 1e3) / 1e9 if tt[3] > 0 else None
    out = {
        "metric": "flex_tx_rx_msps", "value": chans_all * NT * steps / secs / 1e6, "unit": "Msps", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": float(tt[0]) / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "flex_tx_awgn_flex_rx_8192ch_qam16_1500B", "channels_total": C_TOTAL, "channels_per_gpu": S, "samples_per_channel_per_step": NT,
                   "frames_per_channel_per_step": NF, "mod": "QAM16", "fec0": "none", "fec1": "none", "check": "crc24", "snr_db": "6..24 in 1 dB steps, one point per channel",
                   "l2": "capture (%.1f GB per step) larger than L2" % (S * NT * 8 / 1e9)},
        "frames_per_s": float(sent_all.sum()) / secs, "decoded_frames_per_s": float(ok_all.sum()) / secs,
        "snr_db_points": [6 + k for k in range(19)], "per_by_snr_point": per, "frames_per_point": [int(a) for a in sent_all],
        "ms_per_step_parts": {"flex_tx_batch_call": float(tt[1]) / steps, "k_tx": float(tt[3]) / steps, "flex_rx_batch": float(tt[2]) / steps},
        "roofline": {"bound": "hbm", "kernel": "k_tx", "achieved": tx_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": tx_gbs / hbm_peak if tx_gbs else None,
                     "traffic": None, "note": "16 B written per symbol over k_tx's own device time (lqb_tx_last_timing)"},
        "oracle_per": oracle, "failed": failed,
    }
    if emit:
        print(json.dumps(out))
    if world > 1 and standalone:
        dist.destroy_process_group()
    if emit and failed:
        raise SystemExit("tx_rx_per: " + failed)
    return out


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from liquiddsp import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if args.workload == "detector":
        detector_arm(args, rank, local, world)
        return
    if args.workload == "tx":
        tx_arm(args, rank, local, world)
        return
    if args.workload == "mixed_mod":
        mixed_mod_arm(args, rank, local, world)
        return
    if args.workload == "tx_rx_per":
        tx_rx_per_arm(args, rank, local, world)
        return
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # all of this process's torch work and the library's caller-stream ordering use one explicit stream (torch's default
    # stream has handle 0, which the C-ABI reads as "no caller stream")
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    near = bind_near_gpu(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S, N = args.streams, args.samples
    frames, payloads = clean_frames_ours(torch, dev, 1)
    cap, sent = make_capture(torch, frames, S, N, 1, dev, stream_offset=rank * S)
    torch.cuda.synchronize(dev)
    cs = torch.cuda.current_stream(dev)
    rx = capi.Rx(S, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=args.lanes)
    lanes = rx.lanes()

    def step():
        rx.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    pipelined = not args.no_pipeline
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = rx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt = [0.0] * 6
    work = dict(windows=0, aligns=0, symbols=0, samples=0, exact_windows=0, coarse_tiles=0, exact_bins=0)
    fr_tot = va_tot = 0
    def account():
        nonlocal kt, fr_tot, va_tot
        t = rx.timing()
        kt = [a + b for a, b in zip(kt, t)]
        w = rx.work()
        for k in work:
            work[k] += w[k]
        f, v = rx.counts()
        fr_tot += f; va_tot += v

    torch.cuda.synchronize(dev)
    t_wall = time.perf_counter()
    e0.record(cs)
    if pipelined:
        # lqb_rx_submit / lqb_rx_collect: the payload work of step k runs under the search of step k+1; every
        # one of the K steps is submitted AND collected (all its frames on the host side of the API) inside the region
        for i in range(args.steps):
            rx.submit_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
            if i:
                rx.collect(); account()
        rx.collect(); account()
    else:
        for _ in range(args.steps):
            step()
            account()
    e1.record(cs)
    torch.cuda.synchronize(dev)
    t_wall = (time.perf_counter() - t_wall) * 1e3
    if world > 1:
        dist.barrier()
    ms = max(e0.elapsed_time(e1), t_wall)      # the library works on its own streams: the host clock bounds the region too
    clk = clocks.stop()
    launches = rx.launches() - l0
    tt = torch.tensor([ms, float(fr_tot), float(va_tot), float(launches), float(sent)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms = float(mx[0]); fr_all, va_all, launches_all, sent_all = float(sm[1]), float(sm[2]), float(sm[3]), float(sm[4])
    else:
        fr_all, va_all, launches_all, sent_all = float(fr_tot), float(va_tot), float(launches), float(sent)
    secs = ms / 1e3
    value = world * S * N * args.steps / secs / 1e6

    # ---- per-kernel breakdown: in the timed region the lanes' kernels overlap on the GPU, so the kernel times
    # used for the roofline come from the same K steps repeated with one lane (kernels back to back on one stream)
    serial_ms = None
    if lanes > 1:
        rx.close()
        rx = capi.Rx(S, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=1)
        step(); step()
        torch.cuda.synchronize(dev)
        kt = [0.0] * 6
        work = {k: 0 for k in work}
        e0.record(cs)
        for _ in range(args.steps):
            step()
            t = rx.timing()
            kt = [a + b for a, b in zip(kt, t)]
            w = rx.work()
            for k in work:
                work[k] += w[k]
        e1.record(cs)
        torch.cuda.synchronize(dev)
        serial_ms = e0.elapsed_time(e1) / args.steps
    rx.close()

    # ---- e2e: host buffers through the same C-ABI call (H2D + all results D2H inside the timed region)
    e2e = e2e_sc16 = None
    if not args.no_e2e:
        Ne = min(args.e2e_samples, N) if args.e2e_samples else N

        def e2e_leg(host, mem, bytes_per_sample):
            """K steps through the C-ABI with pinned HOST input `host` (H2D inside the timed region) and every result
            (frame records, payload bytes, constellation points) read back to the host inside it."""
            rx2 = capi.Rx(S, device=local, max_frame_samples=65536, flags=0, cuda_stream=cs.cuda_stream, lanes=args.e2e_lanes or args.lanes)
            for _ in range(max(2, args.warmup)):
                rx2.execute_dense_ptr(host.data_ptr(), Ne, Ne, mem)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            d2h = fr = va = 0

            def results():
                rec = rx2.poll_array()
                return int(rec["payload_len"].sum() + 8 * rec["num_framesyms"].sum() + 256 * len(rec)), len(rec), int((rec["payload_valid"] != 0).sum())

            t0 = time.perf_counter()
            e0.record(cs)
            for i in range(args.steps):
                if pipelined:
                    rx2.submit_dense_ptr(host.data_ptr(), Ne, Ne, mem)
                    if not i:
                        continue
                    rx2.collect()
                else:
                    rx2.execute_dense_ptr(host.data_ptr(), Ne, Ne, mem)
                a_, b_, c_ = results(); d2h += a_; fr += b_; va += c_
            if pipelined:
                rx2.collect()
                a_, b_, c_ = results(); d2h += a_; fr += b_; va += c_
            e1.record(cs)
            torch.cuda.synchronize(dev)
            wall = time.perf_counter() - t0
            ems = max(e0.elapsed_time(e1), wall * 1e3)
            te = torch.tensor([ems], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            out_ = {"value": world * S * Ne * args.steps / (float(te[0]) / 1e3) / 1e6, "unit": "Msps",
                    "h2d_bytes_per_step": S * Ne * bytes_per_sample, "d2h_bytes_per_step": d2h // max(args.steps, 1),
                    "samples_per_stream_per_step": Ne, "lanes": rx2.lanes(), "frames_found_per_step": fr / args.steps,
                    "frames_valid_per_step": va / args.steps, "h2d_gbs_per_gpu": S * Ne * bytes_per_sample * args.steps / (float(te[0]) / 1e3) / 1e9}
            rx2.close()
            return out_

        host = torch.empty((S, Ne), dtype=torch.complex64).pin_memory()
        host.copy_(cap[:, :Ne])
        e2e = e2e_leg(host, capi.MEM_HOST, 8)
        e2e["input"] = "complex64 (the reference's gr_complex stream), pinned host memory"
        del host
        # additive input format: interleaved int16 pairs (what SDR front ends deliver), half the PCIe bytes; the capture is
        # the same one scaled by 4096 and rounded (the receiver is gain invariant; the quantisation noise is ~60 dB down)
        host16 = torch.empty((S, Ne, 2), dtype=torch.int16).pin_memory()
        for r0 in range(0, S, 64):
            host16[r0:r0 + 64].copy_(torch.view_as_real(cap[r0:r0 + 64, :Ne]).mul(4096.0).round_().clamp_(-32768, 32767).to(torch.int16))
        e2e_sc16 = e2e_leg(host16, capi.MEM_HOST_SC16, 4)
        e2e_sc16["input"] = "sc16 (LQB_MEM_HOST_SC16: interleaved int16 pairs, value / 32768), pinned host memory"
        del host16

    # ---- CPU baseline + parity on a bounded sample of the same capture (rank 0, N = 1 only): the GPU's frame records of
    # the sampled streams (fresh receiver, one full-size step, host results) are kept for the comparison below
    cpu_sample = gpu_rec = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        Sc = min(S, 8 * cores, 256)
        cpu_sample = cap[:Sc].cpu().numpy()
        rx3 = capi.Rx(S, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=args.lanes)
        rx3.execute_dense_ptr(cap.data_ptr(), N, N, capi.MEM_DEVICE)
        rec = rx3.poll_array()
        rec = rec[rec["stream"] < Sc]
        import ctypes as _C
        gpu_rec = [(int(r["stream"]), int(r["seq"]), int(r["sample_index"]), int(r["header_valid"]), int(r["payload_valid"]), bytes(r["header"]),
                    _C.string_at(int(r["payload"]), int(r["payload_len"])) if (r["header_valid"] and r["payload"]) else b"") for r in rec]
        rx3.close()
    # ---- the single-stream case (what ONE flex_rx block instance sees): the whole capture, row after row, as one stream,
    # decoded by the time-sharded receiver (lqb_rx_execute_sharded: the sequential receiver's frames whatever the cut),
    # beside the same receiver fed the ordinary way (one CTA walks the stream) on a prefix
    single = single_prefix = single_rec = None
    if world == 1 and not args.no_workloads:
        seg_len, preroll = int(os.environ.get("LQB_BENCH_RX_SEG_LEN", 1 << 20)), int(os.environ.get("LQB_BENCH_RX_PREROLL", 1 << 16))
        rx4 = capi.Rx(S, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream, lanes=args.lanes)
        rx4.execute_sharded_ptr(cap.data_ptr(), S * N, capi.MEM_DEVICE, seg_len, preroll)           # warm-up (arenas)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        rx4.execute_sharded_ptr(cap.data_ptr(), S * N, capi.MEM_DEVICE, seg_len, preroll)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        f4, v4 = rx4.counts()
        single = {"samples": S * N, "seg_len": seg_len, "preroll": preroll, "ms": dt * 1e3, "msps": S * N / dt / 1e6,
                  "frames_found": f4, "frames_valid": v4}
        single.update(rx4.shard_info())
        n_pref = min(S * N, 3 << 20)
        rec4 = rx4.poll_array()
        rec4 = rec4[rec4["sample_index"] < n_pref - 70000]
        import ctypes as _C
        single_rec = [(int(r["sample_index"]), int(r["header_valid"]), int(r["payload_valid"]), bytes(r["header"]),
                       _C.string_at(int(r["payload"]), int(r["payload_len"])) if (r["header_valid"] and r["payload"]) else b"") for r in rec4]
        single_prefix = cap.reshape(-1)[:n_pref].cpu().numpy() if not args.no_cpu_baseline else None
        rx4.close()
        n_plain = min(S * N, 16 << 20)
        rx5 = capi.Rx(1, device=local, max_frame_samples=65536, flags=capi.RX_NO_FRAMESYMS, cuda_stream=cs.cuda_stream)
        rx5.execute_dense_ptr(cap.data_ptr(), n_plain, n_plain, capi.MEM_DEVICE)
        rx5.reset()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        rx5.execute_dense_ptr(cap.data_ptr(), n_plain, n_plain, capi.MEM_DEVICE)
        torch.cuda.synchronize(dev)
        single["plain_one_cta"] = {"samples": n_plain, "msps": n_plain / (time.perf_counter() - t0) / 1e6}
        rx5.close()
    del cap
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, short runs (every rank takes part: they shard like the headline)
    workloads = None
    if not args.no_workloads:
        workloads = {}
        workloads["detector"] = detector_arm(args, rank, local, world, steps=3, warmup=1, emit=False)
        workloads["mixed_mod"] = mixed_mod_arm(args, rank, local, world, steps=3, warmup=2, emit=False)
        workloads["tx_rx_per"] = tx_rx_per_arm(args, rank, local, world, steps=5, warmup=2, emit=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (per-kernel device times come from CUDA events on the launching stream, summed over the timed steps)
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    names = ["seek_align_header", "matched_filter", "pll_demod", "fec_crc"]
    t_seek, t_mf, t_pll, t_fec = [k / 1e3 for k in kt[:4]]
    t_coarse = kt[5] / 1e3
    win_bytes = 8.0 * 256.0 * work["windows"]                     # 8 B per new sample a detector window examines
    win_flops = work["exact_windows"] * (50 * 9 * 256 * 10 + 49 * 512 * 9.0)
    # tensor-core pre-filter, ALGORITHMIC flops (dense form): 156 template samples x 49 CFO bins x 8
    # real flops per lag, 128 lags per tile; the issued MMAs are K = 320 x N = 112 (padding +17 %), reported separately.
    # Operands are fp8 (e4m3): the tensor peak for them is twice the bf16 figure; MEASURED_PEAKS.json holds no fp8
    # measurement, so the denominator is 2 x the measured sustained bf16 rate (nominal 4.5 vs 2.25 PFLOP/s dense).
    tc_flops = work["coarse_tiles"] * 128.0 * 156.0 * 49.0 * 8.0
    tc_flops_issued = work["coarse_tiles"] * 128.0 * 320.0 * 112.0 * 2.0
    tc_peak = 2.0 * float(peaks.get("bf16_tflops_sustained", 1400.0))
    mf_bytes = 8.0 * (2.0 * work["symbols"]) + 8.0 * work["symbols"]   # 2 samples read + 1 symbol written per symbol
    fp32_peak = 148 * 128 * 2 * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e12
    kernels = [
        {"name": names[0], "ms_per_step": 1e3 * t_seek / args.steps, "bound": "fp32",
         "hbm_gbs": win_bytes / t_seek / 1e9 if t_seek else None,
         "hbm_frac": win_bytes / t_seek / 1e9 / hbm_peak if t_seek else None,
         "fp32_tflops": win_flops / t_seek / 1e12 if t_seek else None,
         "fp32_frac": win_flops / t_seek / 1e12 / fp32_peak if t_seek else None,
         "windows_per_step": work["windows"] / args.steps, "exact_windows_per_step": work["exact_windows"] / args.steps,
         "cfo_bins_per_exact_window": work["exact_bins"] / max(1, work["exact_windows"]),
         "prefilter_ms_per_step": 1e3 * t_coarse / args.steps,
         "prefilter_tensor_tflops": tc_flops / t_coarse / 1e12 if t_coarse else None,
         "prefilter_tensor_frac": tc_flops / t_coarse / 1e12 / tc_peak if t_coarse else None,
         "prefilter_hbm_gbs": (8.0 * work["samples"]) / t_coarse / 1e9 if t_coarse else None},
        {"name": names[1], "ms_per_step": 1e3 * t_mf / args.steps, "bound": "hbm",
         "hbm_gbs": mf_bytes / t_mf / 1e9 if t_mf else None, "hbm_frac": mf_bytes / t_mf / 1e9 / hbm_peak if t_mf else None},
        {"name": names[2], "ms_per_step": 1e3 * t_pll / args.steps, "bound": "latency",
         "hbm_gbs": 16.0 * work["symbols"] / t_pll / 1e9 if t_pll else None},
        {"name": names[3], "ms_per_step": 1e3 * t_fec / args.steps, "bound": "int-alu"},
    ]
    # the search kernel is tensor-core work (pre-filter) plus a few exact FP32 windows: report it against the tensor peak
    tc_in_seek = t_coarse == 0.0 and work["coarse_tiles"] > 0
    if tc_in_seek:
        kernels[0].update({"bound": "tensor", "tensor_tflops": tc_flops / t_seek / 1e12 if t_seek else None,
                           "tensor_frac": tc_flops / t_seek / 1e12 / tc_peak if t_seek else None,
                           "tensor_tflops_issued": tc_flops_issued / t_seek / 1e12 if t_seek else None,
                           "tensor_frac_of_bf16_peak": tc_flops / t_seek / 1e12 / (tc_peak / 2.0) if t_seek else None,
                           "tensor_tiles_per_step": work["coarse_tiles"] / args.steps})
    dom = max(range(4), key=lambda i: kt[i])
    step_bytes = 8.0 * S * N * args.steps + 8.0 * work["symbols"]   # every input sample once + symbols written
    if dom == 0 and tc_in_seek:
        roof = {"bound": "tensor", "kernel": names[0], "achieved": kernels[0]["tensor_tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                "frac": kernels[0]["tensor_frac"], "traffic": None,
                "peak_source": ("2 x measured bf16_tflops_sustained (MEASURED_PEAKS.json): fp8 operands, no fp8 measurement in the file" if "bf16_tflops_sustained" in peaks
                                else "2 x fallback 1400 TFLOP/s"),
                "note": "algorithmic flops = pre-filter tiles x 128 lags x 156 x 49 x 8 (SURVEY.md 8d dense form; e4m3 in, fp32 accumulate); the kernel "
                        "time also contains the exact FP32 FFT windows, alignment and header decode of every frame.  ncu (profiles/r02_ncu_summary.md): "
                        "the binding resource is shared-memory bandwidth (MMA operand fetch 150 KB + LSU traffic per window), tensor pipe ~40 % active"}
    elif dom == 1:
        roof = {"bound": "hbm", "kernel": names[1], "achieved": kernels[1]["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[1]["hbm_frac"], "traffic": None, "peak_source": peak_src}
    else:
        ach = win_bytes / t_seek / 1e9 if dom == 0 and t_seek else step_bytes / (kt[4] / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                "note": "dominant kernel is the qdetector search, which is FP32-compute-bound (50 FFT-512 per 256 new samples), "
                        "not HBM-bound: see kernels[0].fp32_frac; whole-step HBM fraction in step_hbm_frac"}
    roof["step_hbm_gbs"] = step_bytes / secs / 1e9
    roof["step_hbm_frac"] = roof["step_hbm_gbs"] / hbm_peak
    # measured DRAM traffic of the dominant kernel (one `ncu --set full` capture of this workload, committed under
    # profiles/): reported next to the algorithmic bytes so that wasted re-reads would show
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if tj["config"]["streams"] == S and tj["config"]["samples"] == N:
            kname = {0: "k_seek", 1: "k_mf"}.get(dom)
            if kname in tj["kernels"]:
                roof["traffic"] = tj["kernels"][kname]["dram_gb_per_launch"] * 1e9
                roof["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)"
                roof["algorithmic_bytes"] = (win_bytes if dom == 0 else mf_bytes) / args.steps
                roof["traffic_source"] = "profiles/roofline_traffic.json"
    except Exception:
        pass

    # ---- CPU baseline (timed) and parity of the GPU frame list against the oracle's on that same sample
    cpu = parity = None
    if cpu_sample is not None:
        cores = os.cpu_count() or 1
        Sc, Nc = cpu_sample.shape
        dt, f, v = cpu_rx(cpu_sample, cores)
        cpu = {"value": Sc * Nc / dt / 1e6, "unit": "Msps", "cores": cores, "kind": "port",
               "decoded_frames_per_s": v / dt,
               "sample": "first %d streams x %d samples of the same capture, %d threads, %.1f s" % (Sc, Nc, cores, dt)}
        ref = oracle_frames(cpu_sample, cores)
        mine = {}
        for r in gpu_rec:
            mine.setdefault(r[0], []).append(r)
        n_ref = mism = 0
        for sidx in range(Sc):
            a_, b_ = ref[sidx], sorted(mine.get(sidx, []), key=lambda r: r[1])
            n_ref += len(a_)
            if len(a_) != len(b_):
                mism += abs(len(a_) - len(b_)) + 1
                continue
            for x, y in zip(a_, b_):
                ok = (x["sample_index"] == y[2] and x["header_valid"] == y[3] and x["payload_valid"] == y[4] and x["header"] == y[5]
                      and (not x["header_valid"] or x["payload"] == y[6]))
                mism += 0 if ok else 1
        parity = {"streams": Sc, "frames": n_ref, "mismatches": mism,
                  "compared": "sample_index, header_valid, payload_valid, header bytes, payload bytes; oracle vs one full-size GPU step"}

    if single is not None and single_prefix is not None:
        # ONE sequential oracle receiver over the first samples of the one-stream run
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lqo_py as o
        lim = len(single_prefix) - 70000                  # (frames the prefix holds in full)
        refp = [r for r in o.rx_capture(single_prefix, chunk=4096, max_frames=1 << 16) if r["sample_index"] < lim]
        same = len(refp) == len(single_rec) and all(
            x["sample_index"] == y[0] and x["header_valid"] == y[1] and x["payload_valid"] == y[2] and x["header"] == y[3]
            and (not x["header_valid"] or x["payload"] == y[4]) for x, y in zip(refp, single_rec))
        single["oracle_prefix"] = {"samples": len(single_prefix), "frames": len(refp), "identical": bool(same)}
        if workloads is not None:
            workloads["single_stream"] = single
    elif single is not None and workloads is not None:
        workloads["single_stream"] = single

    out = {
        "metric": "flex_rx_msps", "value": value, "unit": "Msps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": flex_rx_config(args),
        "decoded_frames_per_s": va_all / secs, "frames_per_s": fr_all / secs,
        "frames_sent_per_step": sent_all, "frames_found_per_step": fr_all / args.steps, "frames_valid_per_step": va_all / args.steps,
        "gpu_launches": int(launches_all), "lanes": lanes,
        "api": "lqb_rx_submit + lqb_rx_collect, two calls in flight" if pipelined else "lqb_rx_execute",
        "host_cpus_bound_near_gpu": near,
        "kernel_times": ("CUDA events around each kernel on its launching stream, same K steps repeated with lanes=1 right after the "
                         "timed region (%.2f ms/step serialized; in the timed region the lanes overlap)" % serial_ms) if serial_ms else
                        "CUDA events around each kernel on its launching stream inside the timed region",
        "clocks": clk, "e2e": e2e, "e2e_sc16": e2e_sc16, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "parity_sample": parity,
        "workloads": workloads,
    }
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    if parity and parity["mismatches"]:
        raise SystemExit("flex_rx: GPU frames differ from the oracle's on the sampled streams: %r" % (parity,))
    if single is not None and single.get("oracle_prefix") and not single["oracle_prefix"]["identical"]:
        raise SystemExit("flex_rx: the time-sharded single-stream run differs from the sequential oracle: %r" % (single,))
    bad = [k for k, v in (workloads or {}).items() if v and v.get("failed")]
    if bad:
        raise SystemExit("workloads failed their oracle check: %s" % ", ".join(bad))


if __name__ == "__main__":
    main()
