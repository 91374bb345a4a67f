"""Bulk preamble detection over one long capture (BASELINE.json configs[1]).

qdetector_cccf is strictly sequential (its hop grid re-phases after every detection), so one long stream
is sharded in TIME: overlapping segments are treated as independent streams of a batch detector handle
and detections falling into the overlap are de-duplicated by absolute sample index.  Inside a segment the
behaviour is exactly the sequential detector's; at a seam the hop grid restarts, which can move a detection
by a hop but not lose a frame whose preamble lies wholly inside one segment (overlap >= 512 + 156 samples)."""
import numpy as np

from . import capi

OVERLAP = 1024          # >= nfft (512) + template (156), multiple of 256


def segments(n_total, seg_len):
    """(start, length) of overlapping segments covering [0, n_total)."""
    step = seg_len - OVERLAP
    starts = np.arange(0, max(n_total - OVERLAP, 1), step, dtype=np.int64)
    return [(int(s), int(min(seg_len, n_total - s))) for s in starts]


def dedup(dets, tol=2):
    """dets: iterable of dicts with absolute 'sample_index'; keeps one per frame (first by index)."""
    out = []
    for d in sorted(dets, key=lambda d: d["sample_index"]):
        if out and d["sample_index"] - out[-1]["sample_index"] <= tol:
            continue
        out.append(d)
    return out


class BulkDetector(object):
    def __init__(self, n_segments, device=0, beta=0.0, threshold=0.0, cuda_stream=None):
        self.det = capi.Det(n_segments, device=device, beta=beta, threshold=threshold, cuda_stream=cuda_stream)
        self.n = n_segments

    def run_dense_ptr(self, ptr, stride, seg_len, seg_starts, mem):
        """All segments have seg_len samples at ptr + 8*stride*i; returns de-duplicated absolute detections."""
        self.det.reset()
        self.det.execute_dense_ptr(ptr, stride, seg_len, mem)
        found = self.det.poll()
        for d in found:
            d["sample_index"] += int(seg_starts[d["stream"]])
        return dedup(found)

    def run_host(self, x, seg_len=1 << 18):
        """x: complex64 numpy capture of any length."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        segs = segments(len(x), seg_len)
        out = []
        for i in range(0, len(segs), self.n):
            part = segs[i:i + self.n]
            self.det.reset()
            self.det.execute([x[s:s + l] for s, l in part], list(range(len(part))))
            for d in self.det.poll():
                d["sample_index"] += part[d["stream"]][0]
                out.append(d)
        return dedup(out)
