"""Bulk preamble detection over one long capture (BASELINE.json configs[1]).

qdetector_cccf is strictly sequential (its hop grid re-phases after every detection; reference call site
/root/reference/lib/frame_detector_cc_impl.cc:77), so one long stream is sharded in TIME inside the library
(lqb_det_execute_sharded): segments run side by side as streams of one batch handle, each started speculatively
a pre-roll before its boundary, and a segment's run is accepted only if it entered the segment in exactly the state
the accepted run before it left in -- otherwise the segment is searched again from that state.  The detection list
is therefore the sequential detector's, sample for sample, whatever the cut (tests/test_gpu_tx_det.py); seg_len and
preroll only change the time.  (Round 1 cut the capture into overlapping independent segments and de-duplicated:
decisions could move by a hop at a seam.  That scheme is gone.)"""
import numpy as np

from . import capi


class BulkDetector(object):
    def __init__(self, n_workers, device=0, beta=0.0, threshold=0.0, cuda_stream=None):
        """n_workers: segments searched at a time (streams of the batch handle; >= 444 fills a B200)."""
        self.det = capi.Det(n_workers, device=device, beta=beta, threshold=threshold, cuda_stream=cuda_stream)
        self.n = n_workers

    def run_host(self, x, seg_len=1 << 18, preroll=1 << 14):
        """x: complex64 numpy capture of any length; returns the sequential detector's detections (absolute indices)."""
        self.det.execute_sharded(np.ascontiguousarray(x, dtype=np.complex64), seg_len=seg_len, preroll=preroll)
        return self.det.poll()

    def run_device_ptr(self, ptr, n_samples, seg_len=1 << 18, preroll=1 << 14):
        """The capture already lies in device memory (complex64, contiguous)."""
        self.det.execute_sharded_ptr(ptr, n_samples, capi.MEM_DEVICE, seg_len, preroll)
        return self.det.poll()

    def info(self):
        """segments / runs / rounds / launches of the last call (runs > segments: seams were re-run from the proven state)."""
        return self.det.shard_info()
