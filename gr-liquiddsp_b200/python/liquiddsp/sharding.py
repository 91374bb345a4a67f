"""Channel sharding across GPUs and the host-side gather of per-frame statistics.

Channels are independent, so the data path has no collective: channel c lives on rank c mod G and
each rank runs its own receiver handle.  Only the small per-frame records travel, gathered on the
host with torch.distributed (gloo on CPU in the tests, NCCL-initialised groups also work because the
records are moved as CPU objects)."""
import numpy as np

FRAME_DTYPE = np.dtype([("channel", np.int64), ("seq", np.int64), ("sample_index", np.int64),
                        ("header_valid", np.int8), ("payload_valid", np.int8),
                        ("modulation", np.int16), ("fec0", np.int16), ("fec1", np.int16), ("payload_len", np.int32),
                        ("evm", np.float32), ("rssi", np.float32), ("cfo", np.float32)])


def channels_of_rank(n_channels, rank, world):
    """Global channel ids handled by `rank` (c mod world == rank), in local stream order."""
    return np.arange(rank, n_channels, world, dtype=np.int64)


def local_stream_of(channel, world):
    return channel // world


def records_from_frames(frames, rank, world):
    """frames: dicts from capi.Rx.poll() (local stream ids) -> structured array with global channel ids."""
    rec = np.zeros(len(frames), FRAME_DTYPE)
    for k, f in enumerate(frames):
        rec[k] = (f["stream"] * world + rank, f["seq"], f["sample_index"], f["header_valid"], f["payload_valid"],
                  f["mod_scheme"], f["fec0"], f["fec1"], f["payload_len"], f["evm"], f["rssi"], f["cfo"])
    return rec


def gather_records(rec, dst=0, group=None):
    """Gather every rank's records on rank `dst`, ordered by (channel, seq).  Returns None elsewhere."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.sort(rec, order=["channel", "seq"])
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(rec, out, dst=dst, group=group)
    if out is None:
        return None
    return np.sort(np.concatenate(out), order=["channel", "seq"])
