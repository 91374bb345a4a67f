"""PDU <-> tagged-stream adapters: the step either side of the path in the authors' loopback.

`flex_tx` emits whole frames as PDUs (`cons(PMT_NIL, c32vector(frame_len))`, /root/reference/lib/flex_tx_impl.cc:202-206)
while `flex_rx` consumes a sample STREAM in multiples of 256 items (/root/reference/lib/flex_rx_impl.cc:50,204-215); in a
GNU Radio 3.7 flowgraph the blocks in between are the stock `pdu_to_tagged_stream` / `tagged_stream_to_pdu` (the
un-committed loopback named in /root/reference/python/.idea/workspace.xml:79-86).  These are their flowgraph-free
equivalents with the same semantics -- items plus a length tag (`packet_len` by default) at the first item of each
PDU -- so the loopback can be wired without GNU Radio, and `stream_chunks` turns the tagged stream into the
256-multiples `flex_rx.work` wants (zero padded at the end of the stream, as a throttled flowgraph would idle)."""
import numpy as np

LENGTH_TAG = "packet_len"


class pdu_to_tagged_stream(object):
    """Collects PDUs posted to port 'pdus' and hands them out as one item stream with length tags.

    work(n) returns (items, tags): up to n items (fewer when the queue runs dry) and the tags (offset, key, value) whose
    offsets are absolute item indices since construction, GNU Radio style."""

    def __init__(self, dtype=np.complex64, length_tag=LENGTH_TAG):
        self.dtype = np.dtype(dtype)
        self.length_tag = length_tag
        self._queue = []              # pending PDU payloads
        self._cur = None              # the PDU being streamed and how much of it has gone
        self._pos = 0
        self.nitems_written = 0

    def post(self, port, msg):
        if port != "pdus":
            raise KeyError(port)
        _meta, data = msg
        self._queue.append(np.ascontiguousarray(data, dtype=self.dtype).reshape(-1))

    def pending(self):
        return sum(len(q) for q in self._queue) + (len(self._cur) - self._pos if self._cur is not None else 0)

    def work(self, noutput_items):
        out = np.zeros(noutput_items, self.dtype)
        tags, n = [], 0
        while n < noutput_items:
            if self._cur is None:
                if not self._queue:
                    break
                self._cur, self._pos = self._queue.pop(0), 0
                tags.append((self.nitems_written + n, self.length_tag, len(self._cur)))
            k = min(noutput_items - n, len(self._cur) - self._pos)
            out[n:n + k] = self._cur[self._pos:self._pos + k]
            n += k
            self._pos += k
            if self._pos == len(self._cur):
                self._cur = None
        self.nitems_written += n
        return out[:n], tags


class tagged_stream_to_pdu(object):
    """The reverse: work(items, tags) cuts the stream at every length tag and returns the completed PDUs
    [(None, items)]; a PDU may span several work calls."""

    def __init__(self, dtype=np.complex64, length_tag=LENGTH_TAG):
        self.dtype = np.dtype(dtype)
        self.length_tag = length_tag
        self.nitems_read = 0
        self._need = 0
        self._parts = []

    def work(self, items, tags):
        items = np.asarray(items, dtype=self.dtype).reshape(-1)
        starts = sorted((off - self.nitems_read, val) for off, key, val in tags if key == self.length_tag)
        out, i = [], 0
        while i < len(items):
            if self._need == 0:
                if not starts:
                    break                                  # untagged items between PDUs are dropped, as GNU Radio's block does
                i, self._need = starts.pop(0)
                self._parts = []
            k = min(self._need, len(items) - i)
            self._parts.append(items[i:i + k].copy())
            self._need -= k
            i += k
            if self._need == 0 and self._parts:
                out.append((None, np.concatenate(self._parts)))
                self._parts = []
        self.nitems_read += len(items)
        return out


def stream_chunks(source, chunk=256 * 16, idle_gap=0, max_chunks=None):
    """Drain a pdu_to_tagged_stream into chunks that are multiples of 256 items (what flex_rx.work accepts); the last
    chunk is zero padded, and `idle_gap` zero items follow every drained burst (a receiver needs the samples after a
    frame to finish it: liquid's flexframesync completes a frame on its last sample, the batch receiver when the call
    holding that sample returns)."""
    if chunk <= 0 or chunk % 256:
        raise ValueError("chunk must be a positive multiple of 256 items")
    n = 0
    while source.pending() and (max_chunks is None or n < max_chunks):
        items, _tags = source.work(chunk)
        buf = np.zeros(chunk, source.dtype)
        buf[:len(items)] = items
        n += 1
        yield buf
    for _ in range((idle_gap + chunk - 1) // chunk):
        yield np.zeros(chunk, source.dtype)
