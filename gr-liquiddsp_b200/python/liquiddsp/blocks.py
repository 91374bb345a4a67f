"""Python faces of the three blocks with the reference's names, constructor arguments, message ports
and dict keys (what `import liquiddsp` gives a flowgraph through SWIG in the reference:
/root/reference/swig/liquiddsp_swig.i:19-26, python/__init__.py:4).  They run without GNU Radio --
a tiny message-port layer stands in for gr.basic_block -- and call the same C-ABI the C++ blocks
call.  No CPU fallback: constructing a block without a CUDA device raises."""
import numpy as np

from . import capi

# flex_tx index -> liquid enum.  0..10 / 0..6 / 0..7 are the reference's numbering (lib/flex_tx_impl.cc:77-181) and do not
# move; the entries after them are additive extensions: QAM128 / QAM256, the v27p34 rate the
# reference skips, and the K = 9 family.  A reference transmitter cannot produce them.
N_MOD_REF, N_INNER_REF, N_OUTER_REF = 11, 7, 8
MODULATION = [1, 2, 3, 4, 9, 10, 11, 18, 27, 28, 29] + [30, 31]
INNER_CODE = [1, 11, 15, 17, 18, 19, 20] + [16, 12, 21, 22, 23, 24, 25, 26]
OUTER_CODE = [1, 7, 27, 4, 6, 8, 9, 10]
CRC_24 = 5


def _index(table, scheme, what):
    if scheme in table:
        return table.index(scheme)
    print("Unsupported Received %s." % what)
    return -1


class _MsgBlock(object):
    """Just enough of gr.basic_block's message API for flowgraph-free use and tests."""

    def __init__(self, name):
        self._name = name
        self._out = {}
        self._in = {}

    def name(self):
        return self._name

    def message_port_register_out(self, port):
        self._out[port] = []

    def message_port_register_in(self, port):
        self._in.setdefault(port, None)

    def set_msg_handler(self, port, fn):
        self._in[port] = fn

    def message_ports_out(self):
        return list(self._out)

    def message_ports_in(self):
        return list(self._in)

    def msg_connect(self, port, dst_block, dst_port):
        self._out[port].append((dst_block, dst_port))

    def message_port_pub(self, port, msg):
        for blk, p in self._out[port]:
            blk.post(p, msg)

    def post(self, port, msg):
        self._in[port](msg)


class sink(_MsgBlock):
    """Collects every message posted to port 'in' (test helper)."""

    def __init__(self):
        _MsgBlock.__init__(self, "sink")
        self.msgs = []
        self.message_port_register_in("in")
        self.set_msg_handler("in", self.msgs.append)


class flex_tx(_MsgBlock):
    """flex_tx(modulation, inner_code, outer_code): PDU (meta, bytes) in on 'pdus' -> (None, complex64 frame) out
    on 'pdus'; dict on 'configuration' re-selects the schemes (lib/flex_tx_impl.cc:42-65,183-209)."""

    def __init__(self, modulation, inner_code, outer_code, device=0):
        _MsgBlock.__init__(self, "flex_tx")
        self._tx = capi.Tx(device=device)
        self._check = CRC_24
        self._header = np.zeros(14, np.uint8)
        self.set_inner_code(inner_code)
        self.set_outer_code(outer_code)
        self.set_modulation(modulation)
        self.message_port_register_out("pdus")
        self.message_port_register_in("pdus")
        self.set_msg_handler("pdus", self.send_pkt)
        self.message_port_register_in("configuration")
        self.set_msg_handler("configuration", self.configure)
        self.num_frames = 0

    def set_modulation(self, modulation):
        if 0 <= modulation < len(MODULATION):
            self._ms = MODULATION[modulation]
        else:
            print("Unsupported Modulation Defaulting to BPSK.")
            self._ms = MODULATION[0]

    def set_inner_code(self, inner_code):
        if 0 <= inner_code < len(INNER_CODE):
            self._fec0 = INNER_CODE[inner_code]
        else:
            print("Unsupported FEC Defaulting to none.")
            self._fec0 = INNER_CODE[0]

    def set_outer_code(self, outer_code):
        if 0 <= outer_code < len(OUTER_CODE):
            self._fec1 = OUTER_CODE[outer_code]
        else:
            print("Unsupported FEC Defaulting to none.")
            self._fec1 = OUTER_CODE[0]

    def configure(self, cfg):
        if "modulation" in cfg:
            self.set_modulation(int(cfg["modulation"]))
        if "inner_code" in cfg:
            self.set_inner_code(int(cfg["inner_code"]))
        if "outer_code" in cfg:
            self.set_outer_code(int(cfg["outer_code"]))

    def send_pkt(self, pdu):
        _meta, data = pdu
        payload = np.frombuffer(bytes(bytearray(data)), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.astype(np.uint8)
        frame = self._tx.assemble([(self._ms, self._check, self._fec0, self._fec1)], [payload], [self._header])[0]
        self.num_frames += 1
        self.message_port_pub("pdus", (None, frame))

    def work(self, *_):
        raise RuntimeError("This is not a stream block.")


class flex_rx(_MsgBlock):
    """flex_rx(): feed complex64 samples to work() in multiples of 256; per completed frame publishes
    'constellation' (None, complex64[]), then -- header valid only -- 'payload_data' (None, bytes) and
    'packet_info' {header_valid, payload_valid, modulation, inner_code, outer_code}
    (lib/flex_rx_impl.cc:204-254)."""

    d_inbuf_len = 256

    def __init__(self, n_channels=1, device=0):
        _MsgBlock.__init__(self, "flex_rx")
        self._rx = capi.Rx(n_channels, device=device)
        self._n = n_channels
        for p in ("constellation", "payload_data", "packet_info"):
            self.message_port_register_out(p)
        self.num_frames = 0
        self.last_stats = []

    def output_multiple(self):
        return self.d_inbuf_len

    def work(self, *channels):
        chunks = [np.ascontiguousarray(c, dtype=np.complex64) for c in channels]
        if any(len(c) % self.d_inbuf_len for c in chunks):
            raise ValueError("flex_rx.work needs a multiple of 256 items")
        self._rx.execute(chunks)
        self.last_stats = self._rx.poll()
        for f in self.last_stats:
            if f["flags"] & 1:            # longer than the receive buffer: announced, nothing to publish
                print("flex_rx: dropped a frame of %d payload bytes (longer than the receive buffer)" % f["payload_len"])
                continue
            self.message_port_pub("constellation", (None, f["framesyms"]))
            if not f["header_valid"]:
                continue
            self.message_port_pub("payload_data", (None, f["payload"]))
            self.message_port_pub("packet_info", {
                "header_valid": 1, "payload_valid": int(f["payload_valid"]),
                "modulation": _index(MODULATION, f["mod_scheme"], "Modulation Defaulting to BPSK"),
                "inner_code": _index(INNER_CODE, f["fec0"], "FEC Defaulting to none"),
                "outer_code": _index(OUTER_CODE, f["fec1"], "FEC defaulting to none")})
            self.num_frames += 1
        return len(chunks[0]) if chunks else 0


class frame_detector_cc(_MsgBlock):
    """frame_detector_cc(): work(in) returns the input unchanged and counts preambles
    (lib/frame_detector_cc_impl.cc:67-97)."""

    def __init__(self, device=0):
        _MsgBlock.__init__(self, "frame_detector_cc")
        self._det = capi.Det(1, device=device)
        self.num_frames = 0
        self.detections = []

    def work(self, samples):
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        self._det.execute([x])
        for d in self._det.poll():
            print("Detected %d frames!" % self.num_frames)
            self.num_frames += 1
            self.detections.append(d)
        return x.copy()
