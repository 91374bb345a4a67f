"""liquiddsp -- Python face of the B200-native packet PHY, mirroring the reference package layout
(/root/reference/python/__init__.py:4,9: the SWIG blocks first, then the pure-Python blocks)."""
from . import capi  # noqa: F401
from .blocks import flex_rx, flex_tx, frame_detector_cc  # noqa: F401
from . import policy, sharding, bulk, replay, adapters  # noqa: F401
from .adapters import pdu_to_tagged_stream, tagged_stream_to_pdu  # noqa: F401
