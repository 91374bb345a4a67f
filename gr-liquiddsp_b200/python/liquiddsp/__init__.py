"""liquiddsp -- Python face of the B200-native packet PHY, mirroring the reference package layout
(/root/reference/python/__init__.py:4,9: `from liquiddsp_swig import *` then the pure-Python blocks)."""
from . import capi  # noqa: F401
