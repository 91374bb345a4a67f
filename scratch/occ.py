import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gr-liquiddsp_b200", "python"))
import numpy as np
from liquiddsp import capi
rx = capi.Rx(4)
rx.execute([np.zeros(5000, np.complex64)] * 4)
print("ok", rx.counts())
