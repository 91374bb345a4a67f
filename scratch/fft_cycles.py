import ctypes as C, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gr-liquiddsp_b200", "python"))
from liquiddsp import capi
L = capi.lib()
L.lqb_dbg_fft512_cycles.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
for warps, ctas in ((1, 1), (4, 1), (1, 148 * 3), (4, 148 * 3)):
    out = (C.c_longlong * ctas)()
    for _ in range(2):
        assert L.lqb_dbg_fft512_cycles(warps, ctas, 16, out) == 0
    print("warps/CTA %d, CTAs %d: %.0f cycles per FFT-512 (one warp each, load + transform + store)" % (warps, ctas, sum(out) / ctas / 16))
