"""The C-ABI library loads without a GPU, exports every symbol include/lqb200.h declares,
fails loudly (no CPU fallback), and its host-side tables equal the oracle's."""
import os
import re

import numpy as np
import pytest

import lqo_py as o
from liquiddsp import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "lqb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lqb_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    L = capi.lib()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(capi.DECLARED_SYMBOLS) == names


def test_every_liquid_signature_symbol_is_exported():
    """include/lqb200_liquid.h (liquid-dsp's own names) is part of the boundary too."""
    txt = open(os.path.join(ROOT, "include", "lqb200_liquid.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b((?:flexframesync|flexframegen|flexframegenprops|qdetector_cccf|msequence)_[a-z0-9_]+)\s*\(", txt)))
    assert len(names) >= 20 and "flexframesync_flush" in names
    L = capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure_not_fallback():
    L = capi.lib()
    if L.lqb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.LqbError):
        capi.Rx(1)
    with pytest.raises(capi.LqbError):
        capi.Det(1)
    assert b"no CPU fallback" in L.lqb_last_error()


def test_product_tables_equal_oracle_tables():
    Lo = o.lib()
    for beta in (0.25, 0.3):
        h = np.zeros(32, np.float32)
        Lo.lqo_interp_taps(2, 7, np.float32(beta), o._ptr(h))
        assert np.array_equal(h[:30], capi.tab_interp_taps(beta))
    b = np.zeros((32, 28), np.float32)
    Lo.lqo_pfb_rnyquist(32, 2, 7, np.float32(0.3), o._ptr(b))
    assert np.array_equal(b, capi.tab_pfb_banks(0.3))
    st = np.ctypeslib.as_array(Lo.lqo_nco_sintab(), (1024,))
    assert np.array_equal(st, capi.tab_nco_sintab())


@pytest.mark.parametrize("ms,f0,f1,n", [(2, 1, 1, 256), (2, 11, 27, 1500), (27, 1, 1, 1500), (29, 20, 7, 999), (1, 15, 10, 1)])
def test_packet_length_arithmetic_matches_oracle(ms, f0, f1, n):
    Lo = o.lib()
    enc, nsym = capi.tab_packet_len(n, 5, f0, f1, ms)
    assert enc == Lo.lqo_packetizer_enc_len(n, 5, f0, f1)
    assert nsym == Lo.lqo_qpm_frame_len(n, 5, f0, f1, ms)


def test_unsupported_scheme_is_an_error_not_a_guess():
    with pytest.raises(capi.LqbError):
        capi.tab_packet_len(10, 5, 1, 1, 45)      # ARB16OPT: not implemented


@pytest.mark.parametrize("n", [2, 27, 54, 100, 407, 3458])
def test_interleaver_bit_permutation_matches_oracle(n):
    """The soft-decision path deinterleaves one byte per coded bit through this permutation (host-built)."""
    Lo = o.lib()
    Lo.lqo_deinterleave_bit_perm.argtypes = [__import__("ctypes").c_uint, __import__("ctypes").c_void_p]
    ref = np.zeros(8 * n, np.uint32)
    Lo.lqo_deinterleave_bit_perm(n, o._ptr(ref))
    assert np.array_equal(capi.tab_ilv_bit_perm(n), ref)
    assert np.array_equal(np.sort(ref), np.arange(8 * n))          # a permutation
    # and it really is the byte deinterleaver: permuting the bits of a random block the same way gives lqo_deinterleave
    blk = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    bits = np.unpackbits(blk)
    want = blk.copy()
    Lo.lqo_deinterleave(o._ptr(want), n, 4)
    assert np.array_equal(np.packbits(bits[ref]), want)


@pytest.mark.parametrize("nb", [2, 4, 8])
def test_secded_tables_of_the_library_equal_the_oracles(nb):
    import ctypes as C
    Lo = o.lib()
    Lo.lqo_secded_columns.argtypes = [C.c_uint, C.c_void_p]
    ref = np.zeros(8 * nb, np.uint8)
    Lo.lqo_secded_columns(nb, o._ptr(ref))
    assert np.array_equal(capi.tab_secded_columns(nb), ref)
