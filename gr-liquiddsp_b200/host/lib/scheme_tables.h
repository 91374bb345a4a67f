// Block-API index <-> liquid-dsp enum tables shared by flex_tx and flex_rx.
// The numbering is the reference's: modulation 0..10 (lib/flex_tx_impl.cc:77-115,
// lib/flex_rx_impl.cc:139-178), inner code 0..6 -> fec0 (:119-147 / :107-135; v27p34 is skipped),
// outer code 0..7 -> fec1 (:150-181 / :75-104).  cognitive_engine.py numbers its 616
// configurations from these indices (python/cognitive_engine.py:87), so they must not move.
#pragma once
#include "../../../include/lqb200.h"

namespace gr { namespace liquiddsp { namespace tables {

// Entries past the reference's ranges are ADDITIVE extensions: schemes liquid-dsp and the
// kernels support but the reference's switch ladders never offered.  0..10 / 0..6 / 0..7 are the reference's and do
// not move; a reference transmitter can never produce the extended values, so a reference receiver's "-1" for them
// is unreachable in a pure-reference system.
static const unsigned kNumModRef = 11, kNumInnerRef = 7, kNumOuterRef = 8;
static const unsigned kModulation[13] = { LQB_MODEM_PSK2, LQB_MODEM_PSK4, LQB_MODEM_PSK8, LQB_MODEM_PSK16,
                                          LQB_MODEM_DPSK2, LQB_MODEM_DPSK4, LQB_MODEM_DPSK8, LQB_MODEM_ASK4,
                                          LQB_MODEM_QAM16, LQB_MODEM_QAM32, LQB_MODEM_QAM64,
                                          /* 11, 12 (extension) */ LQB_MODEM_QAM128, LQB_MODEM_QAM256 };
static const unsigned kInner[15] = { LQB_FEC_NONE, LQB_FEC_CONV_V27, LQB_FEC_CONV_V27P23, LQB_FEC_CONV_V27P45,
                                     LQB_FEC_CONV_V27P56, LQB_FEC_CONV_V27P67, LQB_FEC_CONV_V27P78,
                                     /* 7 (extension): the rate the reference skips */ LQB_FEC_CONV_V27P34,
                                     /* 8 .. 14 (extension): K = 9 family */ LQB_FEC_CONV_V29, LQB_FEC_CONV_V29P23,
                                     LQB_FEC_CONV_V29P34, LQB_FEC_CONV_V29P45, LQB_FEC_CONV_V29P56, LQB_FEC_CONV_V29P67,
                                     LQB_FEC_CONV_V29P78 };
static const unsigned kOuter[8] = { LQB_FEC_NONE, LQB_FEC_GOLAY2412, LQB_FEC_RS_M8, LQB_FEC_HAMMING74,
                                    LQB_FEC_HAMMING128, LQB_FEC_SECDED2216, LQB_FEC_SECDED3932, LQB_FEC_SECDED7264 };

template <unsigned N> inline int index_of(const unsigned (&tab)[N], unsigned scheme)
{
    for (unsigned i = 0; i < N; ++i) if (tab[i] == scheme) return (int)i;
    return -1;
}

}}}
