// lqb_lens.h -- scheme properties and packet length arithmetic, usable on host and device.
// Restates liquid-dsp's fec_get_enc_msg_length / packetizer_compute_enc_msg_len /
// qpacketmodem frame length; scheme ids are the liquid enums the
// reference maps at lib/flex_tx_impl.cc:77-181.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define LQB_HD __host__ __device__ __forceinline__
#else
#define LQB_HD inline
#endif

namespace lqb {

LQB_HD bool modem_supported_hd(unsigned ms) { return (ms >= 1 && ms <= 31) || ms == 39 || ms == 40; }

LQB_HD unsigned modem_bps_hd(unsigned ms)
{
    if (ms >= 1 && ms <= 8) return ms;
    if (ms >= 9 && ms <= 16) return ms - 8;
    if (ms >= 17 && ms <= 24) return ms - 16;
    if (ms >= 25 && ms <= 31) return ms - 23;
    if (ms == 39) return 1;
    if (ms == 40) return 2;
    return 0;
}

LQB_HD unsigned crc_len_hd(unsigned c)
{
    return (c == 2 || c == 3) ? 1u : (c == 4) ? 2u : (c == 5) ? 3u : (c == 6) ? 4u : 0u;
}

LQB_HD bool conv_params_hd(unsigned fs, unsigned &K, unsigned &P)
{
    if (fs == 11) { K = 7; P = 1; return true; }
    if (fs == 12) { K = 9; P = 1; return true; }
    if (fs >= 15 && fs <= 20) { K = 7; P = fs - 13; return true; }
    if (fs >= 21 && fs <= 26) { K = 9; P = fs - 19; return true; }
    return false;
}

LQB_HD bool fec_supported_hd(unsigned fs)
{
    unsigned K, P;
    return (fs >= 1 && fs <= 10) || fs == 27 || conv_params_hd(fs, K, P);
}

LQB_HD unsigned blk_len_hd(unsigned n, unsigned m, unsigned k)
{
    unsigned bits = 8 * n, blocks = (bits + m - 1) / m;
    return (blocks * k + 7) / 8;
}

LQB_HD unsigned fec_enc_len_hd(unsigned fs, unsigned n)
{
    unsigned K = 0, P = 0;
    switch (fs) {
    case 1: return n;
    case 2: return 3 * n;
    case 3: return 5 * n;
    case 4: return blk_len_hd(n, 4, 7);
    case 5: return 2 * n;
    case 6: return blk_len_hd(n, 8, 12);
    case 7: return blk_len_hd(n, 12, 24);
    case 8: return n + (n + 1) / 2;
    case 9: return n + (n + 3) / 4;
    case 10: return n + (n + 7) / 8;
    case 27: {
        if (n == 0) return 0;
        unsigned blocks = (n + 222) / 223, dec = (n + blocks - 1) / blocks;
        return blocks * (dec + 32);
    }
    default:
        if (!conv_params_hd(fs, K, P)) return 0;
        if (P == 1) return 2 * n + 2;
        {
            unsigned nb = 8 * n + K - 1, out = nb + (nb + P - 1) / P;
            return (out + 7) / 8;
        }
    }
}

LQB_HD unsigned packetizer_enc_len_hd(unsigned n, unsigned check, unsigned fec0, unsigned fec1)
{
    return fec_enc_len_hd(fec1, fec_enc_len_hd(fec0, n + crc_len_hd(check)));
}

LQB_HD unsigned qpm_frame_len_hd(unsigned n, unsigned check, unsigned fec0, unsigned fec1, unsigned ms)
{
    unsigned bps = modem_bps_hd(ms);
    if (!bps) return 0;
    unsigned bits = 8 * packetizer_enc_len_hd(n, check, fec0, fec1);
    return (bits + bps - 1) / bps;
}

}  // namespace lqb
