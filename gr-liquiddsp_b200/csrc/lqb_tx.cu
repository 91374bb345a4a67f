// lqb_tx.cu -- batched frame generator (flexframegen) on the GPU and its C-ABI.
//
// One CTA assembles one frame: header (CRC-32, whitening, SECDED(72,64), Hamming(8,4),
// interleaving, QPSK, pilots), payload (CRC, whitening, fec0, interleave, fec1, interleave,
// bit packing, modulation) and the 2x ARKAISER interpolation -- i.e. what
// flexframegen_assemble + flexframegen_write_samples do in liquid-dsp.
// Reference call sites replaced: lib/flex_tx_impl.cc:51-56 (props/create), :188 (setprops),
// :198-201 (assemble / getframelen / write_samples).
#include "../../include/lqb200.h"
#include "lqb_dev.cuh"
#include "lqb_tables.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

extern "C" void lqb_internal_set_error(const char *msg);

namespace lqb {

struct TxTables {
    float    h[32];            // interpolator taps (30 used)
    float2   preamble[64];
    float2   pilots[16];
    float2   psk_map[8 * 256];
    uint32_t crc_tab[8][256];
    uint16_t ilv54[4][28], ilv27[4][16];
    uint16_t hperm27f[216], hperm54f[432];   // the header's two interleavers as bit gathers: out bit i = in bit perm[i]
    uint8_t  secded_col[3][64];
    uint8_t  gf_exp[512], gf_log[256], rs_gen[64];
};

struct TxFrame {               // one frame's plan (host-built)
    const unsigned char *pay;  // payload bytes (device: the caller's buffer in place, or the staging buffer)
    float2 *out;               // where the frame's samples go (device: the caller's buffer in place, or the staging buffer)
    unsigned long long pay_off, buf_off, sym_off, out_off;
    unsigned payload_len, check, fec0, fec1, ms, bps;
    unsigned k0, n0, n1, n_sym, buf_len, ilv0_off, ilv1_off, n_samples;
    unsigned char header[16];
};

namespace {

// Threads per frame.  The header encode and the payload check are serial sections (one thread each, side by side);
// the fewer threads wait for them, the more frames an SM keeps in flight to cover them: 64 threads = 32 frames per SM.
#ifndef LQB_TX_THREADS
#define LQB_TX_THREADS 64
#endif
constexpr int kTxThreads = LQB_TX_THREADS;
constexpr int kTxTile = 4 * LQB_TX_THREADS;          // symbols interpolated per pass (two passes of two symbols per thread)
constexpr int kTxBytesWords = (kTxTile + 16) * 8 / 32 + 2;   // words that hold the bits of kTxTile + 16 symbols of up to 8 bits from any bit offset
static_assert(kTxThreads >= 64 && kTxThreads % 32 == 0, "warp 0 encodes the header while warp 1 checks the payload");
constexpr float kPiF = 3.14159274f;
constexpr float kTwoPiF = 6.28318548f;

__device__ __forceinline__ unsigned gray_dec(unsigned s) { unsigned r = s; for (unsigned sh = 1; sh < 32; sh <<= 1) r ^= r >> sh; return r; }
__device__ __forceinline__ unsigned bit_at(const unsigned char *x, long long pos, long long nbits)
{
    return (pos >= 0 && pos < nbits) ? ((x[pos >> 3] >> (7 - (pos & 7))) & 1u) : 0u;
}

__device__ void ilv_forward(unsigned char *x, unsigned n, const unsigned *maps, int tid)
{
    const unsigned n2 = n / 2;
    const unsigned masks[4] = { 0xffu, 0x0fu, 0x55u, 0x33u };
    for (int pass = 0; pass < 4; ++pass) {
        const unsigned *map = maps + (size_t)pass * n2;
        const unsigned mask = masks[pass];
        for (unsigned i = tid; i < n2; i += kTxThreads) {
            unsigned j = map[i], a = x[2 * j + 1], b = x[2 * i];
            x[2 * j + 1] = (unsigned char)((a & ~mask) | (b & mask));
            x[2 * i] = (unsigned char)((a & mask) | (b & ~mask));
        }
        __syncthreads();
    }
}
__device__ void ilv_small(unsigned char *x, const uint16_t *map, unsigned n2, unsigned mask)
{
    for (unsigned i = 0; i < n2; ++i) {
        unsigned j = map[i], a = x[2 * j + 1], b = x[2 * i];
        x[2 * j + 1] = (unsigned char)((a & ~mask) | (b & mask));
        x[2 * i] = (unsigned char)((a & mask) | (b & ~mask));
    }
}

// parity byte of liquid's Hsiao SEC-DED code with nb data bytes: XOR of the columns of the set data bits
__device__ unsigned secded_parity(const TxTables *T, const unsigned char *blk, unsigned nb)
{
    const unsigned char *col = T->secded_col[nb == 2 ? 0 : nb == 4 ? 1 : 2];
    unsigned p = 0;
    for (unsigned bit = 0; bit < nb * 8; ++bit)
        if ((blk[bit >> 3] >> (7 - (bit & 7))) & 1u) p ^= col[bit];
    return p;
}

__constant__ unsigned char c_h84[16] = { 0x00, 0xd2, 0x55, 0x87, 0x99, 0x4b, 0xcc, 0x1e, 0xe1, 0x33, 0xb4, 0x66, 0x78, 0xaa, 0x2d, 0xff };
__constant__ unsigned c_gP[12] = { 0x8ed, 0x1db, 0x3b5, 0x769, 0xed1, 0xda3, 0xb47, 0x68f, 0xd1d, 0xa3b, 0x477, 0xffe };

__device__ __forceinline__ unsigned golay_encode(unsigned m)
{
    unsigned r = 0;
    for (int i = 0; i < 12; ++i) r |= ((unsigned)__popc(m & c_gP[i]) & 1u) << (11 - i);
    return (r << 12) | (m & 0xfffu);
}
__device__ __forceinline__ unsigned h128_encode(unsigned d)
{
    unsigned p = (((unsigned)__popc(d & 0xda) & 1u) << 3) | (((unsigned)__popc(d & 0xb6) & 1u) << 2)
               | (((unsigned)__popc(d & 0x71) & 1u) << 1) | ((unsigned)__popc(d & 0x0f) & 1u);
    return (p << 8) | d;
}

// one FEC stage, src (n bytes) -> dst (enc bytes); whole CTA cooperates
__device__ void fec_encode(const TxTables *T, unsigned fs, unsigned n, unsigned enc, const unsigned char *src, unsigned char *dst, int tid)
{
    switch (fs) {
    case 1: for (unsigned i = tid; i < n; i += kTxThreads) dst[i] = src[i]; break;
    case 2: case 3: {
        const unsigned reps = fs == 2 ? 3u : 5u;
        for (unsigned i = tid; i < n * reps; i += kTxThreads) dst[i] = src[i % n];
        break;
    }
    case 4:   // Hamming(7,4), bit packed: output bit pos -> codeword pos/7 (two per input byte)
        for (unsigned o = tid; o < enc; o += kTxThreads) {
            unsigned v = 0;
            for (unsigned b = 0; b < 8; ++b) {
                unsigned pos = 8 * o + b, cw = pos / 7, k = pos % 7, bit = 0;
                if (cw < 2 * n) {
                    unsigned nib = (cw & 1u) ? (src[cw >> 1] & 15u) : (src[cw >> 1] >> 4);
                    bit = ((c_h84[nib] >> 1) >> (6 - k)) & 1u;
                }
                v = (v << 1) | bit;
            }
            dst[o] = (unsigned char)v;
        }
        break;
    case 5:
        for (unsigned i = tid; i < n; i += kTxThreads) { dst[2 * i] = c_h84[src[i] >> 4]; dst[2 * i + 1] = c_h84[src[i] & 15u]; }
        break;
    case 6:
        for (unsigned o = tid; o < enc; o += kTxThreads) {
            unsigned v = 0;
            for (unsigned b = 0; b < 8; ++b) {
                unsigned pos = 8 * o + b, cw = pos / 12, k = pos % 12, bit = 0;
                if (cw < n) bit = (h128_encode(src[cw]) >> (11 - k)) & 1u;
                v = (v << 1) | bit;
            }
            dst[o] = (unsigned char)v;
        }
        break;
    case 7: {
        const unsigned groups = n / 3, rem = n % 3;
        for (unsigned g = tid; g < groups; g += kTxThreads) {
            const unsigned char *s = src + 3 * g;
            unsigned v0 = golay_encode(((unsigned)s[0] << 4) | (s[1] >> 4)), v1 = golay_encode((((unsigned)s[1] & 15u) << 8) | s[2]);
            unsigned char *d = dst + 6 * g;
            d[0] = (unsigned char)(v0 >> 16); d[1] = (unsigned char)(v0 >> 8); d[2] = (unsigned char)v0;
            d[3] = (unsigned char)(v1 >> 16); d[4] = (unsigned char)(v1 >> 8); d[5] = (unsigned char)v1;
        }
        if ((unsigned)tid < rem) {
            unsigned v = golay_encode(src[3 * groups + tid]);
            unsigned char *d = dst + 6 * groups + 3 * tid;
            d[0] = (unsigned char)(v >> 16); d[1] = (unsigned char)(v >> 8); d[2] = (unsigned char)v;
        }
        break;
    }
    case 8: case 9: case 10: {
        const unsigned nb = fs == 8 ? 2u : fs == 9 ? 4u : 8u;
        const unsigned blocks = (n + nb - 1) / nb;
        for (unsigned b = tid; b < blocks; b += kTxThreads) {
            unsigned r = (n - b * nb >= nb) ? nb : (n - b * nb);
            unsigned char blk[8];
            for (unsigned q = 0; q < nb; ++q) blk[q] = q < r ? src[b * nb + q] : 0;
            unsigned char *d = dst + b * (nb + 1);
            d[0] = (unsigned char)secded_parity(T, blk, nb);
            for (unsigned q = 0; q < r; ++q) d[1 + q] = blk[q];
        }
        break;
    }
    case 27: {   // RS(255,223): one warp per block, lane j keeps parity byte j of the LFSR
        const int warp = tid >> 5, lane = tid & 31;
        const unsigned blocks = (n + 222) / 223, dec_block = (n + blocks - 1) / blocks, enc_block = dec_block + 32;
        for (unsigned b = warp; b < blocks; b += kTxThreads / 32) {
            const unsigned n0 = b * dec_block, take = (n - n0 >= dec_block) ? dec_block : (n - n0);
            unsigned par = 0;
            const unsigned g = T->rs_gen[31 - lane];
            for (unsigned i = 0; i < dec_block; ++i) {
                const unsigned dbyte = i < take ? src[n0 + i] : 0u;
                const unsigned fb = dbyte ^ __shfl_sync(0xffffffffu, par, 0);
                unsigned nxt = __shfl_down_sync(0xffffffffu, par, 1);
                if (lane == 31) nxt = 0;
                par = nxt ^ ((fb && g) ? T->gf_exp[T->gf_log[fb] + T->gf_log[g]] : 0u);
                if (lane == 0) dst[b * enc_block + i] = (unsigned char)dbyte;
            }
            dst[b * enc_block + dec_block + lane] = (unsigned char)par;
        }
        break;
    }
    default: {   // convolutional, thread per output byte
        unsigned K = 7, P = 1, poly0 = 0x6d, poly1 = 0x4f, keep0 = 1, keep1 = 1;
        const unsigned k27[6][2] = { { 0x3, 0x1 }, { 0x3, 0x5 }, { 0xf, 0x1 }, { 0xb, 0x15 }, { 0x17, 0x29 }, { 0x2f, 0x51 } };
        const unsigned k29[6][2] = { { 0x3, 0x1 }, { 0x7, 0x1 }, { 0xd, 0x3 }, { 0xb, 0x15 }, { 0x1b, 0x25 }, { 0x6b, 0x15 } };
        if (fs == 12 || (fs >= 21 && fs <= 26)) { K = 9; poly0 = 0x1af; poly1 = 0x11d; }
        if (fs >= 15 && fs <= 20) { P = fs - 13; keep0 = k27[fs - 15][0]; keep1 = k27[fs - 15][1]; }
        if (fs >= 21 && fs <= 26) { P = fs - 19; keep0 = k29[fs - 21][0]; keep1 = k29[fs - 21][1]; }
        // (column,row) of the q-th kept bit inside one puncturing period
        unsigned char kc[16], kr[16];
        unsigned per = 0;
        for (unsigned c = 0; c < P; ++c) {
            if ((keep0 >> c) & 1u) { kc[per] = (unsigned char)c; kr[per] = 0; ++per; }
            if ((keep1 >> c) & 1u) { kc[per] = (unsigned char)c; kr[per] = 1; ++per; }
        }
        const long long nbits = 8ll * n, T_steps = nbits + K - 1;
        const unsigned long long total = (unsigned long long)(T_steps / P) * per;   // plus a partial period below
        unsigned long long out_bits = total;
        for (unsigned c = 0; c < (unsigned)(T_steps % P); ++c) out_bits += ((keep0 >> c) & 1u) + ((keep1 >> c) & 1u);
        for (unsigned o = tid; o < enc; o += kTxThreads) {
            unsigned v = 0;
            for (unsigned b = 0; b < 8; ++b) {
                const unsigned long long m = 8ull * o + b;
                unsigned bit = 0;
                if (m < out_bits) {
                    const unsigned long long q = m / per;
                    const unsigned w = (unsigned)(m % per);
                    const long long t = (long long)(q * P + kc[w]);
                    unsigned sr = 0;
                    for (unsigned k = 0; k < K; ++k) sr |= bit_at(src, t - k, nbits) << k;
                    bit = (unsigned)__popc(sr & (kr[w] ? poly1 : poly0)) & 1u;
                }
                v = (v << 1) | bit;
            }
            dst[o] = (unsigned char)v;
        }
        break;
    }
    }
    __syncthreads();
}

__device__ float2 modulate(const TxTables *T, unsigned ms, unsigned bps, unsigned s)
{
    const unsigned M = 1u << bps;
    if (ms >= 1 && ms <= 8) return T->psk_map[(bps - 1) * 256 + s];
    if (ms >= 17 && ms <= 24) {
        const float c[9] = { 0, 1.0f, 5.0f, 21.0f, 85.0f, 341.0f, 1365.0f, 5461.0f, 21845.0f };
        const float alpha = __fdiv_rn(1.0f, __fsqrt_rn(c[bps]));
        return make_float2(__fmul_rn((float)(2 * (int)gray_dec(s) - (int)M + 1), alpha), 0.0f);
    }
    if (ms >= 25 && ms <= 31) {
        const float c[9] = { 0, 0, 2.0f, 6.0f, 10.0f, 26.0f, 42.0f, 106.0f, 170.0f };
        const float alpha = __fdiv_rn(1.0f, __fsqrt_rn(c[bps]));
        const unsigned m_i = (bps + 1) >> 1, m_q = bps >> 1;
        const unsigned si = gray_dec(s >> m_q), sq = gray_dec(s & ((1u << m_q) - 1u));
        return make_float2(__fmul_rn((float)(2 * (int)si - (int)(1u << m_i) + 1), alpha),
                           __fmul_rn((float)(2 * (int)sq - (int)(1u << m_q) + 1), alpha));
    }
    if (ms == 39) return make_float2(s ? -1.0f : 1.0f, 0.0f);
    return make_float2((s & 1u) ? -0.707106769f : 0.707106769f, (s & 2u) ? -0.707106769f : 0.707106769f);
}

// 20 CTAs per SM is what the 11 KB of shared memory allow; asking for it caps the registers at 48 (56 without: 18 CTAs
// per SM and 3.08 waves for 8192 frames instead of 2.77).  Measured: 0.241 -> 0.231 ms (profiles/r02_notes.md v26).
#ifndef LQB_TX_CTAS_PER_SM
#define LQB_TX_CTAS_PER_SM 20
#endif
__global__ void __launch_bounds__(kTxThreads, LQB_TX_CTAS_PER_SM)
k_tx(const TxTables *T, const TxFrame *frames, unsigned char *bufA, unsigned char *bufB, const unsigned *ilv, float2 *syms)
{
    __shared__ unsigned char hb[64], hd[32], he[32], hraw[64];  // header: 54 coded bytes; 24 data bytes; 27 SECDED bytes; 54 before the last interleaver
    __shared__ unsigned crc4[4][256];                 // slicing-by-4 tables of the frame's check
    __shared__ float2 smap[256];                      // the frame's symbol map (the scale's divide and square root once per point)
    const int tid = threadIdx.x;
    const TxFrame &f = frames[blockIdx.x];
    unsigned char *A = bufA + f.buf_off, *B = bufB + f.buf_off;
    // the frame's symbols in one array: [14 zeros | 64 preamble | 231 header | n_sym payload | 16 zeros], so that the
    // interpolator reads a plain window (it used to pick the source of every tap with a three-way branch: 375
    // instructions per output sample, profiles/r01_notes.md v23)
    // The symbols never exist in global memory: they are made tile by tile into shared memory (preamble from the table,
    // header from `hsym`, payload from the encoded bytes through the symbol map) and interpolated from there.  Only DPSK,
    // whose modulator carries phase memory from symbol to symbol, still writes its payload symbols to `psym` first.
    __shared__ float2 hsym[232];
    __shared__ float2 tile[kTxTile + 16 + (kTxTile + 16) / 16 + 1];   // entry i at i + i / 16: the interpolator's stride-2 reads (8-byte words) would otherwise collide two by two
    __shared__ unsigned abytes[2][kTxBytesWords];          // the encoded bytes one tile of symbols is cut from (bps <= 8), two tiles in rotation
    float2 *psym = syms + f.sym_off;
    if (f.check >= 3 && f.check <= 6) {
        const unsigned *tab = T->crc_tab[f.check];
        for (int i = tid; i < 256; i += kTxThreads) {
            const unsigned t0 = tab[i];
            const unsigned t1 = (t0 >> 8) ^ tab[t0 & 0xffu], t2 = (t1 >> 8) ^ tab[t1 & 0xffu], t3 = (t2 >> 8) ^ tab[t2 & 0xffu];
            crc4[0][i] = t0; crc4[1][i] = t1; crc4[2][i] = t2; crc4[3][i] = t3;
        }
    }

    if (!(f.ms >= 9 && f.ms <= 16))
        for (unsigned i = tid; i < (1u << f.bps); i += kTxThreads) smap[i] = modulate(T, f.ms, f.bps, i);

    // ---------------- payload bytes into the work buffer
    const unsigned plen = f.payload_len, cl = f.k0 - plen;
    const unsigned char *pay = f.pay;
    // payload -> work buffer, whitened on the way (the check below reads the caller's bytes, not these)
    if ((reinterpret_cast<uintptr_t>(pay) & 3u) == 0) {
        const unsigned *p4 = reinterpret_cast<const unsigned *>(pay);
        unsigned *a4 = reinterpret_cast<unsigned *>(A);                           // buf_off is a multiple of 16
        for (unsigned i = tid; i < (plen >> 2); i += kTxThreads) a4[i] = p4[i] ^ 0xc58b6ab4u;      // b4 6a 8b c5, little endian
        for (unsigned i = (plen & ~3u) + tid; i < plen; i += kTxThreads)
            A[i] = pay[i] ^ (unsigned char)(0xc58b6ab4u >> (8u * (i & 3u)));
    } else {
        for (unsigned i = tid; i < plen; i += kTxThreads) A[i] = pay[i] ^ (unsigned char)(0xc58b6ab4u >> (8u * (i & 3u)));
    }
    __syncthreads();                                      // (crc4 / smap tables are complete)
    // ---------------- the two serial sections run side by side: header on warp 0, payload check on warp 1
    if (tid < 32) {
        // header on warp 0, every step across the lanes (it used to be one thread walking 160 interleaver swaps and 192
        // parity bits through local memory: ~50 us per frame)
        const unsigned lane = (unsigned)tid;
        if (lane < 14) hd[lane] = f.header[lane];
        if (lane == 14) {
            hd[14] = 102; hd[15] = (unsigned char)(f.payload_len >> 8); hd[16] = (unsigned char)f.payload_len;
            hd[17] = (unsigned char)f.ms; hd[18] = (unsigned char)(((f.check & 7u) << 5) | (f.fec0 & 0x1fu)); hd[19] = (unsigned char)(f.fec1 & 0x1fu);
        }
        __syncwarp();
        if (lane == 0) {
            unsigned key = 0xffffffffu;
            for (int i = 0; i < 20; ++i) key = (key >> 8) ^ T->crc_tab[6][(key ^ hd[i]) & 0xffu];
            key = ~key;
            hd[20] = (unsigned char)(key >> 24); hd[21] = (unsigned char)(key >> 16); hd[22] = (unsigned char)(key >> 8); hd[23] = (unsigned char)key;
        }
        __syncwarp();
        if (lane < 24) hd[lane] ^= (unsigned char)(lane & 3u) == 0 ? 0xb4u : (lane & 3u) == 1 ? 0x6au : (lane & 3u) == 2 ? 0x8bu : 0xc5u;
        __syncwarp();
        // SECDED(72,64): parity of block b = XOR of the columns of its set bits; lane handles bits lane and lane + 32
        for (unsigned b = 0; b < 3; ++b) {
            unsigned p = 0;
#pragma unroll
            for (unsigned h2 = 0; h2 < 2; ++h2) {
                const unsigned bit = lane + 32u * h2;
                if ((hd[8 * b + (bit >> 3)] >> (7 - (bit & 7))) & 1u) p ^= T->secded_col[2][bit];
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) p ^= __shfl_xor_sync(0xffffffffu, p, m);
            if (lane == 0) he[9 * b] = (unsigned char)p;
            if (lane < 8) he[9 * b + 1 + lane] = hd[8 * b + lane];
        }
        __syncwarp();
        // interleave(27) as one bit gather, then Hamming(8,4), then interleave(54) as one bit gather
        unsigned char v27 = 0;
        if (lane < 27) {
            unsigned v = 0;
#pragma unroll
            for (unsigned bit = 0; bit < 8; ++bit) {
                const unsigned src = T->hperm27f[8 * lane + bit];
                v |= ((he[src >> 3] >> (src & 7u)) & 1u) << bit;
            }
            v27 = (unsigned char)v;
        }
        __syncwarp();
        if (lane < 27) { hraw[2 * lane] = c_h84[v27 >> 4]; hraw[2 * lane + 1] = c_h84[v27 & 15u]; }
        __syncwarp();
        for (unsigned o = lane; o < 54; o += 32) {
            unsigned v = 0;
#pragma unroll
            for (unsigned bit = 0; bit < 8; ++bit) {
                const unsigned src = T->hperm54f[8 * o + bit];
                v |= ((hraw[src >> 3] >> (src & 7u)) & 1u) << bit;
            }
            hb[o] = (unsigned char)v;
        }
    }
    if (tid >= 32 && tid < 64) {
        // payload check on warp 1, over the caller's bytes.  A CRC is linear over GF(2): lane i runs the table recurrence
        // over its own 1/32 of the message (lane 0 from the all-ones preset, the others from zero); "advance a state by L
        // zero bytes" is a 32 x 32 bit matrix whose column j lane j obtains by running 1 << j through L zero bytes; the
        // lanes' states are then folded in order (Horner), each fold one select and an XOR reduction by shuffle.
        // ~1.1 k warp instructions per 1500-byte frame; the serial slicing-by-4 loop was 6.5 k, a third of the kernel.
        const unsigned lane = (unsigned)tid - 32u;
        unsigned key = 0;
        if (f.check == 2) {
            unsigned sum = 0;
            for (unsigned i = lane; i < plen; i += 32) sum += pay[i];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
            key = (~sum + 1u) & 0xffu;
        } else if (f.check >= 3 && f.check <= 6) {
            const unsigned *tab = crc4[0];
            const unsigned L = max(4u, (((plen + 31u) >> 5) + 3u) & ~3u);          // bytes per lane
            const unsigned F = plen / L, r = plen - F * L;                      // full lanes; bytes of lane F
            unsigned st = lane == 0 ? 0xffffffffu : 0u;
            {
                const unsigned b0 = lane * L, b1 = min(plen, b0 + L);
                for (unsigned i = b0; i < b1; ++i) st = (st >> 8) ^ tab[(st ^ pay[i]) & 0xffu];
            }
            unsigned colL = 1u << lane, colr = 1u << lane;
            for (unsigned i = 0; i < L; ++i) {
                colL = (colL >> 8) ^ tab[colL & 0xffu];
                if (i < r) colr = (colr >> 8) ^ tab[colr & 0xffu];
            }
            auto matvec = [&](unsigned col, unsigned v) {
                unsigned y = ((v >> lane) & 1u) ? col : 0u;
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) y ^= __shfl_xor_sync(0xffffffffu, y, m);
                return y;
            };
            unsigned k;
            if (F == 0) k = __shfl_sync(0xffffffffu, st, 0);                    // (plen < L: everything in lane 0)
            else {
                k = __shfl_sync(0xffffffffu, st, 0);
                for (unsigned i = 1; i < F; ++i) k = matvec(colL, k) ^ __shfl_sync(0xffffffffu, st, (int)i);
                if (r) k = matvec(colr, k) ^ __shfl_sync(0xffffffffu, st, (int)F);
            }
            const unsigned bits = f.check == 3 ? 8u : f.check == 4 ? 16u : f.check == 5 ? 24u : 32u;
            key = (~k) & (bits == 32 ? 0xffffffffu : ((1u << bits) - 1u));
        }
        if (lane == 0)
            for (unsigned i = 0; i < cl; ++i) {           // big endian behind the payload, whitened like the payload
                const unsigned pos = plen + cl - i - 1;
                A[pos] = (unsigned char)((key & 0xffu) ^ (0xc58b6ab4u >> (8u * (pos & 3u))));
                key >>= 8;
            }
    }
    __syncthreads();
    for (int i = tid; i < 231; i += kTxThreads) {
        if ((i & 15) == 0) hsym[i] = T->pilots[i >> 4];
        else {
            int n = i - (i >> 4) - 1;                     // data symbol ordinal
            unsigned s = (hb[n >> 2] >> (6 - 2 * (n & 3))) & 3u;
            hsym[i] = make_float2((s & 1u) ? -0.707106769f : 0.707106769f, (s & 2u) ? -0.707106769f : 0.707106769f);
        }
    }

    __syncthreads();
    // ---------------- fec0 / interleave / fec1 / interleave ("no code" moves nothing: the buffers just keep their roles)
    unsigned char *cur = A, *oth = B;
    if (f.fec0 != 1) {
        fec_encode(T, f.fec0, f.k0, f.n0, cur, oth, tid);
        ilv_forward(oth, f.n0, ilv + f.ilv0_off, tid);
        unsigned char *t_ = cur; cur = oth; oth = t_;
    }
    if (f.fec1 != 1) {
        fec_encode(T, f.fec1, f.n0, f.n1, cur, oth, tid);
        ilv_forward(oth, f.n1, ilv + f.ilv1_off, tid);
        unsigned char *t_ = cur; cur = oth; oth = t_;
    }
    A = cur;                                              // the encoded message

    // ---------------- bits -> symbols -> samples, a tile of kTxTile symbols at a time
    const unsigned bps = f.bps, nbits = 8 * f.n1;
    const bool dpsk = (f.ms >= 9 && f.ms <= 16);
    if (dpsk) {
        // DPSK carries phase memory: phi_i = wrap(phi_{i-1} + inc_i) is a serial float recurrence (the specification's
        // rounding, add then conditional wrap), but only that: the increments (bit extraction, Gray decoding) and the
        // sin / cos of every phase are independent.  All threads make the increments, warp 0 runs the recurrence 32
        // symbols at a time (coalesced load, every lane follows the chain through shuffles and keeps its own phase),
        // all threads take sin / cos.  (One thread doing all three took 0.3 ms for a 4096-symbol frame.)
        const float alpha = __fdiv_rn(kPiF, (float)(1u << bps));
        for (unsigned i = tid; i < f.n_sym; i += kTxThreads) {
            unsigned s = 0;
            for (unsigned b = 0; b < bps; ++b) s = (s << 1) | bit_at(A, (long long)i * bps + b, nbits);
            psym[i].x = __fmul_rn(__fmul_rn((float)gray_dec(s), 2.0f), alpha);
        }
        __syncthreads();
        if (tid < 32) {
            float phi = 0.0f;
            for (unsigned i0 = 0; i0 < f.n_sym; i0 += 32) {
                const unsigned i = i0 + tid;
                const float inc = (i < f.n_sym) ? psym[i].x : 0.0f;
                float mine = 0.0f;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    phi = __fadd_rn(phi, __shfl_sync(0xffffffffu, inc, k));
                    if (phi > kTwoPiF) phi = __fsub_rn(phi, kTwoPiF);
                    if (k == tid) mine = phi;
                }
                if (i < f.n_sym) psym[i].x = mine;
            }
        }
        __syncthreads();
        for (unsigned i = tid; i < f.n_sym; i += kTxThreads) {
            float sn, cs;
            pm_sincosf(psym[i].x, &sn, &cs);
            psym[i] = make_float2(cs, sn);
        }
        __syncthreads();
    }

    // 2x interpolation: out[2t+ph] = sum_n h[ph + 2n] sym[t-n], oldest symbol first.  A thread makes the four samples of
    // two neighbouring symbols from a 16-symbol window; FFMA2 with the tap as the scalar operand is two IEEE fmas, applied
    // in the specification's order (n = 14 first), so the samples are bit-identical to the per-sample form.
    // Symbol index a runs over [14 zeros | 64 preamble | 231 header | n_sym payload | zeros]; tile[j] = symbol a0 + j.
    const int total_syms = 64 + 231 + (int)f.n_sym + 14;
    const int pay0 = 14 + 64 + 231;
    float2 *o = f.out;
    float h[30];
#pragma unroll
    for (int i = 0; i < 30; ++i) h[i] = __ldg(T->h + i);
    const unsigned n1_words = (f.n1 + 3u) >> 2;
    const unsigned *A32 = reinterpret_cast<const unsigned *>(A);              // buf_off is a multiple of 16
    // encoded bytes of the payload symbols [p0, p0 + kTxTile + 16) -> abytes[buf]: whole words from the byte p0 * bps / 8 on
    auto stage_bytes = [&](int a0, int buf) {
        const int p0 = max(a0 - pay0, 0);
        const unsigned w0 = ((unsigned)p0 * bps) >> 5;
        for (int i = tid; i < kTxBytesWords; i += kTxThreads) {
            const unsigned w = w0 + (unsigned)i;
            abytes[buf][i] = (!dpsk && w < n1_words) ? __byte_perm(A32[w], 0u, 0x0123) : 0u;     // MSB-first bit order
        }
    };
    const bool al16 = (reinterpret_cast<uintptr_t>(o) & 15u) == 0;
    __syncthreads();                                      // the encoded message is complete
    stage_bytes(0, 0);
    int buf = 0;
    for (int a0 = 0; a0 < total_syms; a0 += kTxTile, buf ^= 1) {
        __syncthreads();                                  // abytes[buf] is complete; the previous tile has been read
        // ---- fill the tile: symbols a0 .. a0 + kTxTile + 15
        const int p0 = max(a0 - pay0, 0);
        const unsigned wbase = ((unsigned)p0 * bps) >> 5;
        // a tile that lies wholly inside the payload (all but the first two and the last of a frame): no case analysis
        const bool pure = !dpsk && a0 >= pay0 && a0 + kTxTile + 16 <= pay0 + (int)f.n_sym && ((unsigned)(p0 + kTxTile + 16) * bps <= nbits);
        if (pure) {
            const unsigned pos0 = (unsigned)p0 * bps - 32u * wbase;           // bit offset of the tile's first symbol in abytes
            const unsigned *ab = abytes[buf];
#pragma unroll
            for (int k = 0; k < (kTxTile + 16 + kTxThreads - 1) / kTxThreads; ++k) {
                const int j = tid + kTxThreads * k;
                if (j < kTxTile + 16) {
                    const unsigned pos = pos0 + (unsigned)j * bps, wi = pos >> 5;
                    tile[j + (j >> 4)] = smap[__funnelshift_l(ab[wi + 1], ab[wi], pos & 31u) >> (32u - bps)];
                }
            }
        } else {
            // region by region (zeros | preamble | header | payload | zeros), each a plain strided loop over its part of the tile
            constexpr int kJ = kTxTile + 16;
            const int pay1 = pay0 + (int)f.n_sym;
            auto lo = [&](int ra) { return min(max(ra - a0, 0), kJ); };
            for (int j = lo(0) + tid; j < lo(14); j += kTxThreads) tile[j + (j >> 4)] = make_float2(0.0f, 0.0f);
            for (int j = lo(14) + tid; j < lo(14 + 64); j += kTxThreads) tile[j + (j >> 4)] = T->preamble[a0 + j - 14];
            for (int j = lo(14 + 64) + tid; j < lo(pay0); j += kTxThreads) tile[j + (j >> 4)] = hsym[a0 + j - 14 - 64];
            if (dpsk) {
                for (int j = lo(pay0) + tid; j < lo(pay1); j += kTxThreads) tile[j + (j >> 4)] = psym[a0 + j - pay0];
            } else {
                const unsigned *ab = abytes[buf];
                for (int j = lo(pay0) + tid; j < lo(pay1); j += kTxThreads) {
                    // bps <= 8 bits starting at bit i * bps (MSB first); bits past the end of the encoded message read as zero
                    const unsigned i = (unsigned)(a0 + j - pay0), pos = i * bps, wi = (pos >> 5) - wbase;
                    unsigned sv = __funnelshift_l(ab[wi + 1], ab[wi], pos & 31u) >> (32u - bps);
                    if (pos + bps > nbits) sv &= ~((1u << (pos + bps - nbits)) - 1u);
                    tile[j + (j >> 4)] = smap[sv];
                }
            }
            for (int j = lo(pay1) + tid; j < kJ; j += kTxThreads) tile[j + (j >> 4)] = make_float2(0.0f, 0.0f);
        }
        if (a0 + kTxTile < total_syms) stage_bytes(a0 + kTxTile, buf ^ 1);   // the next tile's bytes travel while this one is interpolated
        __syncthreads();
        // ---- interpolate symbols t = a0 .. a0 + kTxTile - 1 (w[q] = sym[t - 14 + q] = tile[t - a0 + q] as a starts 14 early)
        const int t_end = min(a0 + kTxTile, total_syms);
        for (int t = a0 + 2 * tid; t < t_end; t += 2 * kTxThreads) {
            const int j0 = t - a0;
            float2 a0_ = make_float2(0.0f, 0.0f), a1 = a0_, a2 = a0_, a3 = a0_;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float2 x = tile[(j0 + q) + ((j0 + q) >> 4)];
                if (q <= 14) {                            // symbol t: n = 14 - q
                    a0_ = __ffma2_rn(make_float2(h[2 * (14 - q)], h[2 * (14 - q)]), x, a0_);
                    a1 = __ffma2_rn(make_float2(h[1 + 2 * (14 - q)], h[1 + 2 * (14 - q)]), x, a1);
                }
                if (q >= 1) {                             // symbol t + 1: n = 15 - q
                    a2 = __ffma2_rn(make_float2(h[2 * (15 - q)], h[2 * (15 - q)]), x, a2);
                    a3 = __ffma2_rn(make_float2(h[1 + 2 * (15 - q)], h[1 + 2 * (15 - q)]), x, a3);
                }
            }
            // two 16-byte stores per thread when the caller's frame buffer is 16-byte aligned
            if (al16) {
                *reinterpret_cast<float4 *>(o + 2 * t) = make_float4(a0_.x, a0_.y, a1.x, a1.y);
                if (t + 1 < total_syms) *reinterpret_cast<float4 *>(o + 2 * t + 2) = make_float4(a2.x, a2.y, a3.x, a3.y);
            } else {
                o[2 * t] = a0_; o[2 * t + 1] = a1;
                if (t + 1 < total_syms) { o[2 * t + 2] = a2; o[2 * t + 3] = a3; }
            }
        }
    }
}

int tx_fail(int code, const char *msg) { lqb_internal_set_error(msg); return code; }

}  // namespace
}  // namespace lqb

using namespace lqb;

struct lqb_tx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    TxTables *d_tables = nullptr;
    TxFrame *d_frames = nullptr; size_t frames_cap = 0;
    unsigned char *d_pay = nullptr, *d_A = nullptr, *d_B = nullptr; size_t pay_cap = 0, buf_cap = 0, bufB_cap = 0;
    float2 *d_syms = nullptr, *d_out = nullptr; size_t sym_cap = 0, out_cap = 0;
    unsigned *d_ilv = nullptr; size_t ilv_cap = 0, ilv_used = 0;
    // pinned staging: frame plans (two tables in rotation: the host fills one while the upload of the other may still be in
    // flight), host payloads packed back to back (one H2D instead of one copy per frame), host outputs
    TxFrame *h_tab[2] = { nullptr, nullptr }; size_t h_tab_cap[2] = { 0, 0 };
    cudaEvent_t tab_ev[2] = { nullptr, nullptr }; bool tab_busy[2] = { false, false };
    unsigned tab_next = 0;
    // a submitted call that still has to be collected
    bool pending = false, pending_host = false;
    uint32_t pend_n = 0; unsigned pend_tab = 0;
    std::vector<float *> pend_out;
    cudaEvent_t k_ev[2] = { nullptr, nullptr };
    float kernel_ms = 0.0f;
    unsigned char *h_pay = nullptr; size_t h_pay_cap = 0;
    float2 *h_out = nullptr; size_t h_out_cap = 0;
    std::unordered_map<unsigned, size_t> ilv_cache;
    std::vector<unsigned> ilv_host;
    size_t ilv_uploaded = 0;            // entries of ilv_host that are on the device
    uint64_t launches = 0;
};

namespace {
template <typename T> int grow(T *&p, size_t &cap, size_t need)
{
    if (need <= cap) return 0;
    size_t ncap = std::max(need, cap + cap / 2);
    T *q = nullptr;
    if (cudaMalloc(&q, ncap * sizeof(T)) != cudaSuccess) return LQB_ENOMEM;
    if (p) cudaFree(p);
    p = q; cap = ncap;
    return 0;
}
template <typename T> int grow_pinned(T *&p, size_t &cap, size_t need)
{
    if (need <= cap) return 0;
    size_t ncap = std::max(need, cap + cap / 2);
    T *q = nullptr;
    if (cudaMallocHost(&q, ncap * sizeof(T)) != cudaSuccess) return LQB_ENOMEM;
    if (p) cudaFreeHost(p);
    p = q; cap = ncap;
    return 0;
}
}  // namespace

extern "C" {

void lqb_tx_props_init_default(lqb_tx_props *p)
{
    if (!p) return;
    p->check = CRC_16; p->fec0 = FEC_NONE; p->fec1 = FEC_NONE; p->mod_scheme = MODEM_QPSK;
}

int lqb_tx_frame_len(const lqb_tx_props *p, uint32_t payload_len, uint32_t *n_samples)
{
    if (!p || !n_samples) return LQB_EINVAL;
    if (!modem_supported(p->mod_scheme) || !fec_supported(p->fec0) || !fec_supported(p->fec1) || p->check == 0 || p->check >= CRC_NUM || payload_len > 65535)
        return LQB_EINVAL;
    *n_samples = 2 * (64 + 231 + qpm_frame_len(payload_len, p->check, p->fec0, p->fec1, p->mod_scheme) + 14);
    return 0;
}

lqb_tx lqb_tx_create(const lqb_tx_opts *o)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); tx_fail(LQB_ENODEV, "no CUDA device available (no CPU fallback exists)"); return nullptr; }
    int dev = o ? o->device : 0;
    if (dev < 0 || dev >= ndev) { tx_fail(LQB_ENODEV, "device ordinal out of range"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) { tx_fail(LQB_ENODEV, "device is not sm_100 class"); return nullptr; }
    lqb_tx h = new lqb_tx_s;
    h->device = dev;
    cudaSetDevice(dev);
    if (o && o->cuda_stream) h->stream = (cudaStream_t)o->cuda_stream;
    else { cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking); h->own_stream = true; }
    TxTables *T = new TxTables;
    std::memset(T, 0, sizeof *T);
    auto taps = interp_taps(kTxBeta);
    std::memcpy(T->h, taps.data(), 30 * sizeof(float));
    cf pn[64], pil[15];
    preamble_pn(pn); header_pilots(pil);
    for (int i = 0; i < 64; ++i) T->preamble[i] = make_float2(pn[i].re, pn[i].im);
    for (int i = 0; i < 15; ++i) T->pilots[i] = make_float2(pil[i].re, pil[i].im);
    auto maps = psk_maps();
    std::memcpy(T->psk_map, maps.data(), sizeof T->psk_map);
    for (unsigned c = 3; c <= 6; ++c) crc_table(c, T->crc_tab[c]);
    auto m54 = ilv_maps(54), m27 = ilv_maps(27);
    for (unsigned p = 0; p < 4; ++p) {
        for (unsigned i = 0; i < 27; ++i) T->ilv54[p][i] = (uint16_t)m54[p * 27 + i];
        for (unsigned i = 0; i < 13; ++i) T->ilv27[p][i] = (uint16_t)m27[p * 13 + i];
    }
    // the interleaver (passes 0, 1, 2, 3 of masked swaps between byte pairs) moves bits without changing them: run it
    // once on bit labels and keep the permutation, so the device can gather every output byte at once
    auto compose = [](const std::vector<uint32_t> &maps, unsigned n, uint16_t *perm) {
        const unsigned n2 = n / 2, masks[4] = { 0xffu, 0x0fu, 0x55u, 0x33u };
        std::vector<uint16_t> lab(8 * n);
        for (unsigned i = 0; i < 8 * n; ++i) lab[i] = (uint16_t)i;
        for (int pass = 0; pass < 4; ++pass)
            for (unsigned i = 0; i < n2; ++i) {
                const unsigned j = maps[(size_t)pass * n2 + i];
                for (unsigned b = 0; b < 8; ++b)
                    if ((masks[pass] >> b) & 1u) std::swap(lab[8 * (2 * j + 1) + b], lab[8 * (2 * i) + b]);
            }
        for (unsigned i = 0; i < 8 * n; ++i) perm[i] = lab[i];
    };
    compose(m27, 27, T->hperm27f);
    compose(m54, 54, T->hperm54f);
    secded_cols(T->secded_col);
    uint8_t gen[33];
    gf256_tables(T->gf_exp, T->gf_log, gen);
    std::memcpy(T->rs_gen, gen, 33);
    cudaError_t e = cudaMalloc(&h->d_tables, sizeof(TxTables));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, T, sizeof(TxTables), cudaMemcpyHostToDevice);
    delete T;
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&h->tab_ev[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&h->k_ev[k]);
    }
    if (e != cudaSuccess) { tx_fail(LQB_ECUDA, cudaGetErrorString(e)); lqb_tx_destroy(h); return nullptr; }
    return h;
}

void lqb_tx_destroy(lqb_tx h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_tables); cudaFree(h->d_frames); cudaFree(h->d_pay); cudaFree(h->d_A); cudaFree(h->d_B);
    cudaFree(h->d_syms); cudaFree(h->d_out); cudaFree(h->d_ilv);
    for (int k = 0; k < 2; ++k) {
        if (h->h_tab[k]) cudaFreeHost(h->h_tab[k]);
        if (h->tab_ev[k]) cudaEventDestroy(h->tab_ev[k]);
        if (h->k_ev[k]) cudaEventDestroy(h->k_ev[k]);
    }
    if (h->h_pay) cudaFreeHost(h->h_pay);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// wait for everything submitted; host-memory frames are copied out to the caller's buffers
int lqb_tx_collect(lqb_tx h)
{
    if (!h) return LQB_EINVAL;
    if (!h->pending) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return LQB_ECUDA;
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    h->pending = false;
    h->tab_busy[0] = h->tab_busy[1] = false;
    if (e != cudaSuccess) { h->pending_host = false; return tx_fail(LQB_ECUDA, cudaGetErrorString(e)); }
    cudaEventElapsedTime(&h->kernel_ms, h->k_ev[0], h->k_ev[1]);
    if (h->pending_host) {
        const TxFrame *fr = h->h_tab[h->pend_tab];
        for (uint32_t i = 0; i < h->pend_n; ++i) std::memcpy(h->pend_out[i], h->h_out + fr[i].out_off, (size_t)fr[i].n_samples * sizeof(float2));
        h->pending_host = false;
    }
    return 0;
}

int lqb_tx_last_timing(lqb_tx h, float *kernel_ms)
{
    if (!h || !kernel_ms) return LQB_EINVAL;
    *kernel_ms = h->kernel_ms;
    return 0;
}

// Queue the assembly of n frames on the handle's stream and return.  With device buffers nothing waits: frames are
// complete in stream order (the caller's stream when the handle was created on one), and further submits may follow at
// once.  With host buffers the frames reach the caller's memory in lqb_tx_collect().
int lqb_tx_submit(lqb_tx h, uint32_t n, const lqb_tx_props *props, const uint8_t *const *headers,
                  const uint8_t *const *payloads, const uint32_t *lens, float *const *out, int mem)
{
    if (!h || !props || !payloads || !lens || !out) return LQB_EINVAL;
    if (mem != LQB_MEM_HOST && mem != LQB_MEM_DEVICE) return tx_fail(LQB_EINVAL, "frames are complex64 in host or device memory");
    if (!n) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return LQB_ECUDA;
    const bool dev_io = (mem == LQB_MEM_DEVICE);
    if (h->pending_host || (!dev_io && h->pending)) if (int e = lqb_tx_collect(h)) return e;      // one host-memory call at a time
    cudaStream_t st = h->stream;
    const unsigned tab = h->tab_next;
    if (h->tab_busy[tab]) { if (cudaEventSynchronize(h->tab_ev[tab]) != cudaSuccess) return tx_fail(LQB_ECUDA, "event wait failed"); h->tab_busy[tab] = false; }
    if (grow_pinned(h->h_tab[tab], h->h_tab_cap[tab], n)) return tx_fail(LQB_ENOMEM, "cudaMallocHost failed");
    // (the other table too, while it is idle: a pinned allocation in the middle of a run of submits stalls it for milliseconds)
    if (!h->tab_busy[tab ^ 1u] && grow_pinned(h->h_tab[tab ^ 1u], h->h_tab_cap[tab ^ 1u], n)) return tx_fail(LQB_ENOMEM, "cudaMallocHost failed");
    TxFrame *fr = h->h_tab[tab];
    size_t pay_tot = 0, buf_tot = 0, sym_tot = 0, out_tot = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const lqb_tx_props &p = props[i];
        if (lens[i] && !payloads[i]) return tx_fail(LQB_EINVAL, "null payload");
        TxFrame &f = fr[i];
        if (i && lens[i] == lens[i - 1] && std::memcmp(&p, &props[i - 1], sizeof p) == 0) {
            f = fr[i - 1];                 // same scheme and length as the frame before: same plan, new offsets
            std::memset(f.header, 0, sizeof f.header);
        } else {
            uint32_t ns = 0;
            if (int e = lqb_tx_frame_len(&p, lens[i], &ns)) return tx_fail(e, "unsupported props or payload length");
            std::memset(&f, 0, sizeof f);
            f.payload_len = lens[i]; f.check = p.check; f.fec0 = p.fec0; f.fec1 = p.fec1; f.ms = p.mod_scheme; f.bps = modem_bps(p.mod_scheme);
            f.k0 = lens[i] + crc_len(p.check); f.n0 = fec_enc_len(p.fec0, f.k0); f.n1 = fec_enc_len(p.fec1, f.n0);
            f.n_sym = (8 * f.n1 + f.bps - 1) / f.bps; f.n_samples = ns;
            f.buf_len = ((std::max(std::max(f.n1, f.n0), f.k0) + 16) + 15u) & ~15u;
            const unsigned encs[2] = { f.n0, f.n1 }, fss[2] = { f.fec0, f.fec1 };
            for (int s = 0; s < 2; ++s) {
                if (fss[s] == FEC_NONE) continue;
                auto it = h->ilv_cache.find(encs[s]);
                size_t off;
                if (it == h->ilv_cache.end()) {
                    auto maps = ilv_maps(encs[s]);
                    off = h->ilv_host.size();
                    h->ilv_host.insert(h->ilv_host.end(), maps.begin(), maps.end());
                    h->ilv_cache.emplace(encs[s], off);
                } else off = it->second;
                (s ? f.ilv1_off : f.ilv0_off) = (unsigned)off;
            }
        }
        f.pay_off = pay_tot; pay_tot += (lens[i] + 15u) & ~15u;
        f.buf_off = buf_tot; buf_tot += f.buf_len;
        f.sym_off = sym_tot; sym_tot += (f.n_sym + 14 + 64 + 231 + 16 + 1) & ~(size_t)1;      // lead zeros, preamble, header, payload, tail zeros
        f.out_off = out_tot; out_tot += f.n_samples;
        if (headers && headers[i]) std::memcpy(f.header, headers[i], 14);
    }
    if (grow(h->d_frames, h->frames_cap, n) || grow(h->d_A, h->buf_cap, buf_tot + 16)) return tx_fail(LQB_ENOMEM, "cudaMalloc failed");
    if (grow(h->d_B, h->bufB_cap, buf_tot + 16)) return tx_fail(LQB_ENOMEM, "cudaMalloc failed");
    if (grow(h->d_syms, h->sym_cap, sym_tot + 1)) return tx_fail(LQB_ENOMEM, "cudaMalloc failed");
    // the device copy of the interleaver maps follows the host cache by COUNT, not by "this call added maps": a call
    // that added maps and then failed (bad props further down the list, out of memory) leaves them to the next call
    if (h->ilv_host.size() > h->ilv_uploaded || !h->d_ilv) {
        if (h->pending) cudaStreamSynchronize(st);          // growing frees the old arena: no queued kernel may still read it
        if (grow(h->d_ilv, h->ilv_cap, h->ilv_host.size() + 4)) return tx_fail(LQB_ENOMEM, "cudaMalloc failed");
        if (!h->ilv_host.empty() &&
            cudaMemcpyAsync(h->d_ilv, h->ilv_host.data(), h->ilv_host.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return tx_fail(LQB_ECUDA, "interleaver map upload failed");
        h->ilv_uploaded = h->ilv_host.size();
    }
    // Device buffers are used in place: the kernel reads every payload and writes every frame through the caller's
    // pointers.  Host buffers are packed into one pinned staging area each way: one H2D and one D2H per call instead of
    // one copy per frame (8192 frames took 32 ms of copy calls, profiles/r01_notes.md v23).
    if (dev_io) {
        for (uint32_t i = 0; i < n; ++i) { fr[i].pay = payloads[i]; fr[i].out = reinterpret_cast<float2 *>(out[i]); }
    } else {
        if (grow(h->d_pay, h->pay_cap, pay_tot + 16) || grow(h->d_out, h->out_cap, out_tot + 1)) return tx_fail(LQB_ENOMEM, "cudaMalloc failed");
        if (grow_pinned(h->h_pay, h->h_pay_cap, pay_tot + 16) || grow_pinned(h->h_out, h->h_out_cap, out_tot + 1)) return tx_fail(LQB_ENOMEM, "cudaMallocHost failed");
        for (uint32_t i = 0; i < n; ++i) {
            if (lens[i]) std::memcpy(h->h_pay + fr[i].pay_off, payloads[i], lens[i]);
            fr[i].pay = h->d_pay + fr[i].pay_off; fr[i].out = h->d_out + fr[i].out_off;
        }
        if (pay_tot) cudaMemcpyAsync(h->d_pay, h->h_pay, pay_tot, cudaMemcpyHostToDevice, st);
    }
    cudaMemcpyAsync(h->d_frames, fr, n * sizeof(TxFrame), cudaMemcpyHostToDevice, st);
    cudaEventRecord(h->tab_ev[tab], st);
    h->tab_busy[tab] = true;
    h->tab_next = tab ^ 1u;
    cudaEventRecord(h->k_ev[0], st);
    k_tx<<<n, kTxThreads, 0, st>>>(h->d_tables, h->d_frames, h->d_A, h->d_B, h->d_ilv, h->d_syms);
    cudaEventRecord(h->k_ev[1], st);
    h->launches++;
    h->pending = true;
    if (!dev_io) {
        cudaMemcpyAsync(h->h_out, h->d_out, out_tot * sizeof(float2), cudaMemcpyDeviceToHost, st);
        h->pending_host = true; h->pend_n = n; h->pend_tab = tab;
        h->pend_out.assign(out, out + n);
    }
    if (cudaGetLastError() != cudaSuccess) return tx_fail(LQB_ECUDA, "k_tx launch failed");
    return 0;
}

int lqb_tx_assemble(lqb_tx h, uint32_t n, const lqb_tx_props *props, const uint8_t *const *headers,
                    const uint8_t *const *payloads, const uint32_t *lens, float *const *out, int mem)
{
    if (int e = lqb_tx_submit(h, n, props, headers, payloads, lens, out, mem)) return e;
    return lqb_tx_collect(h);
}

}  // extern "C"
