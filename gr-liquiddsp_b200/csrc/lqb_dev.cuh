// lqb_dev.cuh -- device-side data layout and arithmetic primitives shared by the kernels.
//
// Float discipline: the library is compiled with -fmad=false; fused multiply-adds are
// written __fmaf_rn() exactly where the algorithm specification (docs/FRAME_FORMAT.md)
// has one, so every kernel rounds the way the specification says.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lqb_lens.h"

namespace lqb {

// ------------------------------------------------------------------ constant tables in HBM (one copy per handle)
struct DevTables {
    float2   W512[256];        // exp(-j 2 pi k / 512)
    float2   W32[16];          // exp(-j 2 pi k / 32)
    float2   Sc[512];          // conj(FFT512(template))
    float2   sconj[160];       // conj(template), 156 used
    float    s2_sum, threshold;
    int      range, pad0;
    float    sintab[1024];     // NCO sine table
    float    banks[32 * 28];   // matched-filter bank taps, oldest -> newest
    float2   pilots_conj[16];  // conj(header pilots), 15 used
    float2   psk_map[8 * 256]; // PSK-2^b constellation, row b-1
    uint32_t crc_tab[8][256];  // by liquid crc enum (3..6 used)
    uint16_t ilv54[4][28];     // header interleaver maps, n = 54 (27 pairs)
    uint16_t ilv27[4][16];     // n = 27 (13 pairs)
    uint16_t hperm54[432];     // the four deinterleaver passes over 54 bytes as one bit permutation: out bit -> in bit
    uint16_t hperm27[216];     // (bit b of byte i has index 8 i + b, b = 0 the least significant bit)
    uint8_t  h84_dec[256];
    uint8_t  h74_dec[128];
    uint8_t  secded_col[3][64]; // liquid's Hsiao codes, by code (0: (22,16), 1: (39,32), 2: (72,64)) and data bit: parity-byte contribution
    uint8_t  gf_exp[512];
    uint8_t  gf_log[256];
    uint8_t  rs_gen[64];       // 33 used
    alignas(16) uint32_t rs_syn[256][32];  // syndrome products: entry [v][i] packs v*b, v*b^2, v*b^3, v*b^4 (byte 0..3), b = alpha^(i+1)
};

// ------------------------------------------------------------------ per-stream persistent state
struct StreamState {
    long long base;            // absolute index of carry[0]
    long long G;               // samples with absolute index < G read as zero (reset boundary)
    long long wstart;          // SEEK: absolute start of the next 512-sample window
    long long F;               // PENDING: absolute index of the detected frame start
    long long need_until;      // PENDING: do nothing until samples < need_until are available
    long long resume;          // out: absolute index from which samples must be carried over
    unsigned  carry_len;
    int       mode;            // 0 = SEEK, 1 = PENDING
    int       offset;          // PENDING: detected CFO bin
    float     rxy;             // PENDING: detection metric
    unsigned  seq;             // frames emitted since reset
    unsigned  dropped;
    unsigned  carry_sel;       // which of the two carry buffers is current
    unsigned  det_idx;         // PENDING: offset of the frame start in the window that triggered (F = window start + det_idx)
    // time-sharded search (one capture cut into segments that run as streams of their own, lqb_det_execute_sharded):
    long long stop_at;         // windows starting at or beyond this index are left alone (kNoStop: none)
    long long mark_at;         // mark_w <- the first window start >= mark_at this stream's walk reaches (kNoMark: unset)
    long long mark_w, mark_G;  // (the zero boundary G at that point belongs to the state as long as it lies above the window start)
};
constexpr long long kNoStop = 0x3fffffffffffffffll, kNoMark = -0x3fffffffffffffffll;

struct StreamIO {              // one entry per stream fed by this execute call
    const float2      *in;
    unsigned long long n_in;
    unsigned           stream;
    unsigned           pad;
};

// ------------------------------------------------------------------ frame descriptor (device <-> host)
struct FrameDesc {
    long long F, G;
    unsigned long long sym_off;      // payload symbols: offset into the symbol arena (complex samples)
    unsigned long long buf_off;      // offset into byte arena A / B (per-frame stride buf_len)
    unsigned long long pay_off;      // offset into payload output pool
    unsigned long long dec_off;      // offset into Viterbi decision arena (64-bit words)
    unsigned stream, seq, io_index, flags;
    float    tau, gamma, dphi, phi, rxy, mf_scale;
    unsigned mix_theta0, mix_dtheta, pll_theta0, pll_dtheta;
    unsigned pfb_index, tau_neg;
    int      header_valid, payload_valid;
    unsigned payload_len, ms, bps, check, fec0, fec1;
    unsigned n_sym, k0, n0, n1;      // k0 = payload+crc bytes, n0 = after fec0, n1 = after fec1
    unsigned buf_len;                // bytes reserved per byte buffer
    unsigned ilv1_off, ilv0_off;     // offsets (in uint32 entries) of this frame's interleaver maps
    float    evm, rssi, cfo, evm_acc;
    unsigned char header[20];
    unsigned ck_off;                 // first PLL checkpoint of this frame (16-byte entries)
    unsigned det_idx;                // offset of F in the window that triggered (F - det_idx: that window's start)
    unsigned pad;
};

struct Detection {
    long long F;
    unsigned  stream, seq;
    float     tau, gamma, dphi, phi, rxy;
    unsigned  pad;
};

// ------------------------------------------------------------------ arithmetic primitives
// (A packed-pair variant built on sm_100's FFMA2/FMUL2/FADD2 is bit-identical but measured 10 % slower
// in the detector kernel -- see profiles/r01_notes.md -- so the scalar forms stay.)
__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    float2 y;
    y.x = __fmaf_rn(-a.y, b.y, __fmul_rn(a.x, b.x));
    y.y = __fmaf_rn(a.y, b.x, __fmul_rn(a.x, b.y));
    return y;
}
__device__ __forceinline__ float abs2f(float2 a) { return __fmaf_rn(a.y, a.y, __fmul_rn(a.x, a.x)); }
__device__ __forceinline__ float cabsf_(float2 a) { return __fsqrt_rn(abs2f(a)); }

// constellation point of symbol s (the modulator's map; the same operations as the host-built tables of the
// specification, so the points are bit-identical to the oracle's).  psk_map: [8][256] PSK-2^b rows.
__device__ __forceinline__ unsigned gray_decode_dev(unsigned s) { unsigned r = s; for (unsigned sh = 1; sh < 32; sh <<= 1) r ^= r >> sh; return r; }
__device__ inline float2 modem_point(const float2 *psk_map, unsigned ms, unsigned bps, unsigned s)
{
    const unsigned M = 1u << bps;
    if (ms >= 1 && ms <= 8) return psk_map[(bps - 1) * 256 + s];
    if (ms >= 17 && ms <= 24) {
        const float c[9] = { 0, 1.0f, 5.0f, 21.0f, 85.0f, 341.0f, 1365.0f, 5461.0f, 21845.0f };
        const float alpha = __fdiv_rn(1.0f, __fsqrt_rn(c[bps]));
        return make_float2(__fmul_rn((float)(2 * (int)gray_decode_dev(s) - (int)M + 1), alpha), 0.0f);
    }
    if (ms >= 25 && ms <= 31) {
        const float c[9] = { 0, 0, 2.0f, 6.0f, 10.0f, 26.0f, 42.0f, 106.0f, 170.0f };
        const float alpha = __fdiv_rn(1.0f, __fsqrt_rn(c[bps]));
        const unsigned m_i = (bps + 1) >> 1, m_q = bps >> 1;
        const unsigned si = gray_decode_dev(s >> m_q), sq = gray_decode_dev(s & ((1u << m_q) - 1u));
        return make_float2(__fmul_rn((float)(2 * (int)si - (int)(1u << m_i) + 1), alpha),
                           __fmul_rn((float)(2 * (int)sq - (int)(1u << m_q) + 1), alpha));
    }
    if (ms == 39) return make_float2(s ? -1.0f : 1.0f, 0.0f);
    return make_float2((s & 1u) ? -0.707106769f : 0.707106769f, (s & 2u) ? -0.707106769f : 0.707106769f);
}

__device__ __forceinline__ uint32_t nco_constrain_dev(float theta)
{
    // identical to: p = theta/(2 pi); f = p - trunc(p); f < 0 ? f + 1; (uint32)(int64)(f * 2^32)
    // (truncf == (float)(int64)p for |p| < 2^63; f == 1.0 -- a tiny negative plus one -- wraps to 0)
    float p = __fmul_rn(theta, 0.15915494309189535f);
    float f = __fsub_rn(p, truncf(p));
    if (f < 0.0f) f = __fadd_rn(f, 1.0f);
    return f >= 1.0f ? 0u : __float2uint_rz(__fmul_rn(f, 4294967296.0f));
}
__device__ __forceinline__ float2 nco_mix_down(const float *__restrict__ sintab, uint32_t theta, float2 x)
{
    unsigned idx = ((theta + (1u << 21)) >> 22) & 0x3ffu;
    float s = sintab[idx], c = sintab[(idx + 256u) & 0x3ffu];
    float2 y;                                   // x * (c - j s)
    y.x = __fmaf_rn(x.y, s, __fmul_rn(x.x, c));
    y.y = __fmaf_rn(-x.x, s, __fmul_rn(x.y, c));
    return y;
}
__device__ __forceinline__ float nco_get_frequency_dev(uint32_t d_theta)
{
    float d = __fmul_rn((float)d_theta, 1.4629180792671596e-09f);
    return d > 3.14159274f ? __fsub_rn(d, 6.28318548f) : d;
}

// arg() and exp(j t) of the per-symbol loops (generic PSK / DPSK modems), pinned: the specification (oracle/lqo_modem.c,
// lqo_pm_atan2f / lqo_pm_sincosf) fixes Cephes' single-precision algorithms as a sequence of plain IEEE operations, and
// this is that sequence (libdevice's atan2f / sincosf differ from it -- and from every libm -- in the last bit, which
// reached the PLL phase and moved DPSK / PSK8 constellation points by one NCO table step).
__device__ __forceinline__ float pm_atanf_pos(float x)
{
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -__fdiv_rn(1.0f, x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    else y = 0.0f;
    const float z = __fmul_rn(x, x);
    float p = __fsub_rn(__fmul_rn(8.05374449538e-2f, z), 1.38776856032e-1f);
    p = __fadd_rn(__fmul_rn(p, z), 1.99777106478e-1f);
    p = __fsub_rn(__fmul_rn(p, z), 3.33329491539e-1f);
    p = __fadd_rn(__fmul_rn(__fmul_rn(p, z), x), x);
    return __fadd_rn(y, p);
}
__device__ __forceinline__ float pm_atan2f(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    float r = (ax == 0.0f && ay == 0.0f) ? 0.0f : pm_atanf_pos(__fdiv_rn(ay, ax));
    if (x < 0.0f) r = __fsub_rn(3.14159274f, r);
    return y < 0.0f ? -r : r;
}
__device__ __forceinline__ void pm_sincosf(float t, float *sn, float *cs)
{
    float x = fabsf(t);
    int j = __float2int_rz(__fmul_rn(1.27323954473516f, x));
    if (j & 1) j += 1;
    const float y = (float)j;
    j &= 7;
    bool s_neg = t < 0.0f, c_neg = false;
    if (j > 3) { s_neg = !s_neg; c_neg = !c_neg; j -= 4; }
    if (j > 1) c_neg = !c_neg;
    x = __fsub_rn(__fsub_rn(__fsub_rn(x, __fmul_rn(y, 0.78515625f)), __fmul_rn(y, 2.4187564849853515625e-4f)), __fmul_rn(y, 3.77489497744594108e-8f));
    const float z = __fmul_rn(x, x);
    float ps = __fadd_rn(__fmul_rn(-1.9515295891e-4f, z), 8.3321608736e-3f);
    ps = __fsub_rn(__fmul_rn(ps, z), 1.6666654611e-1f);
    ps = __fadd_rn(__fmul_rn(__fmul_rn(ps, z), x), x);
    float pc = __fsub_rn(__fmul_rn(2.443315711809948e-5f, z), 1.388731625493765e-3f);
    pc = __fadd_rn(__fmul_rn(pc, z), 4.166664568298827e-2f);
    pc = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(pc, z), z), __fmul_rn(0.5f, z)), 1.0f);
    const bool swap = (j == 1 || j == 2);
    const float s = swap ? pc : ps, c = swap ? ps : pc;
    *sn = s_neg ? -s : s;
    *cs = c_neg ? -c : c;
}

// sample accessor over [carry | new input] with the zero region below G
struct StreamView {
    const float2 *carry;
    const float2 *in;
    long long     base, G, end;     // end = base + carry_len + n_in (exclusive)
    unsigned      carry_len;
    __device__ __forceinline__ float2 at(long long n) const
    {
        if (n < G || n >= end || n < base) return make_float2(0.0f, 0.0f);
        long long i = n - base;
        return (i < (long long)carry_len) ? carry[i] : in[i - carry_len];
    }
};

__device__ __forceinline__ unsigned brev4(unsigned r) { return __brev(r) >> 28; }
__device__ __forceinline__ unsigned brev5(unsigned r) { return __brev(r) >> 27; }

// ------------------------------------------------------------------ 512-point FFT, one warp, 16 points per lane
// Radix-2 decimation-in-time, identical butterfly order and twiddles to the specification's
// iterative FFT: stages 1-4 in registers, transpose through shared memory, stages 5-8 in
// registers, stage 9 against lane^16 with shuffles.
// In:  v[r] = x[bitrev9((lane << 4) | r)]      Out: v[r] = X[((lane >> 4) << 8) | (r << 4) | (lane & 15)]
template <int DIR>
__device__ __forceinline__ void bfly(float2 &lo, float2 &hi, float2 w)
{
    const float wi = DIR > 0 ? w.y : -w.y;
    const float tr = __fmaf_rn(-wi, hi.y, __fmul_rn(w.x, hi.x));
    const float ti = __fmaf_rn(wi, hi.x, __fmul_rn(w.x, hi.y));
    const float2 u = lo;
    lo.x = __fadd_rn(u.x, tr); lo.y = __fadd_rn(u.y, ti);
    hi.x = __fsub_rn(u.x, tr); hi.y = __fsub_rn(u.y, ti);
}
// twiddle 1 and the exact quarter turn (0,-1) / (0,+1): the products are exact, only the adds remain
__device__ __forceinline__ void bfly_one(float2 &lo, float2 &hi)
{
    const float2 u = lo, t = hi;
    lo.x = __fadd_rn(u.x, t.x); lo.y = __fadd_rn(u.y, t.y);
    hi.x = __fsub_rn(u.x, t.x); hi.y = __fsub_rn(u.y, t.y);
}
template <int DIR>
__device__ __forceinline__ void bfly_quarter(float2 &lo, float2 &hi)
{
    const float2 t = DIR > 0 ? make_float2(hi.y, -hi.x) : make_float2(-hi.y, hi.x);
    const float2 u = lo;
    lo.x = __fadd_rn(u.x, t.x); lo.y = __fadd_rn(u.y, t.y);
    hi.x = __fsub_rn(u.x, t.x); hi.y = __fsub_rn(u.y, t.y);
}

// W  : 256 twiddles exp(-j 2 pi k / 512) (stages 1-4 read compile-time entries, stage 9 reads 16 r + c)
// Wc : per-stage compact copies for stages 5-8 so that the 16 lanes of a half-warp read consecutive
//      words (bank-conflict free): Wc[16 (2^s - 1) + j] = W[j * (16 >> s)], s = 0..3, j < 16 * 2^s
template <int DIR>
__device__ __forceinline__ void fft512_warp(float2 (&v)[16], const float2 *__restrict__ W, const float2 *__restrict__ Wc,
                                            float2 *__restrict__ scratch /* 544 float2, this warp's */, int lane)
{
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int half = 1 << s;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r & half) continue;
            const int e = (r & (half - 1)) * (256 >> s);
            if (e == 0) bfly_one(v[r], v[r + half]);
            else if (e == 128) bfly_quarter<DIR>(v[r], v[r + half]);
            else bfly<DIR>(v[r], v[r + half], W[e]);
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) scratch[17 * lane + r] = v[r];
    __syncwarp();
    const int c = lane & 15, b8 = lane >> 4;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int p = (b8 << 8) | (r << 4) | c;
        v[r] = scratch[p + (p >> 4)];
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int half = 1 << s;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r & half) continue;
            bfly<DIR>(v[r], v[r + half], Wc[16 * (half - 1) + c + 16 * (r & (half - 1))]);
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, v[r].x, 16);
        o.y = __shfl_xor_sync(0xffffffffu, v[r].y, 16);
        float2 lo = b8 ? o : v[r], hi = b8 ? v[r] : o;
        bfly<DIR>(lo, hi, W[(r << 4) | c]);
        v[r] = b8 ? hi : lo;
    }
}

// fill the compact stage tables from the full table (any number of threads; caller synchronises)
__device__ __forceinline__ void fft512_fill_compact(float2 *Wc, const float2 *W, int tid, int nthreads)
{
    for (int i = tid; i < 240; i += nthreads) {
        int s = (i >= 112) ? 3 : (i >= 48) ? 2 : (i >= 16) ? 1 : 0;
        int j = i - 16 * ((1 << s) - 1);
        Wc[i] = W[j * (16 >> s)];
    }
}

__device__ __forceinline__ int fft512_out_index(int lane, int r) { return ((lane >> 4) << 8) | (r << 4) | (lane & 15); }
__device__ __forceinline__ int fft512_in_index(int lane, int r) { return (int)((brev4((unsigned)r) << 5) | brev5((unsigned)lane)); }

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, k, m);
        k = o > k ? o : k;
    }
    return k;
}

}  // namespace lqb
