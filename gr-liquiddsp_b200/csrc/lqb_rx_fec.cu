// lqb_rx_fec.cu -- packetizer_decode on the GPU, frame-parallel:
//   k_deinterleave : the four interleaver passes (CTA per frame, precomputed swap maps)
//   k_blockfec     : NONE / REP3 / REP5 / Hamming(7,4)(8,4)(12,8) / Golay(24,12) / SECDED
//   k_viterbi27x4  : K=7 rate-1/2 (optionally punctured) hard-input Viterbi, four lanes per codeword,
//                    constant-geometry register layout, sign-bit decisions, prefetched traceback
//   k_viterbi      : K=9 variant, warp per codeword: add-compare-select with metrics in shared
//                    memory, decisions by ballot to HBM, then a shuffle-fed traceback from state 0
//   k_rs           : RS(255,223) over GF(256)/0x11d, warp per block: 32 syndromes in parallel
//                    (one per lane), Berlekamp-Massey on lane 0, parallel Chien + Forney
//   k_crc          : unscramble + CRC/checksum + payload copy-out, warp per frame
// This is what liquid-dsp's qpacketmodem_decode/packetizer_decode (and libfec underneath)
// do at the end of flexframesync_execute_rxpayload (reference call site
// lib/flex_rx_impl.cc:213; schemes named at lib/flex_rx_impl.cc:75-136).
#include <cstdlib>
#include <algorithm>
#include "lqb_dev.cuh"
#include "lqb_kernels.h"

namespace lqb {

namespace {

struct StageIO { const unsigned char *src; unsigned char *dst; unsigned enc_len, dec_len, fs; };

__device__ __forceinline__ StageIO stage_io(const PayloadParams &P, const FrameDesc &d, int stage)
{
    StageIO s;
    if (stage == 1) { s.src = P.bufA + d.buf_off; s.dst = P.bufB + d.buf_off; s.enc_len = d.n1; s.dec_len = d.n0; s.fs = d.fec1; }
    else            { s.src = P.bufB + d.buf_off; s.dst = P.bufA + d.buf_off; s.enc_len = d.n0; s.dec_len = d.k0; s.fs = d.fec0; }
    return s;
}

// ------------------------------------------------------------------ interleaver
// The frame's bytes are brought into shared memory once (16 KB window; longer buffers take the
// global-memory route), the four involutive passes run there, and the result is written back.
constexpr unsigned kIlvSmem = 16384;

__global__ void __launch_bounds__(256)
k_deinterleave(PayloadParams P, const unsigned *__restrict__ list, int stage)
{
    __shared__ __align__(16) unsigned char buf[kIlvSmem];
    const FrameDesc &d = P.frames[list[blockIdx.x]];
    unsigned char *g = (stage == 1) ? (P.bufA + d.buf_off) : (P.bufB + d.buf_off);
    const unsigned n = (stage == 1) ? d.n1 : d.n0, n2 = n / 2;
    const unsigned *maps = P.ilv_maps + ((stage == 1) ? d.ilv1_off : d.ilv0_off);
    const unsigned masks[4] = { 0xffu, 0x0fu, 0x55u, 0x33u };
    const bool in_smem = n <= kIlvSmem;
    unsigned char *x = in_smem ? buf : g;
    if (in_smem) {
        const unsigned n16 = (n + 15) / 16;                     // buffers are 16-byte aligned and padded
        for (unsigned i = threadIdx.x; i < n16; i += blockDim.x)
            reinterpret_cast<uint4 *>(buf)[i] = reinterpret_cast<const uint4 *>(g)[i];
        __syncthreads();
    }
    for (int pass = 3; pass >= 0; --pass) {
        const unsigned *map = maps + (size_t)pass * n2;
        const unsigned mask = masks[pass];
        for (unsigned i = threadIdx.x; i < n2; i += blockDim.x) {
            unsigned j = map[i];
            unsigned a = x[2 * j + 1], b = x[2 * i];
            x[2 * j + 1] = (unsigned char)((a & ~mask) | (b & mask));
            x[2 * i] = (unsigned char)((a & mask) | (b & ~mask));
        }
        __syncthreads();
    }
    if (in_smem) {
        const unsigned n16 = (n + 15) / 16;
        for (unsigned i = threadIdx.x; i < n16; i += blockDim.x)
            reinterpret_cast<uint4 *>(g)[i] = reinterpret_cast<const uint4 *>(buf)[i];
    }
}

// ------------------------------------------------------------------ block codes
__device__ __forceinline__ unsigned get_bits(const unsigned char *src, unsigned k, unsigned b)
{
    unsigned s = 0;
    for (unsigned i = 0; i < b; ++i) { unsigned pos = k + i; s = (s << 1) | ((src[pos >> 3] >> (7 - (pos & 7))) & 1u); }
    return s;
}

__constant__ unsigned c_golay_P[12] = { 0x8ed, 0x1db, 0x3b5, 0x769, 0xed1, 0xda3, 0xb47, 0x68f, 0xd1d, 0xa3b, 0x477, 0xffe };

__device__ __forceinline__ unsigned golay_mulP(unsigned v)
{
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r |= ((unsigned)__popc(v & c_golay_P[i]) & 1u) << (11 - i);
    return r;
}
__device__ unsigned golay_decode(unsigned r)
{
    const unsigned rp = (r >> 12) & 0xfffu, rm = r & 0xfffu;
    const unsigned s = rp ^ golay_mulP(rm);
    if (__popc(s) <= 3) return rm;
    for (int i = 0; i < 12; ++i)
        if (__popc(s ^ c_golay_P[i]) <= 2) return rm ^ (1u << (11 - i));
    const unsigned q = golay_mulP(s);
    if (__popc(q) <= 3) return rm ^ q;
    for (int i = 0; i < 12; ++i)
        if (__popc(q ^ c_golay_P[i]) <= 2) return rm ^ q ^ c_golay_P[i];
    return rm;
}

__device__ unsigned h128_decode(unsigned r)
{
    const int pos2bit[13] = { -1, 11, 10, 7, 9, 6, 5, 4, 8, 3, 2, 1, 0 };
    const unsigned d = r & 0xffu, p = (r >> 8) & 0xfu;
    const unsigned z = ((((unsigned)__popc(d & 0xda) ^ (p >> 3)) & 1u) << 0) | ((((unsigned)__popc(d & 0xb6) ^ (p >> 2)) & 1u) << 1)
                     | ((((unsigned)__popc(d & 0x71) ^ (p >> 1)) & 1u) << 2) | ((((unsigned)__popc(d & 0x0f) ^ p) & 1u) << 3);
    if (z >= 1 && z <= 12) r ^= 1u << pos2bit[z];
    return r & 0xffu;
}

__device__ void secded_block(const DevTables *T, const unsigned char *src, unsigned char *dst, unsigned nb, unsigned R, unsigned r)
{
    // src: [parity][r data bytes], block of nb data bytes (missing ones are zero), R parity bits; writes r bytes.
    // Syndrome = parity of the data XOR the received parity bits: zero: nothing; a column of P: that data bit is
    // flipped back; anything else (a parity bit, or two or more errors): left as received.
    const unsigned char *col = T->secded_col[nb == 2 ? 0 : nb == 4 ? 1 : 2];
    unsigned char blk[8];
    unsigned p = 0;
    for (unsigned q = 0; q < nb; ++q) blk[q] = q < r ? src[1 + q] : 0;
    for (unsigned bit = 0; bit < nb * 8; ++bit)
        if ((blk[bit >> 3] >> (7 - (bit & 7))) & 1u) p ^= col[bit];
    const unsigned syn = p ^ (src[0] & ((1u << R) - 1u));
    if (syn)
        for (unsigned bit = 0; bit < nb * 8; ++bit)
            if (col[bit] == syn) { blk[bit >> 3] ^= (unsigned char)(0x80u >> (bit & 7)); break; }
    for (unsigned q = 0; q < r; ++q) dst[q] = blk[q];
}

__global__ void __launch_bounds__(256)
k_blockfec(PayloadParams P, const unsigned *__restrict__ list, int stage)
{
    const FrameDesc &d = P.frames[list[blockIdx.x]];
    const StageIO io = stage_io(P, d, stage);
    const DevTables *T = P.tables;
    const unsigned n = io.dec_len;
    const unsigned char *e = io.src;
    unsigned char *o = io.dst;
    const unsigned tid = threadIdx.x, nt = blockDim.x;
    switch (io.fs) {
    case 1: for (unsigned i = tid; i < n; i += nt) o[i] = e[i]; break;
    case 2:
        for (unsigned i = tid; i < n; i += nt) { unsigned a = e[i], b = e[i + n], c = e[i + 2 * n]; o[i] = (unsigned char)((a & b) | (a & c) | (b & c)); }
        break;
    case 3:
        for (unsigned i = tid; i < n; i += nt) {
            unsigned out = 0;
            for (unsigned bit = 0; bit < 8; ++bit) {
                unsigned cnt = 0;
                for (unsigned r = 0; r < 5; ++r) cnt += (e[i + r * n] >> bit) & 1u;
                if (cnt >= 3) out |= 1u << bit;
            }
            o[i] = (unsigned char)out;
        }
        break;
    case 4:
        for (unsigned i = tid; i < n; i += nt)
            o[i] = (unsigned char)((T->h74_dec[get_bits(e, 14 * i, 7)] << 4) | T->h74_dec[get_bits(e, 14 * i + 7, 7)]);
        break;
    case 5:
        for (unsigned i = tid; i < n; i += nt) o[i] = (unsigned char)((T->h84_dec[e[2 * i]] << 4) | T->h84_dec[e[2 * i + 1]]);
        break;
    case 6:
        for (unsigned i = tid; i < n; i += nt) o[i] = (unsigned char)h128_decode(get_bits(e, 12 * i, 12));
        break;
    case 7: {
        const unsigned groups = n / 3, rem = n % 3;
        for (unsigned g = tid; g < groups; g += nt) {
            const unsigned char *s = e + 6 * g;
            unsigned v0 = ((unsigned)s[0] << 16) | ((unsigned)s[1] << 8) | s[2], v1 = ((unsigned)s[3] << 16) | ((unsigned)s[4] << 8) | s[5];
            unsigned s0 = golay_decode(v0), s1 = golay_decode(v1);
            o[3 * g] = (unsigned char)(s0 >> 4); o[3 * g + 1] = (unsigned char)(((s0 & 15u) << 4) | (s1 >> 8)); o[3 * g + 2] = (unsigned char)s1;
        }
        if (tid < rem) {
            const unsigned char *s = e + 6 * groups + 3 * tid;
            o[3 * groups + tid] = (unsigned char)golay_decode(((unsigned)s[0] << 16) | ((unsigned)s[1] << 8) | s[2]);
        }
        break;
    }
    case 8: case 9: case 10: {
        const unsigned nb = io.fs == 8 ? 2u : io.fs == 9 ? 4u : 8u, nc = io.fs == 8 ? 6u : io.fs == 9 ? 7u : 8u;      // data bytes, parity bits
        const unsigned blocks = (n + nb - 1) / nb;
        for (unsigned b = tid; b < blocks; b += nt) {
            unsigned r = (n - b * nb >= nb) ? nb : (n - b * nb);
            secded_block(T, e + b * (nb + 1), o + b * nb, nb, nc, r);
        }
        break;
    }
    default: break;
    }
}

// ------------------------------------------------------------------ Viterbi (warp per codeword)
struct ConvSpec { unsigned K, P, poly0, poly1, keep0, keep1; };   // keepR: bit c set if row R keeps column c

__device__ __forceinline__ ConvSpec conv_spec(unsigned fs)
{
    ConvSpec c;
    c.K = 7; c.P = 1; c.poly0 = 0x6d; c.poly1 = 0x4f; c.keep0 = 1; c.keep1 = 1;
    // puncturing matrices as column bitmasks, column 0 = bit 0
    const unsigned k27[6][2] = { { 0x3, 0x1 }, { 0x3, 0x5 }, { 0xf, 0x1 }, { 0xb, 0x15 }, { 0x17, 0x29 }, { 0x2f, 0x51 } };
    const unsigned k29[6][2] = { { 0x3, 0x1 }, { 0x7, 0x1 }, { 0xd, 0x3 }, { 0xb, 0x15 }, { 0x1b, 0x25 }, { 0x6b, 0x15 } };
    if (fs == 12 || (fs >= 21 && fs <= 26)) { c.K = 9; c.poly0 = 0x1af; c.poly1 = 0x11d; }
    if (fs >= 15 && fs <= 20) { c.P = fs - 13; c.keep0 = k27[fs - 15][0]; c.keep1 = k27[fs - 15][1]; }
    if (fs >= 21 && fs <= 26) { c.P = fs - 19; c.keep0 = k29[fs - 21][0]; c.keep1 = k29[fs - 21][1]; }
    return c;
}

__device__ __forceinline__ unsigned soft_bit(const unsigned char *enc, unsigned ib) { return ((enc[ib >> 3] >> (7 - (ib & 7))) & 1u) ? 255u : 0u; }

constexpr int kVitWarps = 4;

// SOFT: the received values are bytes (one per kept coded bit, already deinterleaved: LQB_RX_SOFT) instead of bits
template <bool SOFT>
__global__ void __launch_bounds__(32 * kVitWarps)
k_viterbi(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list, int stage)
{
    __shared__ unsigned metrics[kVitWarps][2][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned gi = blockIdx.x * kVitWarps + warp;
    if (gi >= n_list) return;
    const FrameDesc &d = P.frames[list[gi]];
    StageIO io = stage_io(P, d, stage);
    if (SOFT) { io.src = P.soft_d + P.soft[list[gi]].d_off; io.enc_len *= 8u; }     // one byte per coded bit
    const ConvSpec cs = conv_spec(io.fs);
    const unsigned ns = 1u << (cs.K - 1), half = ns >> 1, words = ns >> 5;
    const unsigned nbits = 8 * io.dec_len, T = nbits + cs.K - 1;
    unsigned *dec = reinterpret_cast<unsigned *>(P.decisions + d.dec_off);
    unsigned (*m)[256] = metrics[warp];

    // kept bits per puncturing period and prefix counts per column
    unsigned per = 0, pre[8];
    for (unsigned c = 0; c < cs.P; ++c) { pre[c] = per; per += ((cs.keep0 >> c) & 1u) + ((cs.keep1 >> c) & 1u); }

    for (unsigned s = lane; s < ns; s += 32) m[0][s] = 63u;
    __syncwarp();
    if (lane == 0) m[0][0] = 0u;
    __syncwarp();

    // The received bits are consumed in order, so they are read as 32-bit big-endian words, the next word always
    // requested one word ahead: a byte load per trellis step put a global-memory round trip on every step's critical
    // path (this kernel was 15x slower per frame than the K = 7 one, profiles/r01_notes.md v29).
    const unsigned enc_words = (io.enc_len + 3u) >> 2;
    auto load_word = [&](unsigned w) -> unsigned {
        if (w >= enc_words) return 0u;
        const unsigned char *p = io.src + 4u * w;
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) v = (v << 8) | ((4u * w + k < io.enc_len) ? (unsigned)p[k] : 0u);
        return v;
    };
    unsigned wbase = 0, wcur = load_word(0), wnext = load_word(1);
    auto take_bit = [&](unsigned ib) -> unsigned {               // ib never decreases
        if (SOFT) {                                              // four soft bytes per word, first byte in the top bits
            while ((ib >> 2) != wbase) { ++wbase; wcur = wnext; wnext = load_word(wbase + 1); }
            return (wcur >> (24u - 8u * (ib & 3u))) & 0xffu;
        }
        while ((ib >> 5) != wbase) { ++wbase; wcur = wnext; wnext = load_word(wbase + 1); }
        return ((wcur >> (31u - (ib & 31u))) & 1u) ? 255u : 0u;
    };
    // branch labels for this lane's butterflies (state pair i, i+half -> 2i, 2i+1)
    unsigned cur = 0;
    for (unsigned t = 0; t < T; ++t) {
        const unsigned col = t % cs.P;
        unsigned ib = (t / cs.P) * per + pre[col];
        unsigned sym0 = 127u, sym1 = 127u;
        if ((cs.keep0 >> col) & 1u) { sym0 = take_bit(ib); ++ib; }
        if ((cs.keep1 >> col) & 1u) { sym1 = take_bit(ib); }
        const unsigned *mo = m[cur];
        unsigned *mn = m[cur ^ 1];
        for (unsigned q = 0; q < half; q += 32) {
            const unsigned i = q + lane;
            const unsigned b0 = (__popc((2 * i) & cs.poly0) & 1) ? 255u : 0u;
            const unsigned b1 = (__popc((2 * i) & cs.poly1) & 1) ? 255u : 0u;
            const unsigned metric = (b0 ^ sym0) + (b1 ^ sym1);
            const unsigned a0 = mo[i], a1 = mo[i + half];
            unsigned m0 = a0 + metric, m1 = a1 + (510u - metric);
            const unsigned d0 = (int)(m0 - m1) > 0;
            mn[2 * i] = d0 ? m1 : m0;
            m0 = a0 + (510u - metric); m1 = a1 + metric;
            const unsigned d1 = (int)(m0 - m1) > 0;
            mn[2 * i + 1] = d1 ? m1 : m0;
            const unsigned w0 = __ballot_sync(0xffffffffu, d0), w1 = __ballot_sync(0xffffffffu, d1);
            // decision layout: word 2*(q/32) holds even states 2(q+lane), word 2*(q/32)+1 the odd ones
            if (lane == 0) { dec[(size_t)t * words + 2 * (q >> 5)] = w0; dec[(size_t)t * words + 2 * (q >> 5) + 1] = w1; }
        }
        __syncwarp();
        cur ^= 1;
    }
    __syncwarp();
    __threadfence_block();

    // traceback from state 0: every lane tracks the state; decision words arrive by shuffle
    unsigned char *out = io.dst;
    const unsigned steps_per_blk = 32 / words;
    unsigned state = 0, byte_acc = 0;
    for (long long t_hi = (long long)T - 1; t_hi >= 0; t_hi -= steps_per_blk) {
        // lanes hold words for steps t_hi, t_hi-1, ... : lane l -> step t_hi - l / words, word l % words
        const long long my_t = t_hi - (long long)(lane / words);
        unsigned w = 0;
        if (my_t >= 0) w = dec[(size_t)my_t * words + (lane % words)];
        for (unsigned k = 0; k < steps_per_blk; ++k) {
            const long long t = t_hi - k;
            if (t < 0) break;
            // state s: pair index i = s >> 1 lives in ballot group q = i / 32 at bit i % 32, odd/even selects the word
            const unsigned i = state >> 1, wi = 2 * (i >> 5) + (state & 1u);
            const unsigned word = __shfl_sync(0xffffffffu, w, k * words + wi);
            const unsigned bit = (word >> (i & 31u)) & 1u;
            if (t >= (long long)(cs.K - 1)) {
                const unsigned bi = (unsigned)t - (cs.K - 1);
                byte_acc |= bit << (7 - (bi & 7u));
                if ((bi & 7u) == 0) { if (lane == 0) out[bi >> 3] = (unsigned char)byte_acc; byte_acc = 0; }
            }
            state = (state >> 1) | (bit << (cs.K - 2));
        }
    }
}


// ------------------------------------------------------------------ Viterbi K=7, four lanes per codeword
// A thread-per-codeword kernel (all 64 metrics in one thread's registers, 281 instructions per step) is limited
// by how many codewords there are: a warp issues about every fourth cycle and 31 k codewords are 1.6 warps per
// scheduler (8.9 ms, profiles/r01_notes.md).  Here the 64 path metrics of a codeword are spread over
// 4 lanes x 16 registers in a CONSTANT-GEOMETRY layout: the metric of state s at step t
// lives at position p = rotr6(s, t mod 6) (lane = p >> 4, register = p & 15).  The two predecessors of a
// butterfly then always sit at positions that differ in bit k = 5 - (t mod 6) and its two successors are
// written back in place, so four steps out of six are register-only and two exchange with one other lane
// through one shuffle per metric.  Branch labels are parities of position bits: the register part is a
// compile-time constant, the lane part a per-thread mask XORed into the received symbols once per step.
// Decisions are stored per position ([step pair][codeword][lane] 32-bit words: the lane's 16 positions of the even
// step in the high nibbles of the four bytes, of the odd step in the low nibbles); the traceback works in position
// coordinates.  Same integer metrics, comparisons and tie-breaks as the specification.
constexpr int kV4Lanes = 4, kV4PosBits = 4;

__host__ __device__ constexpr unsigned v4_par6(unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1u; }
// position bits that enter the branch label of polynomial `poly` in phase r (label = parity(2 i & poly), 2 i = rotl6(p, r + 1) & ~1)
__host__ __device__ constexpr unsigned v4_phase_mask(unsigned poly, int r)
{
    unsigned m = 0;
    for (int q = 0; q < 6; ++q) {
        const int tpos = (q + r + 1) % 6;
        if (tpos != 0 && ((poly >> tpos) & 1u)) m |= 1u << q;
    }
    return m;
}
__host__ __device__ constexpr unsigned v4_label(int r, unsigned pos_bits)
{
    return v4_par6(pos_bits & v4_phase_mask(0x6d, r)) | (v4_par6(pos_bits & v4_phase_mask(0x4f, r)) << 1);
}

// Packed 16-bit path metrics (profiles/r01_notes.md v13).  A lane keeps its 16 positions in eight registers, position j
// in the low half of word j & 7 when j < 8 and in the high half otherwise, so one VIADD.16x2 / VIMNMX.U16x2 serves two
// positions.  16 bits are enough because the metric spread of a K = 7 trellis is at most 6 * 510 (every state is
// reachable from the best state of six steps ago and metrics never decrease): the common minimum is subtracted every 96
// steps, which leaves 3060 + 96 * 510 < 65536 and changes no comparison.  Same integer metrics relative to each other,
// same comparisons and tie-breaks as the specification's 32-bit decoder.
__device__ __forceinline__ unsigned v2_add(unsigned a, unsigned b) { return __vadd2(a, b); }
__device__ __forceinline__ unsigned v2_min(unsigned a, unsigned b) { return __vminu2(a, b); }

// packed branch metrics by label for one step: Ap[lab] = metric(lab) in the low half (positions 0..7 of the lane) and
// metric(lab ^ delta) in the high half (positions 8..15, whose labels differ by the constant delta of the phase)
__host__ __device__ inline uint4 v4_metrics(int r, unsigned s0, unsigned s1)
{
    const unsigned delta = v4_label(r, 8u);
    unsigned A[4];
    A[0] = s0 + s1; A[1] = (s0 ^ 255u) + s1; A[2] = s0 + (s1 ^ 255u); A[3] = 510u - A[0];
    return make_uint4(A[0] | (A[0 ^ delta] << 16), A[1] | (A[1 ^ delta] << 16), A[2] | (A[2 ^ delta] << 16), A[3] | (A[3 ^ delta] << 16));
}

// One trellis step.  G[w] receives, in its two sign bits (15, 31), the RAW decision flags of positions w and w + 8:
// a candidate difference minus one, whose sign is the complement of "first candidate > second" for the position that
// keeps successor 2i and the condition itself for the position that keeps 2i + 1 in the exchange phases; the caller
// undoes that with a constant XOR on the packed word.
template <int R>
__device__ __forceinline__ void v4_step(unsigned (&W)[8], const uint4 c, unsigned role, unsigned (&G)[8])
{
    constexpr int k = 5 - R;                       // position bit that separates the two predecessors
    const unsigned Ap[4] = { c.x, c.y, c.z, c.w };
    if constexpr (k < 3) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w & (1 << k)) continue;
            const int w1 = w | (1 << k);
            const unsigned lab = v4_label(R, (unsigned)w);
            const unsigned a = Ap[lab], b = Ap[3 - lab];
            // candidate sums never overflow a half (renormalisation keeps them below 65536), so a plain 32-bit add is
            // a packed add -- and may issue on the FMA pipe (IMAD.IADD), which the ALU-bound loop leaves idle
            const unsigned m0 = W[w] + a, m1 = W[w1] + b;                    // into successor 2i   (kept at w)
            const unsigned q0 = W[w] + b, q1 = W[w1] + a;                    // into successor 2i+1 (kept at w1)
            G[w] = v2_add(m0, ~m1);                                          // m0 - m1 - 1: sign clear <=> m0 > m1
            G[w1] = v2_add(q0, ~q1);
            W[w] = v2_min(m0, m1);
            W[w1] = v2_min(q0, q1);
        }
    } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned lab = v4_label(R, (unsigned)w);
            const unsigned a = Ap[lab], b = Ap[3 - lab];
            unsigned other;
            if constexpr (k == 3) other = __byte_perm(W[w], 0u, 0x1032);
            else other = __shfl_xor_sync(0xffffffffu, W[w], 1 << (k - kV4PosBits));
            const unsigned c_own = W[w] + a, c_oth = other + b;
            // F = own - oth - 1.  Lower position (keeps 2i): decision own > oth <=> sign(F) clear.
            // Upper position (keeps 2i+1): decision oth > own <=> F + 1 < 0 <=> sign(F + 1) set.
            G[w] = v2_add(v2_add(c_own, ~c_oth), role);
            W[w] = v2_min(c_own, c_oth);
        }
    }
}

// sign bits of eight packed words -> the high nibble of each byte of one word: position p = w + 8 h lands in byte
// 2 (w & 1) + h, bit 7 - (w >> 1); the low nibbles are garbage
__device__ __forceinline__ unsigned v4_pack(const unsigned (&G)[8])
{
    const unsigned p0 = __byte_perm(G[0], G[1], 0x7531), p1 = __byte_perm(G[2], G[3], 0x7531);
    const unsigned p2 = __byte_perm(G[4], G[5], 0x7531), p3 = __byte_perm(G[6], G[7], 0x7531);
    unsigned r = (p0 & 0x80808080u) | ((p1 >> 1) & ~0x80808080u);
    r = (r & 0xc0c0c0c0u) | ((p2 >> 2) & ~0xc0c0c0c0u);
    r = (r & 0xe0e0e0e0u) | ((p3 >> 3) & ~0xe0e0e0e0u);
    return r;
}

// Traceback coordinates.  The walk keeps u, a bit permutation of the POSITION of the current state (position =
// rotr6(state, phase)): in position coordinates one traceback step replaces a single bit by the decision bit, and u is
// laid out so that the decision can be fetched from the pair word with byte permutes:
//   u5 u4 = lane that owns the position, u3 u2 = (p0, p3) = byte of that lane's word, u1 u0 = ~(p2 p1) = bit in the nibble.
__host__ __device__ constexpr int v4_ubit(int pos_bit) { return pos_bit == 0 ? 3 : pos_bit == 1 ? 0 : pos_bit == 2 ? 1 : pos_bit == 3 ? 2 : pos_bit; }
__host__ __device__ constexpr bool v4_uinv(int pos_bit) { return pos_bit == 1 || pos_bit == 2; }
constexpr unsigned kV4UZero = 3u;                    // u of position 0

// one traceback step at chunk offset J (step = 24 c + J): reads the decision of the current position from the pair
// word, shifts it into the output accumulator and replaces the position bit that leaves the state
template <int J>
__device__ __forceinline__ void v4_back(unsigned &u, unsigned &acc, const uint4 &pw)
{
    constexpr int ph = (J + 1) % 6;                 // rotation of this step
    constexpr int kb = (6 - ph) % 6;                // position bit replaced by the decision
    constexpr int ub = v4_ubit(kb);
    const unsigned sel = (u >> 2) & 7u;
    const unsigned lo = __byte_perm(pw.x, pw.y, sel), hi = __byte_perm(pw.z, pw.w, sel);
    const unsigned byte = __byte_perm(lo, hi, (u >> 3) & 4u);
    // even steps sit in the high nibbles
    const unsigned raw = byte >> ((J & 1) ? (u & 3u) : ((u & 3u) | 4u));       // decision in bit 0, garbage above
    acc = __funnelshift_r(acc, raw, 1);
    const unsigned t = raw << ub;
    if (v4_uinv(kb)) u = (u & ~(1u << ub)) | (~t & (1u << ub));
    else u = (u & ~(1u << ub)) | (t & (1u << ub));
}

constexpr int kV4Threads = 64;
constexpr unsigned kV4Renorm = 96;                    // steps between metric renormalisations (multiple of 6)
constexpr int kV4Warm = 96;                           // speculative traceback warm-up, steps (debug override: LQB_V4_WARM)

// SOFT: the received values are bytes (one per kept coded bit, deinterleaved: LQB_RX_SOFT, csrc/lqb_rx_soft.cu) instead of
// bits; the packed branch metrics are then computed per step from the two bytes (v4_metrics is plain arithmetic on them)
// instead of being looked up by the received pair.  Everything else -- metrics layout, decisions, traceback -- is shared.
template <bool PUNCT, bool SOFT = false>
__global__ void __launch_bounds__(kV4Threads)
k_viterbi27x4(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list, int stage, unsigned *__restrict__ dec, int warm)
{
    // [lane][phase][received pair] -> v4_metrics; a received symbol is 0, 1 or (punctured codes) 2 = erased
    constexpr int kPairs = PUNCT ? 9 : 4;
    __shared__ uint4 lut[kV4Lanes][6][kPairs];
    const unsigned gt = blockIdx.x * kV4Threads + threadIdx.x;
    const unsigned l = gt % kV4Lanes;
    unsigned gi = gt / kV4Lanes;
    // groups past the end of the list shadow the last codeword without storing anything, so that every lane of a
    // warp runs the same number of steps and the shuffles can use the full mask
    const bool active = gi < n_list;
    if (!active) gi = n_list - 1;
    const FrameDesc &d = P.frames[list[gi]];
    StageIO io = stage_io(P, d, stage);
    if (SOFT) { io.src = P.soft_d + P.soft[list[gi]].d_off; io.enc_len *= 8u; }       // one byte per coded bit
    const ConvSpec cs = conv_spec(io.fs);
    const unsigned nbits = 8 * io.dec_len, T = nbits + 6;          // T is even
    unsigned Tw = T;                                  // longest codeword in this warp
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) Tw = max(Tw, __shfl_xor_sync(0xffffffffu, Tw, m));
    const unsigned char *enc = io.src;
    const unsigned enc_words = max((io.enc_len + 3u) / 4u, 1u);

    // lane part of the branch labels, as XOR masks on the received symbols, per phase
    auto lane_mask = [](unsigned lane, int r, unsigned poly) { return (__popc((lane << kV4PosBits) & v4_phase_mask(poly, r)) & 1) ? 255u : 0u; };
    // SOFT: the twelve lane masks (six phases x two polynomials) as bits 2 r, 2 r + 1 of one word
    unsigned lmask12 = 0;
    if (SOFT) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
            lmask12 |= ((lane_mask(l, r, 0x6d) & 1u) << (2 * r)) | ((lane_mask(l, r, 0x4f) & 1u) << (2 * r + 1));
    }
    if (!SOFT) {
        const unsigned soft[3] = { 0u, 255u, 127u };
        for (unsigned i = threadIdx.x; i < kV4Lanes * 6 * kPairs; i += kV4Threads) {
            const unsigned ll = i / (6 * kPairs), v = i % kPairs;
            const int r = (int)((i / kPairs) % 6u);
            const unsigned s0 = PUNCT ? soft[v / 3u] : ((v & 2u) ? 255u : 0u), s1 = PUNCT ? soft[v % 3u] : ((v & 1u) ? 255u : 0u);
            lut[ll][r][v] = v4_metrics(r, s0 ^ lane_mask(ll, r, 0x6d), s1 ^ lane_mask(ll, r, 0x4f));
        }
        __syncthreads();
    }
    // phases 0 and 1 exchange with the lane that differs in bit 1 / bit 0; the lane holding the upper predecessors
    // adds one to the raw difference (see v4_step) and its flags are stored uninverted
    const unsigned up5 = (l >> 1) & 1u, up4 = l & 1u;
    const unsigned role0 = up5 ? 0x00010001u : 0u, role1 = up4 ? 0x00010001u : 0u;
    const unsigned xm0 = (up5 ? 0u : 0xf0f0f0f0u) | (up4 ? 0u : 0x0f0f0f0fu);     // steps 6n, 6n+1
    constexpr unsigned xm1 = 0x00f000f0u | 0x0f0f0f0fu;                           // steps 6n+2 (halves exchange), 6n+3
    constexpr unsigned xm2 = 0xffffffffu;                                         // steps 6n+4, 6n+5
    unsigned W[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) W[j] = 63u | (63u << 16);
    if (l == 0) W[0] = 63u << 16;                   // state 0 sits at position 0 at t = 0

    const unsigned *enc32 = reinterpret_cast<const unsigned *>(enc);     // byte arenas are 16-byte aligned per frame
    // 32 encoded bits from bit offset o (MSB first, in the top bits); six steps use at most 12 of them
    auto fetch32 = [&](unsigned o) {
        const unsigned idx = o >> 5;
        const unsigned a = __byte_perm(__ldg(enc32 + min(idx, enc_words - 1u)), 0u, 0x0123);
        const unsigned b = __byte_perm(__ldg(enc32 + min(idx + 1u, enc_words - 1u)), 0u, 0x0123);
        return __funnelshift_l(b, a, o & 31u);
    };
    // unpunctured: the window of the next group is fetched one group ahead; punctured: the bit position of the next group
    // depends on the puncturing column, the window is fetched at the group's start
    unsigned bits = 0, bits_next = (PUNCT || SOFT) ? 0u : fetch32(0u);
    unsigned bitpos = 0, col = 0;                   // punctured: encoded bits consumed, column t mod P
    const char *lut_l = reinterpret_cast<const char *>(&lut[l][0][0]);
    // SOFT, unpunctured: the twelve bytes of a group of six steps (three words, byte 0 of the stream in the low byte)
    unsigned sw0 = 0, sw1 = 0, sw2 = 0;
    const unsigned soft_words = max((io.enc_len + 3u) / 4u, 1u);
    auto soft_byte = [&](unsigned i) -> unsigned { return i < io.enc_len ? (unsigned)__ldg(enc + i) : 0u; };
    auto metrics = [&](const int r) -> uint4 {
        if (SOFT) {
            unsigned s0, s1;
            if (PUNCT) {
                const unsigned k0 = (cs.keep0 >> col) & 1u, k1 = (cs.keep1 >> col) & 1u;
                s0 = k0 ? soft_byte(bitpos) : 127u;
                s1 = k1 ? soft_byte(bitpos + k0) : 127u;
                bitpos += k0 + k1;
                col = (col + 1u == cs.P) ? 0u : col + 1u;
            } else {
                const unsigned w = (r < 2) ? sw0 : (r < 4) ? sw1 : sw2;
                s0 = (w >> (16 * (r & 1))) & 0xffu;
                s1 = (w >> (16 * (r & 1) + 8)) & 0xffu;
            }
            // (an erased position enters as 127 and takes the label mask like any other value, exactly as in the table of
            // the hard-input form: soft[] = 0, 255, 127 XOR mask)
            const unsigned m0 = ((lmask12 >> (2 * r)) & 1u) ? 255u : 0u, m1 = ((lmask12 >> (2 * r + 1)) & 1u) ? 255u : 0u;
            return v4_metrics(r, s0 ^ m0, s1 ^ m1);
        }
        if (PUNCT) {
            const unsigned k0 = (cs.keep0 >> col) & 1u, k1 = (cs.keep1 >> col) & 1u;
            const unsigned c0 = k0 ? bits >> 31 : 2u;
            bits <<= k0;
            const unsigned c1 = k1 ? bits >> 31 : 2u;
            bits <<= k1;
            bitpos += k0 + k1;
            col = (col + 1u == cs.P) ? 0u : col + 1u;
            return *reinterpret_cast<const uint4 *>(lut_l + 16 * kPairs * r + 16 * (3u * c0 + c1));
        } else {
            const int sh = 26 - 2 * r;                 // the step's two bits -> byte offset 16 * v
            const unsigned off = (bits >> sh) & 0x30u;
            return *reinterpret_cast<const uint4 *>(lut_l + 64 * r + off);
        }
    };
    // decisions: [step pair][codeword][lane] 32-bit words, even step in the high nibbles, odd step in the low ones
    unsigned *out_dec = dec + (size_t)gi * kV4Lanes + l;
    const size_t dstride = (size_t)n_list * kV4Lanes;
    unsigned t = 0, since = 0;
#define LQB_V4_PAIR(R, ROLE_E, ROLE_O, XM)                                        \
    {                                                                             \
        unsigned G[8];                                                            \
        v4_step<R>(W, metrics(R), ROLE_E, G);                                     \
        const unsigned re = v4_pack(G);                                           \
        v4_step<R + 1>(W, metrics(R + 1), ROLE_O, G);                             \
        const unsigned ro = v4_pack(G);                                           \
        const unsigned dd = ((re & 0xf0f0f0f0u) | ((ro >> 4) & 0x0f0f0f0fu)) ^ (XM); \
        if (active && t + R < T) out_dec[(size_t)((t + R) >> 1) * dstride] = dd;  \
    }
    // whole groups of six phases; steps past a codeword's end run on arbitrary symbols and store nothing
    for (; t < Tw; t += 6) {
        if (SOFT) {
            if (!PUNCT) {
                const unsigned wi = 3u * (t / 6u);            // bytes 2 t .. 2 t + 11
                sw0 = __ldg(enc32 + min(wi, soft_words - 1u)); sw1 = __ldg(enc32 + min(wi + 1u, soft_words - 1u)); sw2 = __ldg(enc32 + min(wi + 2u, soft_words - 1u));
            }
        }
        else if (!PUNCT) { bits = bits_next; bits_next = fetch32(2u * (t + 6u)); }
        else bits = fetch32(bitpos);
        LQB_V4_PAIR(0, role0, role1, xm0)
        LQB_V4_PAIR(2, 0x00010000u, 0u, xm1)
        LQB_V4_PAIR(4, 0u, 0u, xm2)
        since += 6;
        if (since == kV4Renorm) {
            since = 0;
            unsigned mn = v2_min(v2_min(v2_min(W[0], W[1]), v2_min(W[2], W[3])), v2_min(v2_min(W[4], W[5]), v2_min(W[6], W[7])));
            mn = v2_min(mn, __byte_perm(mn, 0u, 0x1032));
            mn = v2_min(mn, __shfl_xor_sync(0xffffffffu, mn, 1));
            mn = v2_min(mn, __shfl_xor_sync(0xffffffffu, mn, 2));
            const unsigned neg = __vsub2(0u, mn);
#pragma unroll
            for (int j = 0; j < 8; ++j) W[j] = v2_add(W[j], neg);
        }
    }
#undef LQB_V4_PAIR
    __syncwarp();

    // ---- traceback, four lanes per codeword.  Lane q owns output bytes [B_q, B_q+1), i.e. steps [lo, hi) with
    // lo = 8 B_q + 6: it starts kV4Warm..kV4Warm+23 steps above hi from an arbitrary state (the top lane: at T - 1 from
    // state 0) and walks down to lo.  Survivor paths merge within a few constraint lengths, so the state a lane has
    // when it reaches hi nearly always equals the state the lane above ends with; that is CHECKED afterwards, top
    // down, and a lane whose start was wrong walks its segment again from the proven state - the output is exactly
    // the serial traceback's.  Steps are processed in chunks of 24 (phase, byte and pair alignment all repeat).
    const int Ti = (int)T;
    const int lo = 8 * (int)((io.dec_len * l) >> 2) + 6, hi = 8 * (int)((io.dec_len * (l + 1u)) >> 2) + 6;
    const uint4 *dec128 = reinterpret_cast<const uint4 *>(dec) + gi;
    unsigned char *out = io.dst;
    unsigned ustart = kV4UZero, uend = kV4UZero;
    auto load4 = [&](uint4 (&w)[4], int ts_top) {     // pair words of steps ts_top, ts_top - 2, .. (ts_top odd)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int ts = ts_top - 2 * kk;
            w[kk] = (ts >= 0 && ts < Ti) ? dec128[(size_t)(ts >> 1) * n_list] : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    // walk chunks c_top .. lo / 24; when the walk reaches step force_ts - 1 its state is replaced by u_force
    auto walk = [&](int c_top, int force_ts, unsigned u_force) {
        unsigned u = kV4UZero, acc = 0;
        uint4 w[4], wn[4];
        load4(w, 24 * c_top + 23);
        const int c_bot = lo / 24;
        for (int c = c_top; c >= c_bot; --c) {
            const int base = 24 * c;
            // before step base + J (J = 21, 13, 5): segment boundaries hi / lo / force_ts all lie at base + J + 1
#define LQB_V4_MARK(J)                                                             \
            {                                                                      \
                const int tb = base + (J) + 1;                                     \
                if (tb == force_ts) u = u_force;                                   \
                if (tb == hi) ustart = u;                                          \
                if (tb == lo) uend = u;                                            \
            }
            // after step base + J (J = 22, 14, 6) the accumulator's top byte is output byte (base + J - 6) / 8
#define LQB_V4_EMIT(J)                                                             \
            {                                                                      \
                const int ts = base + (J);                                         \
                if (active && ts >= lo && ts < hi) out[(ts - 6) >> 3] = (unsigned char)(acc >> 24); \
            }
            load4(wn, base + 15);
            v4_back<23>(u, acc, w[0]); v4_back<22>(u, acc, w[0]); LQB_V4_EMIT(22) LQB_V4_MARK(21)
            v4_back<21>(u, acc, w[1]); v4_back<20>(u, acc, w[1]);
            v4_back<19>(u, acc, w[2]); v4_back<18>(u, acc, w[2]);
            v4_back<17>(u, acc, w[3]); v4_back<16>(u, acc, w[3]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) w[kk] = wn[kk];
            load4(wn, base + 7);
            v4_back<15>(u, acc, w[0]); v4_back<14>(u, acc, w[0]); LQB_V4_EMIT(14) LQB_V4_MARK(13)
            v4_back<13>(u, acc, w[1]); v4_back<12>(u, acc, w[1]);
            v4_back<11>(u, acc, w[2]); v4_back<10>(u, acc, w[2]);
            v4_back<9>(u, acc, w[3]); v4_back<8>(u, acc, w[3]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) w[kk] = wn[kk];
            load4(wn, base - 1);
            v4_back<7>(u, acc, w[0]); v4_back<6>(u, acc, w[0]); LQB_V4_EMIT(6) LQB_V4_MARK(5)
            v4_back<5>(u, acc, w[1]); v4_back<4>(u, acc, w[1]);
            v4_back<3>(u, acc, w[2]); v4_back<2>(u, acc, w[2]);
            v4_back<1>(u, acc, w[3]); v4_back<0>(u, acc, w[3]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) w[kk] = wn[kk];
#undef LQB_V4_MARK
#undef LQB_V4_EMIT
        }
    };
    const int c_last = (Ti - 1) / 24;                // chunk of the last step
    const bool top = hi + warm >= Ti;                // this lane's walk starts at the codeword's end: exact
    walk(top ? c_last : (hi + warm) / 24, top ? Ti : -1, kV4UZero);
    // verify top down; `ok` = this lane's segment is proven equal to the serial traceback's
    bool ok = top;
#pragma unroll 1
    for (int qq = kV4Lanes - 2; qq >= 0; --qq) {
        const unsigned need = __shfl_down_sync(0xffffffffu, uend, 1);       // end state of the lane above (same codeword for l < 3)
        const bool okup = __shfl_down_sync(0xffffffffu, (int)ok, 1) != 0;
        if ((int)l == qq && !ok) {
            // the lane above is proven by now (okup) - its end state is the true state at step hi
            if (okup && ustart != need) {
                if (hi > lo) walk((hi - 1) / 24, hi, need);
                else uend = need;                    // empty segment: pass the proven state on
            }
            ok = true;
        }
    }
}

// ------------------------------------------------------------------ Viterbi K=9, sixteen lanes per codeword
// The 256 path metrics of a codeword sit in 16 lanes x 8 registers of packed 16-bit pairs: lane l holds the OLD states
// 16 q + l (low half of word q) and 16 (q + 8) + l (high half), q = 0 .. 7 -- the two predecessors i and i + 128 of
// butterfly i = 16 q + l share a word, so the add-compare-select is lane-local:
//   X = W + (m | 510 - m << 16),  Y = W + (510 - m | m << 16)      candidate sums into new states 2 i and 2 i + 1
//   Z1 = (X.lo | Y.lo << 16), Z2 = (X.hi | Y.hi << 16)             the first / second candidate of both
//   N = min16x2(Z1, Z2),  decision = Z1 > Z2 (sign of Z1 - Z2 - 1, per half)
// with m the branch metric of the butterfly's label: label(i) = label(16 q) ^ label(l), the lane part XORed into the
// received values once per step, the register part a compile-time constant.  The new states 32 q + 2 l + x then move to
// where the next step wants them (lane 2 (l & 7) + x, register 2 q + (l >> 3)): sixteen shuffles and eight byte permutes.
// 16 bits suffice: the metric spread of a K = 9 trellis is at most 8 * 510 and the common minimum is subtracted every 96
// steps (4080 + 96 * 510 < 65536); relative metrics, comparisons and tie-breaks are the specification's.
// Decisions: one 32-bit word per lane and PAIR of steps (even step in the high nibbles of the four bytes, odd step in
// the low ones; the bit of new state 32 q + 2 l + x: byte 2 (q & 1) + x, bit 7 - (q >> 1)), [pair][lane] in the frame's
// private slice of the decision arena.  Traceback: the sixteen lanes walk the state together, the lane that owns it
// hands its bit over by shuffle.  Hard or SOFT input, punctured or not (the generic symbol fetch).  Two codewords per warp.
constexpr int kV9Lanes = 16;
constexpr int kV9Threads = 128;

__host__ __device__ constexpr unsigned v9_par(unsigned x) { x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1u; }
// label of butterfly i: bit 0 = parity(2 i & poly0), bit 1 = parity(2 i & poly1)   (K = 9: 0x1af, 0x11d)
__host__ __device__ constexpr unsigned v9_label(unsigned i) { return v9_par((2u * i) & 0x1afu) | (v9_par((2u * i) & 0x11du) << 1); }

template <bool SOFT>
__global__ void __launch_bounds__(kV9Threads)
k_viterbi29x16(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list, int stage)
{
    const unsigned lane = threadIdx.x & 31u, l = lane & 15u, grp_base = lane & 16u;
    unsigned gi = (blockIdx.x * kV9Threads + threadIdx.x) / kV9Lanes;
    // a group past the end of the list shadows the last codeword without storing anything (full-mask shuffles)
    const bool active = gi < n_list;
    if (!active) gi = n_list - 1;
    const FrameDesc &d = P.frames[list[gi]];
    StageIO io = stage_io(P, d, stage);
    if (SOFT) { io.src = P.soft_d + P.soft[list[gi]].d_off; io.enc_len *= 8u; }     // one byte per coded bit
    const ConvSpec cs = conv_spec(io.fs);
    const unsigned nbits = 8 * io.dec_len, T = nbits + 8;          // T is even
    unsigned Tw = T;                                  // longest codeword in this warp
    Tw = max(Tw, __shfl_xor_sync(0xffffffffu, Tw, 16));
    unsigned *dec = reinterpret_cast<unsigned *>(P.decisions + d.dec_off);        // [pair][lane]

    // received values, in order (as the warp-per-codeword kernel reads them: words, one ahead)
    unsigned per = 0, pre[8];
    for (unsigned c = 0; c < cs.P; ++c) { pre[c] = per; per += ((cs.keep0 >> c) & 1u) + ((cs.keep1 >> c) & 1u); }
    const unsigned enc_words = (io.enc_len + 3u) >> 2;
    auto load_word = [&](unsigned w) -> unsigned {
        if (w >= enc_words) return 0u;
        const unsigned char *p = io.src + 4u * w;
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) v = (v << 8) | ((4u * w + k < io.enc_len) ? (unsigned)p[k] : 0u);
        return v;
    };
    unsigned wbase = 0, wcur = load_word(0), wnext = load_word(1);
    auto take = [&](unsigned ib) -> unsigned {                   // ib never decreases
        if (SOFT) {
            while ((ib >> 2) != wbase) { ++wbase; wcur = wnext; wnext = load_word(wbase + 1); }
            return (wcur >> (24u - 8u * (ib & 3u))) & 0xffu;
        }
        while ((ib >> 5) != wbase) { ++wbase; wcur = wnext; wnext = load_word(wbase + 1); }
        return ((wcur >> (31u - (ib & 31u))) & 1u) ? 255u : 0u;
    };
    const unsigned lab_l = v9_label(l);
    const unsigned lm0 = (lab_l & 1u) ? 255u : 0u, lm1 = (lab_l & 2u) ? 255u : 0u;

    unsigned W[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) W[q] = 63u | (63u << 16);
    if (l == 0) W[0] = 63u << 16;                    // state 0 starts at 0
    const unsigned src0 = grp_base | (l >> 1), src1 = src0 + 8u;     // the two lanes this lane's next states come from
    const unsigned sel = (l & 1u) ? 0x7632u : 0x5410u;               // their half: x = l & 1
    unsigned col = 0, tcol = 0, since = 0, pair_word = 0;
    for (unsigned t = 0; t < Tw; ++t) {
        // ---- received pair and the four packed branch metrics
        unsigned ib = tcol * per + pre[col];
        unsigned sym0 = 127u, sym1 = 127u;
        if (t < T) {
            if ((cs.keep0 >> col) & 1u) { sym0 = take(ib); ++ib; }
            if ((cs.keep1 >> col) & 1u) { sym1 = take(ib); }
        }
        if (++col == cs.P) { col = 0; ++tcol; }
        const unsigned s0 = sym0 ^ lm0, s1 = sym1 ^ lm1;
        unsigned M[4];
        M[0] = s0 + s1; M[1] = (s0 ^ 255u) + s1; M[2] = s0 + (s1 ^ 255u); M[3] = 510u - M[0];
        const unsigned A[4] = { M[0] | (M[3] << 16), M[1] | (M[2] << 16), M[2] | (M[1] << 16), M[3] | (M[0] << 16) };
        // ---- add-compare-select, lane-local
        unsigned N[8], G[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            constexpr unsigned kLab[8] = { v9_label(0), v9_label(16), v9_label(32), v9_label(48), v9_label(64), v9_label(80), v9_label(96), v9_label(112) };
            const unsigned X = W[q] + A[kLab[q]], Y = W[q] + A[3u - kLab[q]];
            const unsigned Z1 = __byte_perm(X, Y, 0x5410), Z2 = __byte_perm(X, Y, 0x7632);
            G[q] = v2_add(Z1, ~Z2);                  // sign clear <=> Z1 > Z2 <=> the predecessor i + 128 wins
            N[q] = v2_min(Z1, Z2);                   // new states 2 i (low) and 2 i + 1 (high)
        }
        // ---- decisions: sign bits of the eight words -> the high nibbles of four bytes (byte 2 (q & 1) + x, bit 7 - (q >> 1))
        {
            const unsigned p0 = __byte_perm(G[0], G[1], 0x7531), p1 = __byte_perm(G[2], G[3], 0x7531);
            const unsigned p2 = __byte_perm(G[4], G[5], 0x7531), p3 = __byte_perm(G[6], G[7], 0x7531);
            unsigned r = (p0 & 0x80808080u) | ((p1 >> 1) & 0x40404040u) | ((p2 >> 2) & 0x20202020u) | ((p3 >> 3) & 0x10101010u);
            r ^= 0xf0f0f0f0u;                        // stored sense: 1 = the predecessor i + 128 wins
            if (t & 1u) {
                pair_word |= r >> 4;
                if (active && t < T) dec[(size_t)(t >> 1) * kV9Lanes + l] = pair_word;
            } else pair_word = r;
        }
        // ---- the new states go where the next step reads them
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned a0 = __shfl_sync(0xffffffffu, N[j], src0), a4 = __shfl_sync(0xffffffffu, N[j + 4], src0);
            const unsigned b0 = __shfl_sync(0xffffffffu, N[j], src1), b4 = __shfl_sync(0xffffffffu, N[j + 4], src1);
            W[2 * j] = __byte_perm(a0, a4, sel);     // register 2 j     <- butterflies j, j + 4 of lane (l >> 1)
            W[2 * j + 1] = __byte_perm(b0, b4, sel); // register 2 j + 1 <- the same of lane (l >> 1) + 8
        }
        if (++since == 96u) {
            since = 0;
            unsigned mn = v2_min(v2_min(v2_min(W[0], W[1]), v2_min(W[2], W[3])), v2_min(v2_min(W[4], W[5]), v2_min(W[6], W[7])));
            mn = v2_min(mn, __byte_perm(mn, 0u, 0x1032));
#pragma unroll
            for (int m = 1; m <= 8; m <<= 1) mn = v2_min(mn, __shfl_xor_sync(0xffffffffu, mn, m));
            const unsigned neg = __vsub2(0u, mn);
#pragma unroll
            for (int q = 0; q < 8; ++q) W[q] = v2_add(W[q], neg);
        }
    }
    __syncwarp();
    __threadfence_block();

    // ---- traceback from state 0: the sixteen lanes track the state, the owner of its bit hands it over
    unsigned char *out = io.dst;
    unsigned state = 0, acc = 0;
    const int pairs = (int)(T >> 1);
    unsigned w_cur = active ? dec[(size_t)(pairs - 1) * kV9Lanes + l] : 0u;
    for (int pr = pairs - 1; pr >= 0; --pr) {
        const unsigned w_next = (active && pr > 0) ? dec[(size_t)(pr - 1) * kV9Lanes + l] : 0u;     // (requested one pair ahead)
#pragma unroll
        for (int odd = 1; odd >= 0; --odd) {
            const unsigned t = 2u * (unsigned)pr + (unsigned)odd;
            // new state `state` after step t = 32 q + 2 l' + x
            const unsigned x = state & 1u, owner = (state >> 1) & 15u, q = state >> 5;
            const unsigned byte = (w_cur >> (8u * (2u * (q & 1u) + x))) & 0xffu;
            const unsigned mine = (byte >> ((odd ? 3u : 7u) - (q >> 1))) & 1u;
            const unsigned k = __shfl_sync(0xffffffffu, mine, grp_base | owner);
            if (t >= 8u) {
                const unsigned bi = t - 8u;          // the bit shifted out at step t entered at t - 8
                acc |= k << (7u - (bi & 7u));
                if ((bi & 7u) == 0u) { if (active && l == 0) out[bi >> 3] = (unsigned char)acc; acc = 0; }
            }
            state = (state >> 1) | (k << 7);
        }
        w_cur = w_next;
    }
}

// ------------------------------------------------------------------ Reed-Solomon (warp per 255-byte block)
constexpr int kRsWarps = 4;
constexpr int kRsSynBytes = 256 * 32 * 4;

__global__ void __launch_bounds__(32 * kRsWarps)
k_rs(PayloadParams P, const unsigned *__restrict__ blocks, unsigned n_blocks, int stage)
{
    // syn[v][lane]: lane's bank is its own, so a lookup with a different v in every lane is one conflict-free wavefront
    extern __shared__ uint4 rs_dyn[];
    unsigned (*syn)[32] = reinterpret_cast<unsigned (*)[32]>(rs_dyn);
    __shared__ unsigned char gexp[512], glog[256];
    __shared__ __align__(16) unsigned char data[kRsWarps][256];
    __shared__ unsigned char synd[kRsWarps][32], lam[kRsWarps][36], omg[kRsWarps][32];
    __shared__ int deg_s[kRsWarps];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) gexp[i] = P.tables->gf_exp[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) glog[i] = P.tables->gf_log[i];
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(&P.tables->rs_syn[0][0]);
        for (int i = threadIdx.x; i < 256 * 32 / 4; i += blockDim.x) rs_dyn[i] = g[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warps stride over the block list so that the 32 KB table is loaded once per CTA, not once per four blocks
    for (unsigned gi = blockIdx.x * kRsWarps + warp; gi < n_blocks; gi += gridDim.x * kRsWarps) {
    const FrameDesc &d = P.frames[blocks[2 * gi]];
    const unsigned blk = blocks[2 * gi + 1];
    const StageIO io = stage_io(P, d, stage);
    const unsigned n = io.dec_len;
    const unsigned nblocks = (n + 222) / 223, dec_block = (n + nblocks - 1) / nblocks, enc_block = dec_block + 32, pad = 223 - dec_block;
    const unsigned nn = 255 - pad;
    unsigned char *x = data[warp];
    const unsigned char *src = io.src + (size_t)blk * enc_block;
    __syncwarp();
    for (unsigned i = lane; i < enc_block; i += 32) x[i] = src[i];
    __syncwarp();

    // syndrome lane: S_lane = r(b), b = alpha^(lane+1), by Horner from the highest-degree byte, four bytes per
    // dependent lookup: S <- S b^4 + x0 b^3 + x1 b^2 + x2 b + x3.  The running value lives in the top byte of `r`.
    unsigned r = 0;
    const char *tl = reinterpret_cast<const char *>(&syn[0][lane]);          // + 128 v
    auto T = [&](unsigned off) { return *reinterpret_cast<const unsigned *>(tl + off); };
    const unsigned ng = nn >> 2;
    const unsigned *xw = reinterpret_cast<const unsigned *>(x);
    for (unsigned g = 0; g < ng; ++g) {
        const unsigned w = xw[g];                                            // x0 | x1 << 8 | x2 << 16 | x3 << 24
        const unsigned ts = T((r >> 17) & 0x7f80u);
        const unsigned t0 = T((w << 7) & 0x7f80u), t1 = T((w >> 1) & 0x7f80u), t2 = T((w >> 9) & 0x7f80u);
        r = ts ^ (t0 << 8) ^ (t1 << 16) ^ (t2 << 24) ^ w;                    // only the top byte is meaningful
    }
    unsigned s = r >> 24;
    for (unsigned j = 4 * ng; j < nn; ++j) s = x[j] ^ (T(s << 7) & 0xffu);
    synd[warp][lane] = (unsigned char)s;
    const unsigned any = __ballot_sync(0xffffffffu, s != 0);
    __syncwarp();
    if (any) {
        unsigned char *S = synd[warp], *L = lam[warp], *O = omg[warp];
        // Berlekamp-Massey with one locator coefficient per lane (lane i holds lambda[i], b[i]; coefficient 32 is
        // carried by lane 31 in a second register).  Same recurrence as the serial specification: every
        // quantity is a GF(256) value, so the result is identical.
        {
            unsigned lam_i = (lane == 0) ? 1u : 0u, b_i = (lane == 0) ? 1u : 0u;     // coefficients 0..31
            unsigned lam32 = 0u, b32 = 0u;                                            // coefficient 32 (meaningful on lane 31)
            unsigned el = 0;
            for (unsigned r = 1; r <= 32; ++r) {
                // discrepancy = sum_{i<r} lambda[i] * s[r-i-1]
                unsigned term = 0;
                if ((unsigned)lane < r) {
                    const unsigned sv_ = S[r - lane - 1];
                    if (lam_i && sv_) term = gexp[glog[lam_i] + glog[sv_]];
                }
#pragma unroll
                for (int k = 16; k >= 1; k >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, k);
                const unsigned dsc = term;
                // x * b(x): coefficient i takes b[i-1]
                unsigned b_shift = __shfl_up_sync(0xffffffffu, b_i, 1);
                if (lane == 0) b_shift = 0;
                const unsigned b31 = __shfl_sync(0xffffffffu, b_i, 31);               // becomes coefficient 32
                if (dsc == 0) {
                    b_i = b_shift; b32 = b31;
                } else {
                    const unsigned ld = glog[dsc];
                    const unsigned t_i = lam_i ^ (b_shift ? gexp[ld + glog[b_shift]] : 0u);
                    const unsigned t32 = lam32 ^ (b31 ? gexp[ld + glog[b31]] : 0u);
                    if (2 * el <= r - 1) {
                        el = r - el;
                        const unsigned dinv = 255 - ld;
                        b_i = lam_i ? gexp[glog[lam_i] + dinv] : 0u;
                        b32 = lam32 ? gexp[glog[lam32] + dinv] : 0u;
                    } else {
                        b_i = b_shift; b32 = b31;
                    }
                    lam_i = t_i; lam32 = t32;
                }
            }
            L[lane] = (unsigned char)lam_i;
            if (lane == 31) L[32] = (unsigned char)lam32;
            __syncwarp();
            // degree and omega = (S * lambda) mod x^deg, one coefficient per lane
            unsigned nzmask = __ballot_sync(0xffffffffu, lam_i != 0);
            int deg = nzmask ? 31 - __clz(nzmask) : 0;
            if (L[32]) deg = 32;
            if (lane == 0) deg_s[warp] = deg;
            if (lane < deg) {
                unsigned acc = 0;
                for (int j = 0; j <= lane; ++j) if (S[lane - j] && L[j]) acc ^= gexp[glog[S[lane - j]] + glog[L[j]]];
                O[lane] = (unsigned char)acc;
            }
        }
        __syncwarp();
        const int deg = deg_s[warp];
        // Chien search: lane tests i = lane+1, lane+33, ... ; roots kept in a per-lane bit mask
        // lambda(alpha^i) for this lane's eight positions at once, coefficient by coefficient: the exponent of term j moves
        // by 32 j from one position to the next, so it is stepped (one add, one conditional subtract) instead of being
        // formed with a multiply and a division by 255 per term and position -- the same field elements, summed in GF(256)
        unsigned mine = 0, count = 0;
        {
            unsigned q8[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) q8[it] = 1u;
            for (int j = 1; j <= deg; ++j) {
                const unsigned c = L[j];
                if (!c) continue;                    // (uniform: L is the warp's)
                unsigned e = (glog[c] + (unsigned)(lane + 1) * (unsigned)j) % 255u;
                const unsigned step = (32u * (unsigned)j) % 255u;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    q8[it] ^= gexp[e];
                    e += step;
                    e -= (e >= 255u) ? 255u : 0u;
                }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const unsigned i = (unsigned)it * 32u + (unsigned)lane + 1u;
                const bool root = (i <= 255u) && (q8[it] == 0u);
                if (root) mine |= 1u << it;
                count += __popc(__ballot_sync(0xffffffffu, root));
            }
        }
        if ((int)count == deg) {
            for (unsigned it = 0; it < 8; ++it) {
                if (!((mine >> it) & 1u)) continue;
                const unsigned root = it * 32 + lane + 1, loc = root - 1;
                unsigned num = 0, den = 0;
                const unsigned rr = root % 255u, rr2 = (2u * rr) % 255u;         // exponent steps of root^i, root^(2 i)
                unsigned ex = 0;
                for (int i = 0; i < deg; ++i) {
                    if (O[i]) { unsigned e = glog[O[i]] + ex; e -= (e >= 255u) ? 255u : 0u; num ^= gexp[e]; }
                    ex += rr; ex -= (ex >= 255u) ? 255u : 0u;
                }
                const int top = (deg < 31 ? deg : 31) & ~1;
                ex = 0;
                for (int i = 0; i <= top; i += 2) {
                    if (L[i + 1]) { unsigned e = glog[L[i + 1]] + ex; e -= (e >= 255u) ? 255u : 0u; den ^= gexp[e]; }
                    ex += rr2; ex -= (ex >= 255u) ? 255u : 0u;
                }
                if (num != 0 && loc >= pad) x[loc - pad] ^= gexp[(glog[num] + 255 - glog[den]) % 255];
            }
        }
        __syncwarp();
    }
    const unsigned n0 = blk * dec_block;
    const unsigned take = (n - n0 >= dec_block) ? dec_block : (n - n0);
    unsigned char *dst = io.dst + n0;
    for (unsigned i = lane; i < take; i += 32) dst[i] = x[i];
    }
}

// ------------------------------------------------------------------ unscramble + CRC + copy-out (warp per frame)
__global__ void __launch_bounds__(128)
k_crc(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list)
{
    // slicing-by-4 tables of the check the CTA's first frame uses (frames of one call nearly always share it; a warp
    // whose frame uses another check falls back to the byte-at-a-time loop on the global table)
    __shared__ unsigned crc4[4][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned cta_check = P.frames[list[min(blockIdx.x * 4u, n_list - 1u)]].check;
    if (cta_check >= 3 && cta_check <= 6) {
        const unsigned *tab = P.tables->crc_tab[cta_check];
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            const unsigned t0 = tab[i];
            const unsigned t1 = (t0 >> 8) ^ tab[t0 & 0xffu], t2 = (t1 >> 8) ^ tab[t1 & 0xffu], t3 = (t2 >> 8) ^ tab[t2 & 0xffu];
            crc4[0][i] = t0; crc4[1][i] = t1; crc4[2][i] = t2; crc4[3][i] = t3;
        }
    }
    __syncthreads();
    const unsigned gi = blockIdx.x * 4 + warp;
    if (gi >= n_list) return;
    FrameDesc &d = P.frames[list[gi]];
    unsigned char *buf = P.bufA + d.buf_off;        // k0 bytes: payload + crc, still whitened
    unsigned char *out = P.payload + d.pay_off;
    const unsigned plen = d.payload_len, k0 = d.k0, cl = k0 - plen;
    for (unsigned i = lane; i < k0; i += 32) {
        unsigned mask = (i & 3u) == 0 ? 0xb4u : (i & 3u) == 1 ? 0x6au : (i & 3u) == 2 ? 0x8bu : 0xc5u;
        unsigned v = buf[i] ^ mask;
        buf[i] = (unsigned char)v;
        if (i < plen) out[i] = (unsigned char)v;
    }
    __syncwarp();
    if (lane == 0) {
        unsigned key = 0, rx = 0;
        for (unsigned i = 0; i < cl; ++i) rx = (rx << 8) | buf[plen + i];
        if (d.check == 2) {
            unsigned sum = 0;
            for (unsigned i = 0; i < plen; ++i) sum += buf[i];
            key = (~sum + 1u) & 0xffu;
        } else if (d.check >= 3 && d.check <= 6) {
            const unsigned *tab = P.tables->crc_tab[d.check];
            unsigned k = 0xffffffffu;
            unsigned i0 = 0;
            if (d.check == cta_check) {
                // four bytes per dependent step (buf is 16-byte aligned: buf_off is a multiple of 16)
                const unsigned *b4 = reinterpret_cast<const unsigned *>(buf);
                const unsigned nw = plen >> 2;
                for (unsigned i = 0; i < nw; ++i) {
                    k ^= b4[i];
                    k = crc4[3][k & 0xffu] ^ crc4[2][(k >> 8) & 0xffu] ^ crc4[1][(k >> 16) & 0xffu] ^ crc4[0][k >> 24];
                }
                i0 = 4 * nw;
            }
            for (unsigned i = i0; i < plen; ++i) k = (k >> 8) ^ tab[(k ^ buf[i]) & 0xffu];
            const unsigned bits = d.check == 3 ? 8u : d.check == 4 ? 16u : d.check == 5 ? 24u : 32u;
            key = (~k) & (bits == 32 ? 0xffffffffu : ((1u << bits) - 1u));
        }
        d.payload_valid = (key == rx) ? 1 : 0;
    }
}

}  // namespace

void launch_deinterleave(const PayloadParams &P, const unsigned *list, unsigned n, int stage, cudaStream_t s)
{
    if (n) k_deinterleave<<<n, 256, 0, s>>>(P, list, stage);
}
void launch_blockfec(const PayloadParams &P, const unsigned *list, unsigned n, int stage, cudaStream_t s)
{
    if (n) k_blockfec<<<n, 256, 0, s>>>(P, list, stage);
}
void launch_viterbi(const PayloadParams &P, const unsigned *list, unsigned n, int stage, unsigned K, bool punct, cudaStream_t s)
{
    if (!n) return;
    if (K == 7) {
        // all frames of one launch share the stage's scheme class only loosely: punctured and plain rate-1/2 codes may be
        // mixed in one list, so the generic symbol fetch is used unless the caller's list is known to be plain (punct == 0)
        const unsigned threads = n * kV4Lanes;
        // test hook: a short warm-up makes the speculative traceback start wrong often, which exercises the re-walk
        const char *we = std::getenv("LQB_V4_WARM");
        const int warm = we ? std::max(0, std::atoi(we)) : kV4Warm;
        if (punct) k_viterbi27x4<true><<<(threads + kV4Threads - 1) / kV4Threads, kV4Threads, 0, s>>>(P, list, n, stage, reinterpret_cast<unsigned *>(P.decisions), warm);
        else k_viterbi27x4<false><<<(threads + kV4Threads - 1) / kV4Threads, kV4Threads, 0, s>>>(P, list, n, stage, reinterpret_cast<unsigned *>(P.decisions), warm);
    }
    else if (!std::getenv("LQB_V9_GENERIC")) k_viterbi29x16<false><<<(n * kV9Lanes + kV9Threads - 1) / kV9Threads, kV9Threads, 0, s>>>(P, list, n, stage);
    else k_viterbi<false><<<(n + kVitWarps - 1) / kVitWarps, 32 * kVitWarps, 0, s>>>(P, list, n, stage);      // (the warp-per-codeword form, kept for A/B)
}
void launch_viterbi_soft(const PayloadParams &P, const unsigned *list, unsigned n, int stage, unsigned K, bool punct, cudaStream_t s)
{
    if (!n) return;
    if (K == 7) {
        const unsigned threads = n * kV4Lanes;
        const char *we = std::getenv("LQB_V4_WARM");
        const int warm = we ? std::max(0, std::atoi(we)) : kV4Warm;
        if (punct) k_viterbi27x4<true, true><<<(threads + kV4Threads - 1) / kV4Threads, kV4Threads, 0, s>>>(P, list, n, stage, reinterpret_cast<unsigned *>(P.decisions), warm);
        else k_viterbi27x4<false, true><<<(threads + kV4Threads - 1) / kV4Threads, kV4Threads, 0, s>>>(P, list, n, stage, reinterpret_cast<unsigned *>(P.decisions), warm);
    }
    else if (!std::getenv("LQB_V9_GENERIC")) k_viterbi29x16<true><<<(n * kV9Lanes + kV9Threads - 1) / kV9Threads, kV9Threads, 0, s>>>(P, list, n, stage);
    else k_viterbi<true><<<(n + kVitWarps - 1) / kVitWarps, 32 * kVitWarps, 0, s>>>(P, list, n, stage);
}
void launch_rs(const PayloadParams &P, const unsigned *blocks, unsigned n_blocks, int stage, cudaStream_t s)
{
    if (!n_blocks) return;
    const int sms = sm_count_of_this_device();
    const unsigned want = (n_blocks + kRsWarps - 1) / kRsWarps, cap = (unsigned)sms * 6u;     // six 34 KB CTAs per SM
    k_rs<<<std::min(want, cap), 32 * kRsWarps, kRsSynBytes, s>>>(P, blocks, n_blocks, stage);
}
void launch_crc(const PayloadParams &P, const unsigned *list, unsigned n, cudaStream_t s)
{
    if (n) k_crc<<<(n + 3) / 4, 128, 0, s>>>(P, list, n);
}

}  // namespace lqb
