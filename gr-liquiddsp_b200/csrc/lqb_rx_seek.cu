// lqb_rx_seek.cu -- frame discovery: preamble search (qdetector seek), alignment
// (tau/gamma/dphi/phi estimation), header matched filter + pilot sync + header decode.
//
// Replaces, for a batch of independent streams, the part of liquid-dsp's flexframesync
// state machine that decides WHERE frames are (reference call site
// lib/flex_rx_impl.cc:213 -> flexframesync_execute; lib/frame_detector_cc_impl.cc:77 ->
// qdetector_cccf_execute).  One CTA walks one stream's hop grid (a 512-sample window every 256
// samples).  Every window first goes through a tensor-core PRE-FILTER: the correlation of its 356
// linear lags with the template at all 49 CFO bins as one dense contraction (tcgen05.mma,
// kind::f8f6f4: e4m3 operands, fp32 accumulators in TMEM, the Hankel operand never materialised),
// with a RIGOROUS bound on what the e4m3 rounding can have changed.  A window the bound proves
// unable to trigger is skipped; only the rest (the real detections plus a fraction of a percent)
// take qdetector's exact evaluation (1 forward FFT + 49 frequency-shifted inverse FFTs, one warp
// per FFT, 16 points per lane), so decisions are those of the specification.  On a hit the CTA runs
// the serial alignment / header steps and moves its hop grid past the frame.  Four worker warps
// stage samples and read accumulators, a fifth only issues MMAs; three CTAs share an SM.
// The payload itself is left to the frame-parallel kernels in lqb_rx_payload.cu.
#include "lqb_dev.cuh"
#include "lqb_kernels.h"
#include "lqb_tc.cuh"
#include <cmath>

namespace lqb {

namespace {

#ifndef LQB_SEEK_REORDER
#define LQB_SEEK_REORDER 0               // stage A of block k+2 between the two tile read-outs of block k (0: after the decision; measured A/B on B200: 21.6 ms without, 22.8 ms with -- shared memory is the busy resource, not the wait)
#endif
#ifndef LQB_SEEK_BINS
#define LQB_SEEK_BINS 1                  // exact windows visit only the CFO bins the tensor cores cannot rule out (0: all 49)
#endif
#ifndef LQB_SEEK_THR_PRUNE
#define LQB_SEEK_THR_PRUNE 1             // the bin scan also drops bins that cannot reach the threshold; every window that is not skipped takes it
#endif
#ifndef LQB_SEEK_SLEEP
#define LQB_SEEK_SLEEP 0                 // scale of the nanosleep back-off in mbarrier waits (0: plain polling)
#endif
constexpr int kWarps = 4;                 // worker warps (TMEM lane quarters: warp w reads accumulator rows 32w .. 32w+31)
constexpr int kThreads = 32 * kWarps;     // worker threads
constexpr int kCtaThreads = kThreads + 32; // + one warp that only issues tcgen05.mma
constexpr float kPiF = 3.14159274f;     // (float)M_PI

#ifdef LQB_SEEK_PROF
// per-phase cycle accounting (thread 0 of every CTA), debug builds only: see lqb_dbg_seek_prof()
__device__ unsigned long long g_seek_prof[24];
#define PROF_DECL long long prof_last = clock64(); long long prof_acc[24] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }
#define PROF_MARK(i) do { if (threadIdx.x == 0) { long long now_ = clock64(); prof_acc[i] += now_ - prof_last; prof_last = now_; } } while (0)
#define PROF_ARGS , long long &prof_last, long long (&prof_acc)[24]
#define PROF_PASS , prof_last, prof_acc
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_ARGS
#define PROF_PASS
#endif

#ifdef LQB_SEEK_TRACE
// one record per CTA (debug builds only): start / end in ns (globaltimer), SM, stream.  See lqb_dbg_seek_trace().
__device__ unsigned long long g_seek_trace[4][1 << 16];
__device__ unsigned g_seek_trace_n;
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid_of() { unsigned v; asm volatile("mov.u32 %0, %smid;" : "=r"(v)); return v; }
#endif

// acquire / release accesses for the slice queue (a stream's state passes from one CTA to the next inside one launch)
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
constexpr unsigned kQueueEmpty = 0xffffffffu;
// a draw that waits longer than this many polls (200 ns apart: about two seconds) gives up: the CTA leaves, the host
// reports the stall (lqb_dbg_seek_stall) instead of the GPU hanging
constexpr unsigned kQueueMaxPolls = 10u * 1000u * 1000u;
__device__ unsigned g_seek_stall[8];

// barrier over the worker warps only (the MMA warp never joins it)
__device__ __forceinline__ void wsync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// fused tensor-core pre-filter geometry: a block correlates n_t * 128 consecutive lags
constexpr int kTcRows = 2 * 128 + 160;          // 416 rows of 16 bytes per component
constexpr int kTcZBytes = kTcRows * 16;         // 6656

struct SeekShared {
    // The exact-evaluation buffers and the second Z strip never live at the same time (the pre-filter
    // pipeline is drained before a window is evaluated exactly), so they share storage; Sc and W are
    // reloaded from the constant tables when the strip has overwritten them (tables_dirty).
    union {
        struct {
            float2 Xw[512];           // time-domain window / aligned buffer
            float2 Xf[512];           // its spectrum (also reused for the CFO spectrum)
            float2 Sc[512];           // conj(S)
            float2 W[256];            // twiddles
        };
        unsigned char Z1[2 * kTcZBytes];
    };
    float2 Wc[240];           // per-stage compact twiddles (stages 5-8)
    float2 scr[2176];         // per-warp FFT transpose scratch (4 x 544); flat header-stage scratch; Z strip 0 + staged halfs
    unsigned long long best[kWarps];
    float  energy[2];
    float2 y3[3];             // align: y[511], y[0], y[1]
    float  cfo_nb[2];         // align: |CFO spectrum| left and right of its peak
    // control (written by thread 0)
    int    trig, idx, off, stop, hv;
    float  rxy, tau, gamma, dphi, phi, mf_scale;
    unsigned theta0, dtheta, pfb, tau_neg;
    // fused tensor-core pre-filter pipeline
    uint64_t z_full[2];       // workers -> MMA warp: strip b is staged (count = 128)
    uint64_t acc_full;        // tcgen05.commit -> workers: accumulators hold the block
    uint64_t acc_empty;       // workers -> MMA warp: accumulators have been read out (count = 128)
    unsigned tmem_base;
    int    tc_cmd[2];         // tiles in strip b (0 = the MMA warp exits)
    int    tables_dirty;
    float  part_max[4][kWarps], part_en[4][kWarps];   // block statistics, four slots in rotation (three blocks are alive at a time)
    float  chunk_e[16], chunk_d[16];  // per 32 staged samples: energy, e4m3 residual energy (scaled units)
    float  span_e[4], span_d[4];      // per block: max over any 6 consecutive chunks (covers every lag's 156-sample span)
    float  red[kWarps][4];
    float  tc_red[2 * kWarps];
    unsigned binmax[tc::kNBins + 7];  // bin_candidates: max over all 512 circular lags of |C_q|^2 per CFO bin (float bits)
    unsigned long long cand;
    unsigned char hbytes[64]; // header: 54 demodulated bytes
    unsigned char hdec[32];   // decoded header (24 bytes incl. CRC)
};

// flat offsets inside scr for the align/header stage
constexpr int kVbuf = 0;      // 640: derotated samples n = 128 .. 616
constexpr int kHsym = 640;    // 232: header symbols
constexpr int kZbuf = 1088;   // 512: x * conj(s) for CFO estimation
constexpr int kVsum = 1600;   // 160: derotated products for the phase estimate
constexpr int kPil  = 1760;   // 64 : pilot FFT in/out

__device__ __forceinline__ void load_window(SeekShared &sh, const StreamView &sv, long long start, int tid)
{
#pragma unroll
    for (int k = 0; k < 512 / kThreads; ++k) {
        int i = tid + kThreads * k;
        sh.Xw[i] = sv.at(start + i);
    }
}

// energy of 256 samples as a balanced pairwise tree in index order
__device__ __forceinline__ float half_energy(const float2 *x, int lane)
{
    float e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = abs2f(x[lane * 8 + q]);
    float a = __fadd_rn(__fadd_rn(__fadd_rn(e[0], e[1]), __fadd_rn(e[2], e[3])),
                        __fadd_rn(__fadd_rn(e[4], e[5]), __fadd_rn(e[6], e[7])));
#pragma unroll
    for (int m = 1; m <= 16; m <<= 1) a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, m));
    return a;
}

__device__ __forceinline__ void forward_fft_to(SeekShared &sh, const float2 *src, float2 *dst, int lane)
{
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = src[fft512_in_index(lane, r)];
    fft512_warp<+1>(v, sh.W, sh.Wc, sh.scr, lane);      // warp 0's scratch
#pragma unroll
    for (int r = 0; r < 16; ++r) dst[fft512_out_index(lane, r)] = v[r];
}

// inverse FFT of Xf .* conj(S shifted by off); leaves y in registers
__device__ __forceinline__ void cross_ifft(SeekShared &sh, int off, float2 (&v)[16], float2 *scratch, int lane)
{
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        int i = fft512_in_index(lane, r);
        v[r] = cmulf(sh.Xf[i], sh.Sc[(i - off) & 511]);
    }
    fft512_warp<-1>(v, sh.W, sh.Wc, scratch, lane);
}

// evaluate the window in sh.Xw; thread 0 publishes trig/idx/off/rxy.
// cand: CFO bins (bit offi) that can hold the global maximum; the others are proven smaller (bin_candidates) and skipped.
__device__ void eval_window(SeekShared &sh, const DevTables *T, int tid, unsigned long long cand PROF_ARGS)
{
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 2) {
        float e = half_energy(sh.Xw + warp * 256, lane);
        if (lane == 0) sh.energy[warp] = e;
    }
    if (warp == 0) forward_fft_to(sh, sh.Xw, sh.Xf, lane);
    wsync();
    PROF_MARK(15);
    const float g0 = __fmul_rn(__fsqrt_rn(__fadd_rn(sh.energy[0], sh.energy[1])), __fsqrt_rn(156.0f / 512.0f));
    unsigned long long best = 0ull;
    if (g0 >= 1e-10f) {
        const int range = T->range;
        float2 v[16];
        // per-lane running maximum; strict '>' keeps the earliest (offset, lag) on ties within the lane
        float bv = 0.0f;
        unsigned border = 0;
        // the candidate bins are dealt round-robin to the warps, each warp taking its bins in increasing order
        int turn = 0;
        for (int offi = 0; offi <= 2 * range; ++offi) {
            if (!((cand >> offi) & 1ull)) continue;
            if ((turn++ & (kWarps - 1)) != warp) continue;
            cross_ifft(sh, offi - range, v, sh.scr + warp * 544, lane);
            const unsigned obase = (unsigned)(offi * 512 + fft512_out_index(lane, 0));
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float a2 = abs2f(v[r]);
                const bool gt = a2 > bv;
                bv = gt ? a2 : bv;
                border = gt ? obase + 16u * r : border;
            }
        }
        best = ((unsigned long long)__float_as_uint(bv) << 32) | (0xffffffffu - border);
        best = warp_max_u64(best);
    }
    if (lane == 0) sh.best[warp] = best;
    wsync();
    PROF_MARK(16);
    if (tid == 0) {
        unsigned long long b = sh.best[0];
        for (int w = 1; w < kWarps; ++w) b = sh.best[w] > b ? sh.best[w] : b;
        int trig = 0, idx = 0, off = 0;
        float rxy = 0.0f;
        if (g0 >= 1e-10f) {
            float peak2 = __uint_as_float((unsigned)(b >> 32));
            unsigned order = 0xffffffffu - (unsigned)(b & 0xffffffffu);
            if (peak2 > 0.0f) { idx = (int)(order & 511u); off = (int)(order >> 9) - T->range; }
            float g = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(512.0f, g0), __fsqrt_rn(T->s2_sum)));
            rxy = __fmul_rn(__fsqrt_rn(peak2), g);
            trig = (rxy > T->threshold) && (idx < 512 - 156);
        }
        sh.trig = trig; sh.idx = idx; sh.off = off; sh.rxy = rxy;
    }
    wsync();
}


// ------------------------------------------------------------------ fused tensor-core pre-filter (pipelined)
// A block correlates n_t * 128 consecutive lags starting at absolute sample a0 against the template at
// all 49 CFO bins on the tensor cores (e4m3 operands, fp32 accumulation in TMEM).  Work is split three ways
// so that the tensor pipe never waits for the CUDA cores of its own CTA:
//   tc_stage_a / tc_stage_b (workers) : quantise the samples of block k+2 to e4m3 (A), build its Z strip and hand it to
//                                       the MMA warp (B)
//   mma_warp  (warp 4)                : per 128-lag tile: wait for the (single) accumulator to be free, issue 10 MMAs, commit
//   tc_retire_tile (workers)          : per tile: read the accumulator (max_b |C|^2 per lag), release it
// Strip layout: Z[c][16 m + e] = e4m3(x_c[a0 + m + e] * 2^(4 - ex)), two component planes, 416 rows of 16 bytes:
// row m IS the K-major core-matrix row "16 consecutive samples from a0 + m", so a no-swizzle UMMA descriptor with
// row-group stride 128 B and K-chunk stride 256 B reads the Hankel matrix A[lag][k] = x[a0 + lag + k] in place.
//
// Error bound (why skipping on the fp8 result is exact).  With x_q, t_q the quantised samples / template,
//   |C[l,b] - C_q[l,b]| <= ||x_l - x_q,l|| ||t_b|| + ||x_q,l|| ||t_b - t_q,b||        (Cauchy-Schwarz over the 156-span)
// The rounding residual v 2^k - e4m3(v 2^k) is exact in fp32, so its energy is ACCUMULATED EXACTLY while quantising,
// per chunk of 32 samples; a lag's 156-sample span touches at most six consecutive chunks, so
//   ||x_l - x_q,l||^2 <= D6 = max over six consecutive chunks of the residual energy, ||x_l||^2 <= E6 likewise,
// ||t_b|| = ||s||, and max_b ||t_b - t_q,b|| / ||s|| = P.b_err is computed on the host from the very bytes of B.
// With g0 = sqrt(E_window 156 / 512) (qdetector's normaliser):
//   rxy <= rxy_q + (sqrt(D6) + (sqrt(E6) + sqrt(D6)) b_err) / g0
// (typically 0.06; measured differences are several times smaller).  fp32 accumulation of 320 exact products adds
// < 1e-4; the decision keeps a further 0.008 in hand.  Saturated blocks (scaled maximum >= 384) are not trusted.
struct TcBlk {
    long long a0, e_lo, e_hi, w, next_a0;   // first lag; energy range; window this block belongs to; first lag of the block after it
    int n_t, cold, ex, buf, slot;           // tiles; cold-start block of a hop grid; scale exponent; strip buffer; statistics slot
    long long wrap_w;                       // >= LLONG_MIN + 1: samples are taken circularly from the 512-sample window at wrap_w
};
constexpr long long kNoWrap = -(1ll << 62);
struct TcRes { float m0_all, m0_ge28, m1_all, m1_ge28, mx, en, span_e, span_d; };

constexpr int kXsWords = 112;       // staged e4m3 samples: words per component plane (448 bytes >= 416 + 15 rows)

__device__ __forceinline__ unsigned char *tc_strip(SeekShared &sh, int buf)
{
    return buf ? sh.Z1 : reinterpret_cast<unsigned char *>(sh.scr);
}
__device__ __forceinline__ uint32_t *tc_xs(SeekShared &sh)
{
    return reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(sh.scr) + 2 * kTcZBytes);   // [2][kXsWords], behind strip 0
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tc::smem_u32(bar)) : "memory");
}

// Stage A of a block: quantise its samples to e4m3 into the staging planes `xs` and take its statistics (maximum, window
// energy, per-span energy and residual maxima) into statistics slot b.slot.  Touches neither strip, so it can run while
// the tensor core still reads the strip this block will later be built into.
// pre[] / pre_a0: samples already fetched for the block starting at pre_a0 (by the previous stage B).
// Thread t owns the four consecutive samples 4 t .. 4 t + 3 of the block (threads 106.. hold zeros).
__device__ void tc_stage_a(SeekShared &sh, const StreamView &sv, const TcBlk &b, int tid, float2 (&pre)[4], long long pre_a0 PROF_ARGS)
{
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t *xs = tc_xs(sh);
    const int n_samp = 128 * b.n_t + 168;
    const float sc = __uint_as_float((uint32_t)(127 + 4 - b.ex) << 23);      // exact power of two: block maximum -> [16, 32)
    wsync();                                   // the previous strip build has finished reading xs
    float mx = 0.0f, en = 0.0f, dd = 0.0f, ee = 0.0f;
    if (b.wrap_w != kNoWrap) {                 // (uniform) circular lags of one window: sample index modulo 512
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * tid + k;
            pre[k] = (i < 424) ? sv.at(b.wrap_w + (((b.a0 - b.wrap_w) + i) & 511)) : make_float2(0.0f, 0.0f);
        }
    } else if (pre_a0 != b.a0) {               // (uniform) nothing usable was prefetched: fetch now
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * tid + k;
            pre[k] = (i < 424) ? sv.at(b.a0 + i) : make_float2(0.0f, 0.0f);
        }
    }
    uint32_t pk[2] = { 0u, 0u };               // (re0 im0 re1 im1), (re2 im2 re3 im3) as e4m3 bytes
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * tid + k;
        const long long n = b.a0 + i;
        float2 v = pre[k];
        if (i >= n_samp) v = make_float2(0.0f, 0.0f);
        mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
        const float a2 = fmaf(v.y, v.y, v.x * v.x);
        if (i < n_samp && n >= b.e_lo && n < b.e_hi) en += a2;
        ee += a2;
        const float2 vs = make_float2(v.x * sc, v.y * sc);
        const __nv_fp8x2_storage_t q = __nv_cvt_float2_to_fp8x2(vs, __NV_SATFINITE, __NV_E4M3);      // low byte: re
        const __half2_raw hb = __nv_cvt_fp8x2_to_halfraw2(q, __NV_E4M3);
        const float2 back = __half22float2(*reinterpret_cast<const __half2 *>(&hb));
        const float rx = vs.x - back.x, ry = vs.y - back.y;                  // exact: what the tensor core will not see
        dd += fmaf(ry, ry, rx * rx);
        pk[k >> 1] |= (uint32_t)q << (16 * (k & 1));
    }
    if (tid < kXsWords) {
        xs[tid] = __byte_perm(pk[0], pk[1], 0x6420);                         // re plane: bytes 0, 2 of each word
        xs[kXsWords + tid] = __byte_perm(pk[0], pk[1], 0x7531);              // im plane
    }
    // non-negative floats order like their bit patterns: one REDUX instead of five shuffle + max rounds
    mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(mx)));
    // chunk sums: eight consecutive lanes hold one chunk of 32 samples
#pragma unroll
    for (int m = 1; m <= 4; m <<= 1) { ee += __shfl_xor_sync(0xffffffffu, ee, m); dd += __shfl_xor_sync(0xffffffffu, dd, m); }
    if ((lane & 7) == 0) { sh.chunk_e[tid >> 3] = ee; sh.chunk_d[tid >> 3] = dd; }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) en += __shfl_xor_sync(0xffffffffu, en, m);
    if (lane == 0) { sh.part_max[b.slot][warp] = mx; sh.part_en[b.slot][warp] = en; }
    wsync();
    if (warp == 0) {
        // lag p (0 .. 255) spans samples p .. p + 155 = chunks p / 32 .. (p + 155) / 32: six consecutive chunks from j = 0 .. 7
        float se = 0.0f, sd = 0.0f;
        if (lane < 9) {
#pragma unroll
            for (int q = 0; q < 6; ++q) { if (lane + q < 16) { se += sh.chunk_e[lane + q]; sd += sh.chunk_d[lane + q]; } }
        }
        se = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(se)));
        sd = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(sd)));
        if (lane == 0) { sh.span_e[b.slot] = se; sh.span_d[b.slot] = sd; }
    }
    PROF_MARK(1);
}

// Stage B: build the block's Z strip from the staging planes, hand it to the MMA warp, request the samples of the block
// that will be staged next (they travel while this one is multiplied).
__device__ void tc_stage_b(SeekShared &sh, const StreamView &sv, const TcBlk &b, int tid, float2 (&pre)[4], long long &pre_a0 PROF_ARGS)
{
    using namespace tc;
    unsigned char *Z = tc_strip(sh, b.buf);
    const uint32_t *xs = tc_xs(sh);
    // ---- Z[c][16 m + e] = xs[c][m + e]: row m is the 16 bytes at byte offset m of the plane
    const int rows = 128 * b.n_t + 160;
    for (int m = tid; m < rows; m += kThreads) {
        const int w0 = m >> 2;
        const unsigned sh8 = 8u * (unsigned)(m & 3);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t *s = xs + kXsWords * c + w0;
            const uint32_t q0 = s[0], q1 = s[1], q2 = s[2], q3 = s[3], q4 = s[4];
            uint4 o;
            o.x = __funnelshift_r(q0, q1, sh8); o.y = __funnelshift_r(q1, q2, sh8);
            o.z = __funnelshift_r(q2, q3, sh8); o.w = __funnelshift_r(q3, q4, sh8);
            *reinterpret_cast<uint4 *>(Z + c * kTcZBytes + 16 * m) = o;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) sh.tc_cmd[b.buf] = b.n_t;
    mbar_arrive(&sh.z_full[b.buf]);
    PROF_MARK(2);
    // a block that lies wholly inside the new input (the usual case) is read straight through the pointer
    if (b.next_a0 == kNoWrap) pre_a0 = kNoWrap;        // (nothing follows: bin_candidates)
    else {
        const long long next_a0 = b.next_a0;
        const long long i0 = next_a0 - sv.base - (long long)sv.carry_len;
        if (i0 >= 0 && next_a0 >= sv.G && next_a0 + 424 <= sv.end) {
            const float2 *src = sv.in + i0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * tid + k;
                pre[k] = (i < 424) ? __ldg(src + i) : make_float2(0.0f, 0.0f);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * tid + k;
                pre[k] = (i < 424) ? sv.at(next_a0 + i) : make_float2(0.0f, 0.0f);
            }
        }
        pre_a0 = next_a0;
    }
    PROF_MARK(3);
}

// the MMA warp: one lane issues, the warp only ever waits on mbarriers.  One accumulator tile (128 lags x 112
// columns, 128 TMEM columns per CTA so that three CTAs fit an SM): a tile's MMAs start when the workers have read
// the previous tile out; the tensor pipe is kept busy by the other CTAs of the SM meanwhile.
__device__ void mma_warp(SeekShared &sh, const unsigned char *Bsm, int lane)
{
    using namespace tc;
    unsigned pz[2] = { 0u, 0u }, pe = 1u;      // parity 1 on a fresh barrier: "the accumulator starts out free"
    int buf = 0;
    const uint32_t tmem = sh.tmem_base;
    // instruction descriptor: D = f32 (bit 4), A = B = e4m3 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b_addr = smem_u32(Bsm);
    while (true) {
        mbar_wait(&sh.z_full[buf], pz[buf], 200u * LQB_SEEK_SLEEP);
        pz[buf] ^= 1u;
        const int n_t = *reinterpret_cast<volatile int *>(&sh.tc_cmd[buf]);
        if (n_t == 0) break;
        const unsigned char *Z = tc_strip(sh, buf);
        for (int t = 0; t < n_t; ++t) {
            mbar_wait(&sh.acc_empty, pe, 100u * LQB_SEEK_SLEEP);
            pe ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                uint32_t acc = 0;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t za = smem_u32(Z + c * kTcZBytes) + 2048u * t;        // 128 lags further: 128 rows
#pragma unroll
                    for (int j = 0; j < kMmaPerComp; ++j) {
                        // A: K-chunk (16 samples) stride 256 B, row-group (8 lags) stride 128 B; 32 samples per MMA = 512 B
                        mma_f8(tmem, make_desc(za + 512u * j, 256u, 128u),
                               make_desc(b_addr + (uint32_t)((c * kMmaPerComp + j) * 2) * kBChunkBytes, kBChunkBytes, 128u), idesc, acc);
                        acc = 1;
                    }
                }
                mma_commit(&sh.acc_full);
            }
            __syncwarp();
        }
        buf ^= 1;
    }
}

// wait for the next accumulator tile, take max_b |C|^2 of this thread's lag, release the accumulator.
// Columns: re of bin b at column b, im at column 56 + b (b < 49; the padding columns of B are zero).
__device__ float tc_retire_tile(SeekShared &sh, unsigned &ph_full, int tid, bool discard PROF_ARGS)
{
    using namespace tc;
    PROF_MARK(0);
    mbar_wait(&sh.acc_full, ph_full, 40u * LQB_SEEK_SLEEP);
    ph_full ^= 1u;
    PROF_MARK(4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float best = 0.0f;
    if (!discard) {
        const uint32_t taddr = sh.tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
        float bt0 = 0.0f, bt1 = 0.0f, bt2 = 0.0f, bt3 = 0.0f;       // independent running maxima: no single dependent chain
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            uint32_t re[2][8], im[2][8];
            const int nc = (round == 3) ? 1 : 2;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u < nc) { tmem_ld8(taddr + 8u * (2 * round + u), re[u]); tmem_ld8(taddr + 56u + 8u * (2 * round + u), im[u]); }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u >= nc) continue;
                float2 a[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 r2 = make_float2(__uint_as_float(re[u][2 * k]), __uint_as_float(re[u][2 * k + 1]));
                    const float2 i2 = make_float2(__uint_as_float(im[u][2 * k]), __uint_as_float(im[u][2 * k + 1]));
                    a[k] = __ffma2_rn(i2, i2, __fmul2_rn(r2, r2));           // two bins per instruction (FMUL2 / FFMA2)
                }
                bt0 = max3f(bt0, a[0].x, a[0].y); bt1 = max3f(bt1, a[1].x, a[1].y);
                bt2 = max3f(bt2, a[2].x, a[2].y); bt3 = max3f(bt3, a[3].x, a[3].y);
            }
        }
        best = fmaxf(fmaxf(bt0, bt1), fmaxf(bt2, bt3));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&sh.acc_empty);
    PROF_MARK(5);
    return best;
}

// reduce the per-lag maxima of a block's tiles to the four range maxima the decisions need, and fetch its statistics
__device__ void tc_retire_reduce(SeekShared &sh, const TcBlk &b, float best0, float best1, int tid, TcRes &r PROF_ARGS)
{
    const int warp = tid >> 5, lane = tid & 31;
    // |C|^2 >= 0: the maxima are taken on the bit patterns with one REDUX each
    const float a = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best0)));
    const float bb = __uint_as_float(__reduce_max_sync(0xffffffffu, tid >= 28 ? __float_as_uint(best0) : 0u));
    const float c = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best1)));
    const float d = __uint_as_float(__reduce_max_sync(0xffffffffu, tid >= 28 ? __float_as_uint(best1) : 0u));
    if (lane == 0) { sh.red[warp][0] = a; sh.red[warp][1] = bb; sh.red[warp][2] = c; sh.red[warp][3] = d; }
    wsync();
    r.m0_all = 0.0f; r.m0_ge28 = 0.0f; r.m1_all = 0.0f; r.m1_ge28 = 0.0f; r.mx = 0.0f; r.en = 0.0f;
    r.span_e = sh.span_e[b.slot]; r.span_d = sh.span_d[b.slot];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        r.m0_all = fmaxf(r.m0_all, sh.red[w][0]); r.m0_ge28 = fmaxf(r.m0_ge28, sh.red[w][1]);
        r.m1_all = fmaxf(r.m1_all, sh.red[w][2]); r.m1_ge28 = fmaxf(r.m1_ge28, sh.red[w][3]);
        r.mx = fmaxf(r.mx, sh.part_max[b.slot][w]); r.en += sh.part_en[b.slot][w];
    }
    wsync();
    PROF_MARK(6);
}

// carry between consecutive windows of one hop grid (uniform across the workers)
struct ScanCarry {
    bool valid, unsafe, ex_valid;
    long long w;            // window the carried tail / half-energy belong to
    float tail, half;       // max |C|^2 over its first 100 lags; energy of its first 256 samples
    float tail_d, tail_e;   // of the block that produced the tail: largest residual energy / energy of a 156-sample span
    int ex;                 // scale exponent for the next strips
};

__device__ __forceinline__ float fast_sqrt(float x)
{
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// the scale is only a choice (any power of two keeps the bound exact); clamped so that the factors below stay normal
__device__ __forceinline__ int exponent_of(float mx)
{
    int ex = 0;
    if (mx > 0.0f) ex = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;
    return max(-40, min(40, ex));
}

// Walk the hop grid from st.wstart with the pre-filter while windows can be ruled out.
// Returns 1 when window st.wstart (updated) needs the exact evaluation, 0 when the input is exhausted.
// The pipeline is empty on return.
//
// Order of work per window (block k is being decided, k+1 is staged, k+2 is new), chosen so that the workers never idle
// while the tensor core multiplies -- a CTA has ONE accumulator tile, so tile 1 of block k can only be multiplied once
// tile 0 has been read out:
//   read tile 0 of k | quantise k+2 (stage A: no strip touched) while tile 1 of k is multiplied | read tile 1 of k |
//   decide k | build the strip of k+2 in the buffer k used (stage B) -- tile 0 of k+1 is multiplied meanwhile.
__device__ int fused_scan(SeekShared &sh, const StreamView &sv, const DevTables *T, float b_err, StreamState &st, int tid,
                          ScanCarry &c, float2 (&pre)[4], long long &pre_a0, unsigned &ph_full, int &buf_w,
                          unsigned &n_windows, unsigned &n_tiles, float &rxy_q PROF_ARGS)
{
    rxy_q = 0.0f;                                      // pre-filter's estimate of rxy for the window that needs the exact evaluation
    const float s_norm = sqrtf(T->s2_sum);             // (once per scan, not per window)
    const float k_err = s_norm, k_thr = (T->threshold - 0.008f) * sqrtf(156.0f / 512.0f) * s_norm;
    long long g_w = st.wstart;                         // block generator: next window to stage
    bool g_cold = !(c.valid && c.w == g_w);
    long long w_done = g_w;                            // first window not yet decided
    if (g_w + 512 > sv.end || g_w >= st.stop_at) return 0;
    sh.tables_dirty = 1;                               // (every worker writes the same value)
    if (!c.ex_valid) {
        // first strip of this launch: take the scale from the samples themselves
        float mx = 0.0f;
        const long long a0 = g_cold ? g_w - 28 : g_w + 100;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = tid + kThreads * k;
            const float2 v = (i < 424) ? sv.at(a0 + i) : make_float2(0.0f, 0.0f);
            mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        wsync();
        if ((tid & 31) == 0) sh.tc_red[tid >> 5] = mx;
        wsync();
        mx = fmaxf(fmaxf(sh.tc_red[0], sh.tc_red[1]), fmaxf(sh.tc_red[2], sh.tc_red[3]));
        c.ex = exponent_of(mx);
        c.ex_valid = true;
    }
    TcBlk q0, q1, nb;                                   // q0, q1: handed to the MMA warp (q0 the older); nb: stage A done only
    int inflight = 0, result = 0, slot_w = 0;
    auto gen = [&](TcBlk &b) {
        b.w = g_w; b.ex = c.ex; b.buf = buf_w; buf_w ^= 1; b.slot = slot_w; slot_w = (slot_w + 1) & 3; b.wrap_w = kNoWrap;
        if (g_cold) { b.a0 = g_w - 28; b.n_t = 1; b.e_lo = g_w; b.e_hi = g_w + 256; b.cold = 1; b.next_a0 = g_w + 100; g_cold = false; }
        else { b.a0 = g_w + 100; b.n_t = 2; b.e_lo = g_w + 256; b.e_hi = g_w + 512; b.cold = 0; b.next_a0 = g_w + 356; g_w += 256; }
        n_tiles += (unsigned)b.n_t;
    };
    const long long stop_at = st.stop_at;
    auto more = [&]() { return g_w + 512 <= sv.end && g_w < stop_at; };
    auto block_max = [&](const TcBlk &b) {
        return fmaxf(fmaxf(sh.part_max[b.slot][0], sh.part_max[b.slot][1]), fmaxf(sh.part_max[b.slot][2], sh.part_max[b.slot][3]));
    };
    while (inflight < 2 && more()) {
        TcBlk &b = inflight ? q1 : q0;
        gen(b);
        tc_stage_a(sh, sv, b, tid, pre, pre_a0 PROF_PASS);
        tc_stage_b(sh, sv, b, tid, pre, pre_a0 PROF_PASS);
        // the scale of the following blocks follows the samples (statistics are visible: stage A ends behind a barrier)
        c.ex = exponent_of(block_max(b));
        ++inflight;
    }
    while (inflight) {
        const TcBlk d = q0;
        const float best0 = tc_retire_tile(sh, ph_full, tid, false PROF_PASS);
        bool have_nb = false;
#if LQB_SEEK_REORDER
        if (more()) {
            gen(nb);
            tc_stage_a(sh, sv, nb, tid, pre, pre_a0 PROF_PASS);          // runs while tile 1 of d is multiplied
            have_nb = true;
        }
#endif
        const float best1 = d.n_t == 2 ? tc_retire_tile(sh, ph_full, tid, false PROF_PASS) : 0.0f;
        TcRes r;
        tc_retire_reduce(sh, d, best0, best1, tid, r PROF_PASS);
        q0 = q1; --inflight;
        // accumulators hold sum (x 2^(4-ex)) (t 2^5): |C|^2 = acc^2 2^(2 ex - 18); residual energy in sample units: 2^(2 ex - 8)
        const float s2 = __uint_as_float((uint32_t)(127 + 2 * d.ex - 8 - 2 * tc::kBScaleLog2) << 23);
        const float d_true = r.span_d * __uint_as_float((uint32_t)(127 + 2 * d.ex - 8) << 23);
        const float scaled_mx = r.mx * __uint_as_float((uint32_t)(127 + 4 - d.ex) << 23);
        // a block whose scaled maximum reaches e4m3's saturation (448) has an unbounded rounding error: do not trust it
        // (small values need no such test: their rounding error, including flush to zero, is inside the residual)
        const bool unsafe_blk = !(scaled_mx < 384.0f);
        bool skip = true;
        if (d.cold) {
            c.tail = r.m0_ge28 * s2; c.tail_d = d_true; c.tail_e = r.span_e; c.half = r.en; c.unsafe = unsafe_blk; c.w = d.w; c.valid = true;
        } else {
            ++n_windows;
            const float mm = fmaxf(c.tail, fmaxf(r.m0_all, r.m1_all) * s2), E = c.half + r.en;
            skip = false;
            if (E > 0.0f && !c.unsafe && !unsafe_blk) {
                // rxy_q + margin < thr - 0.008, multiplied through by g0 ||s|| (no division, four approximate square
                // roots -- their 2^-22 relative error and the rounding of the sums are covered by the factor 1.001):
                //   sqrt(mm) + ||s|| (sqrt(D6) + (sqrt(E6) + sqrt(D6)) b_err) < (thr - 0.008) sqrt(156 / 512) ||s|| sqrt(E)
                const float rd = fast_sqrt(fmaxf(c.tail_d, d_true)), re = fast_sqrt(fmaxf(c.tail_e, r.span_e));
                const float lhs = fast_sqrt(mm) + k_err * (rd + (re + rd) * b_err);
                skip = lhs * 1.001f < k_thr * fast_sqrt(E);
                if (!skip) rxy_q = fast_sqrt(mm) / (fast_sqrt(E) * sqrtf(156.0f / 512.0f) * s_norm);
            }
            c.tail = r.m1_ge28 * s2; c.tail_d = d_true; c.tail_e = r.span_e; c.half = r.en; c.unsafe = unsafe_blk; c.w = d.w + 256; c.valid = true;
        }
        if (!skip) {
            // speculative work behind this window is dropped: a block that only went through stage A was never handed
            // to the MMA warp (undo its strip-buffer turn), handed-over blocks have their accumulators released unread
            if (have_nb) { buf_w ^= 1; n_tiles -= (unsigned)nb.n_t; }
            while (inflight) {
                for (int t = 0; t < q0.n_t; ++t) (void)tc_retire_tile(sh, ph_full, tid, true PROF_PASS);
                q0 = q1; --inflight;
            }
            w_done = d.w;
            result = 1;
            break;
        }
        if (!d.cold) w_done = d.w + 256;
#if !LQB_SEEK_REORDER
        if (more()) {
            gen(nb);
            tc_stage_a(sh, sv, nb, tid, pre, pre_a0 PROF_PASS);
            have_nb = true;
        }
#endif
        if (have_nb) {
            tc_stage_b(sh, sv, nb, tid, pre, pre_a0 PROF_PASS);          // into the strip buffer d used: its MMAs are complete
            c.ex = exponent_of(block_max(nb));
            if (inflight) q1 = nb; else q0 = nb;
            ++inflight;
        }
    }
    if (tid == 0) st.wstart = w_done;
    wsync();
    return result;
}

// per-bin form of the accumulator read-out: every thread keeps max |C_q|^2 per CFO bin over ITS lag of each of the four
// tiles in registers (acc[b], float bit patterns); the reduction over the 128 lags happens once, after the last tile.
__device__ void tc_retire_tile_bins(SeekShared &sh, unsigned &ph_full, int tid, unsigned (&acc)[tc::kNBins])
{
    using namespace tc;
    mbar_wait(&sh.acc_full, ph_full, 40u * LQB_SEEK_SLEEP);
    ph_full ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = sh.tmem_base + ((uint32_t)((tid >> 5) * 32) << 16);
#pragma unroll
    for (int c8 = 0; c8 < 7; ++c8) {
        uint32_t re[8], im[8];
        tmem_ld8(taddr + 8u * c8, re);
        tmem_ld8(taddr + (uint32_t)kImCol0 + 8u * c8, im);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int bin = 8 * c8 + k;                                          // compile-time
            if (bin < kNBins) {
                const float r_ = __uint_as_float(re[k]), i_ = __uint_as_float(im[k]);
                acc[bin] = max(acc[bin], __float_as_uint(fmaf(i_, i_, r_ * r_)));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&sh.acc_empty);
}

// Which CFO bins can hold the global maximum of the window at w (the exact evaluation's arg max over all 512 CIRCULAR
// lags x 49 bins)?  The window is not skippable, so its exact evaluation would cost 49 inverse FFTs; the tensor cores
// first correlate all 512 circular lags (two 256-lag blocks, the second reading the window modulo 512), keep the
// maximum per bin, and the same rigorous e4m3 error bound as the pre-filter's rules out every bin whose largest
// possible value lies below the smallest possible value of the best bin.  Typically 3 .. 7 bins survive.
// Returns the bin mask (all bins when a block cannot be trusted).  The MMA pipeline must be empty on entry and is on exit.
__device__ unsigned long long bin_candidates(SeekShared &sh, const StreamView &sv, const DevTables *T, float b_err, long long w,
                                             int tid, float2 (&pre)[4], long long &pre_a0, unsigned &ph_full, int &buf_w,
                                             unsigned &n_tiles PROF_ARGS)
{
    const unsigned long long all = (1ull << (2 * T->range + 1)) - 1ull;
    for (int i = tid; i < tc::kNBins + 7; i += kThreads) sh.binmax[i] = 0u;
    // the scale comes from the window itself (a window that is not skippable usually holds the start of a frame, whose
    // level the pre-filter's running scale has not seen yet)
    int ex;
    {
        float m = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 v = sv.at(w + tid + kThreads * k);
            m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y)));
        }
        m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));
        wsync();
        if ((tid & 31) == 0) sh.tc_red[tid >> 5] = m;
        wsync();
        ex = exponent_of(fmaxf(fmaxf(sh.tc_red[0], sh.tc_red[1]), fmaxf(sh.tc_red[2], sh.tc_red[3])));
    }
    TcBlk b[2];
    float worst_d = 0.0f, worst_e = 0.0f, mx = 0.0f, en = 0.0f;
    for (int h = 0; h < 2; ++h) {
        TcBlk &q = b[h];
        q.w = w; q.a0 = w + 256 * h; q.e_lo = q.a0; q.e_hi = q.a0 + 256; q.next_a0 = kNoWrap;
        q.n_t = 2; q.cold = 0; q.ex = ex; q.buf = buf_w; buf_w ^= 1; q.slot = h; q.wrap_w = w;
        n_tiles += 2u;
        tc_stage_a(sh, sv, q, tid, pre, pre_a0 PROF_PASS);
        tc_stage_b(sh, sv, q, tid, pre, pre_a0 PROF_PASS);
    }
    {
        unsigned acc[tc::kNBins];
#pragma unroll
        for (int i = 0; i < tc::kNBins; ++i) acc[i] = 0u;
        for (int t = 0; t < 4; ++t) tc_retire_tile_bins(sh, ph_full, tid, acc);
        const int lane = tid & 31;
        unsigned mine0 = 0u, mine1 = 0u;               // lane b & 31 collects bin b
#pragma unroll
        for (int i = 0; i < tc::kNBins; ++i) {
            const unsigned m = __reduce_max_sync(0xffffffffu, acc[i]);
            if (i < 32) mine0 = (lane == i) ? m : mine0; else mine1 = (lane == i - 32) ? m : mine1;
        }
        atomicMax(&sh.binmax[lane], mine0);                    // four warps x two instructions
        if (lane < tc::kNBins - 32) atomicMax(&sh.binmax[32 + lane], mine1);
    }
    wsync();
    for (int h = 0; h < 2; ++h) {
        worst_d = fmaxf(worst_d, sh.span_d[h]); worst_e = fmaxf(worst_e, sh.span_e[h]);
#pragma unroll
        for (int k = 0; k < kWarps; ++k) { mx = fmaxf(mx, sh.part_max[h][k]); en += sh.part_en[h][k]; }
    }
    const float scaled_mx = mx * __uint_as_float((uint32_t)(127 + 4 - ex) << 23);
    unsigned long long cand = all;
    if (scaled_mx < 384.0f && en > 0.0f) {
        // everything in |C| units: |C_q| = sqrt(binmax) 2^(ex - 9); err as in fused_scan; slack for the fp32 rounding of
        // the exact evaluation itself (0.004 in rxy units on either side)
        const float s2 = __uint_as_float((uint32_t)(127 + 2 * ex - 8 - 2 * tc::kBScaleLog2) << 23);
        const float d_true = worst_d * __uint_as_float((uint32_t)(127 + 2 * ex - 8) << 23);
        const float s_norm = sqrtf(T->s2_sum);
        const float rd = fast_sqrt(d_true), re = fast_sqrt(worst_e);
        const float err = (s_norm * (rd + (re + rd) * b_err) + 0.004f * fast_sqrt(en * (156.0f / 512.0f)) * s_norm) * 1.001f;
        float best = 0.0f;
        for (int i = 0; i < tc::kNBins; ++i) best = fmaxf(best, __uint_as_float(sh.binmax[i]));
        float floor_ = fast_sqrt(best * s2) - err;
#if LQB_SEEK_THR_PRUNE
        // A bin matters only if it can TRIGGER: the exact evaluation's result is used for nothing else (a window that
        // does not trigger just moves the hop grid on), and when something does trigger the global maximum is at least the
        // threshold, so no bin that cannot reach the threshold can hold it.  In |C| units the threshold is
        // thr * g0 * ||s||, g0 = sqrt(E 156 / 512) of this very window.  An EMPTY set proves "no trigger" outright.
        floor_ = fmaxf(floor_, T->threshold * fast_sqrt(en * (156.0f / 512.0f)) * s_norm * 0.999f);
#endif
        cand = 0ull;
        for (int i = 0; i <= 2 * T->range; ++i)
            if (fast_sqrt(__uint_as_float(sh.binmax[i]) * s2) + err >= floor_) cand |= 1ull << i;
#if !LQB_SEEK_THR_PRUNE
        if (!cand) cand = all;
#endif
    }
    wsync();                                   // binmax / statistics slots are free again
    return cand;
}

// alignment on the 512 samples in sh.Xw (x[F .. F+512)) with CFO bin sh.off
__device__ void align_frame(SeekShared &sh, const DevTables *T, int tid PROF_ARGS)
{
    const int warp = tid >> 5, lane = tid & 31;
    float2 *zbuf = sh.scr + kZbuf, *vsum = sh.scr + kVsum;
    if (warp == 0) forward_fft_to(sh, sh.Xw, sh.Xf, lane);
    // CFO product (other warps start while warp 0 transforms; zbuf does not alias warp 0's scratch)
    for (int i = tid; i < 512; i += kThreads)
        zbuf[i] = (i < 156) ? cmulf(sh.Xw[i], T->sconj[i]) : make_float2(0.0f, 0.0f);
    wsync();
    PROF_MARK(17);
    if (warp == 0) {
        float2 v[16];
        cross_ifft(sh, sh.off, v, sh.scr, lane);
        if (lane == 0) sh.y3[1] = v[0];
        if (lane == 1) sh.y3[2] = v[0];
        if (lane == 31) sh.y3[0] = v[15];
    } else if (warp == 1) {
        // CFO spectrum, at the same time on its own warp: only the peak bin and its two neighbours are kept
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = zbuf[fft512_in_index(lane, r)];
        fft512_warp<+1>(v, sh.W, sh.Wc, sh.scr + 544, lane);
        unsigned long long best = 0ull;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int p = fft512_out_index(lane, r);
            unsigned long long key = ((unsigned long long)__float_as_uint(abs2f(v[r])) << 32) | (0xffffffffu - (unsigned)p);
            best = key > best ? key : best;
        }
        best = warp_max_u64(best);
        const float v2 = __uint_as_float((unsigned)(best >> 32));
        const unsigned i0 = (v2 > 0.0f) ? (0xffffffffu - (unsigned)(best & 0xffffffffu)) : 0u;
        const unsigned pn = (i0 + 511u) & 511u, pp = (i0 + 1u) & 511u;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const unsigned p = (unsigned)fft512_out_index(lane, r);
            if (p == pn) sh.cfo_nb[0] = cabsf_(v[r]);
            if (p == pp) sh.cfo_nb[1] = cabsf_(v[r]);
        }
        if (lane == 0) sh.best[0] = best;
    }
    wsync();
    PROF_MARK(18);
    if (tid == 0) {
        float yneg = __fsqrt_rn(cabsf_(sh.y3[0])), y0 = __fsqrt_rn(cabsf_(sh.y3[1])), ypos = __fsqrt_rn(cabsf_(sh.y3[2]));
        float a = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(ypos, yneg)), y0);
        float b = __fmul_rn(0.5f, __fsub_rn(ypos, yneg));
        float c = y0;
        float tau = __fdiv_rn(-b, __fmul_rn(2.0f, a));
        float g_hat = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(a, tau), tau), __fmul_rn(b, tau)), c);
        sh.tau = tau;
        sh.gamma = __fdiv_rn(__fmul_rn(g_hat, g_hat), __fmul_rn(512.0f, T->s2_sum));
        unsigned long long bb = sh.best[0];
        float v2 = __uint_as_float((unsigned)(bb >> 32));
        unsigned i0 = (v2 > 0.0f) ? (0xffffffffu - (unsigned)(bb & 0xffffffffu)) : 0u;
        float v0 = __fsqrt_rn(v2);
        float vneg = sh.cfo_nb[0], vpos = sh.cfo_nb[1];
        a = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(vpos, vneg)), v0);
        b = __fmul_rn(0.5f, __fsub_rn(vpos, vneg));
        float idx = __fdiv_rn(-b, __fmul_rn(2.0f, a));
        float index = __fadd_rn((float)i0, idx);
        float base = (i0 > 256u) ? __fsub_rn(index, 512.0f) : index;
        sh.dphi = __fdiv_rn(__fmul_rn(__fmul_rn(base, 2.0f), kPiF), 512.0f);
    }
    wsync();
    const float dphi = sh.dphi;
    for (int i = tid; i < 156; i += kThreads) {
        float ang = __fmul_rn(-dphi, (float)i);
        float sn, cs;
        pm_sincosf(ang, &sn, &cs);
        vsum[i] = cmulf(zbuf[i], make_float2(cs, sn));
    }
    wsync();
    PROF_MARK(19);
    if (tid == 0) {
        // summed in sample order (the specification's order); unrolled so the shared-memory loads run ahead of the adds
        float mr = 0.0f, mi = 0.0f;
#pragma unroll 12
        for (int i = 0; i < 156; ++i) { const float2 q = vsum[i]; mr = __fadd_rn(mr, q.x); mi = __fadd_rn(mi, q.y); }
        sh.phi = pm_atan2f(mi, mr);
        sh.theta0 = nco_constrain_dev(sh.phi);
        sh.dtheta = nco_constrain_dev(sh.dphi);
        if (sh.tau > 0.0f) {
            sh.pfb = (unsigned)__float2int_rz(__fmul_rn(sh.tau, 32.0f)) % 32u;
            sh.tau_neg = 0u;
        } else {
            sh.pfb = (unsigned)__float2int_rz(__fmul_rn(__fadd_rn(1.0f, sh.tau), 32.0f)) % 32u;
            sh.tau_neg = 1u;
        }
        sh.mf_scale = __fdiv_rn(0.5f, sh.gamma);
    }
    wsync();
}

// in-place interleaver pass on a small byte buffer (serial; header only)
__device__ void hdr_ilv_pass(unsigned char *x, const uint16_t *map, unsigned n2, unsigned mask)
{
    for (unsigned i = 0; i < n2; ++i) {
        unsigned j = map[i];
        unsigned a = x[2 * j + 1], b = x[2 * i];
        x[2 * j + 1] = (unsigned char)((a & ~mask) | (b & mask));
        x[2 * i] = (unsigned char)((a & mask) | (b & ~mask));
    }
}

// header: matched filter over x[F+128 .. F+617), pilot sync, QPSK demod, decode.
// Publishes sh.hv and the decoded bytes; returns PLL init through pll_theta0/pll_dtheta.
__device__ void decode_header(SeekShared &sh, const DevTables *T, const StreamView &sv, long long F, int tid,
                              unsigned *pll_theta0, unsigned *pll_dtheta PROF_ARGS)
{
    float2 *vbuf = sh.scr + kVbuf, *hsym = sh.scr + kHsym, *pil = sh.scr + kPil;
    const unsigned theta0 = sh.theta0, dtheta = sh.dtheta, tau_neg = sh.tau_neg;
    for (int m = tid; m < 489; m += kThreads) {
        unsigned n = 128u + (unsigned)m;
        vbuf[m] = nco_mix_down(T->sintab, theta0 + n * dtheta, sv.at(F + n));
    }
    wsync();
    const float *h = T->banks + sh.pfb * 28;
    for (int k = tid; k < 231; k += kThreads) {
        int nt = 2 * (78 + k) - (int)tau_neg;           // sample index of symbol t = 78 + k
        const float2 *w = vbuf + (nt - 27 - 128);
        float ar = 0.0f, ai = 0.0f;
#pragma unroll 4
        for (int j = 0; j < 28; ++j) { ar = __fmaf_rn(h[j], w[j].x, ar); ai = __fmaf_rn(h[j], w[j].y, ai); }
        hsym[k] = make_float2(__fmul_rn(ar, sh.mf_scale), __fmul_rn(ai, sh.mf_scale));
    }
    wsync();
    PROF_MARK(20);
    // ---- pilot sync (warp 0): 15 pilots -> FFT-32 (one point per lane, the specification's radix-2 DIT butterflies
    // exchanged by shuffle) -> residual dphi / phi / gain
    if (tid < 32) {
        const unsigned lane = (unsigned)tid;
        const float2 bt = (lane < 15u) ? cmulf(hsym[16 * lane], T->pilots_conj[lane]) : make_float2(0.0f, 0.0f);
        // lane i starts with bt[brev5(i)]
        const unsigned srcl = brev5(lane);
        float2 v = make_float2(__shfl_sync(0xffffffffu, bt.x, srcl), __shfl_sync(0xffffffffu, bt.y, srcl));
#pragma unroll
        for (int half = 1; half < 32; half <<= 1) {
            const float2 w = T->W32[(lane & (half - 1)) * (16 / half)];
            const float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v.x, half), __shfl_xor_sync(0xffffffffu, v.y, half));
            const bool hi_side = (lane & half) != 0;
            float2 lo = hi_side ? o : v, hi = hi_side ? v : o;
            bfly<+1>(lo, hi, w);
            v = hi_side ? hi : lo;
        }
        const float a_me = cabsf_(v);
        // first maximum wins
        float y0 = a_me; unsigned i0 = lane;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const float oy = __shfl_xor_sync(0xffffffffu, y0, m);
            const unsigned oi = __shfl_xor_sync(0xffffffffu, i0, m);
            if (oy > y0 || (oy == y0 && oi < i0)) { y0 = oy; i0 = oi; }
        }
        const float ypos = __shfl_sync(0xffffffffu, a_me, (i0 + 1) & 31), yneg = __shfl_sync(0xffffffffu, a_me, (i0 + 31) & 31);
        float a = __fsub_rn(__fmul_rn(0.5f, __fadd_rn(ypos, yneg)), y0);
        float b = __fmul_rn(0.5f, __fsub_rn(ypos, yneg));
        float idx = __fdiv_rn(-b, __fmul_rn(2.0f, a));
        float index = __fadd_rn((float)i0, idx);
        float base = (i0 > 16u) ? __fsub_rn(index, 32.0f) : index;
        float dphi = __fdiv_rn(__fmul_rn(__fmul_rn(base, 2.0f), kPiF), 512.0f);
        // derotated pilots, one per lane, then summed in pilot order
        float2 r = make_float2(0.0f, 0.0f);
        if (lane < 15u) {
            float ang = __fmul_rn(__fmul_rn(-dphi, (float)lane), 16.0f);
            float sn, cs;
            pm_sincosf(ang, &sn, &cs);
            r = cmulf(bt, make_float2(cs, sn));
        }
        float mr = 0.0f, mi = 0.0f;
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            mr = __fadd_rn(mr, __shfl_sync(0xffffffffu, r.x, i));
            mi = __fadd_rn(mi, __shfl_sync(0xffffffffu, r.y, i));
        }
        if (lane == 0) {
            float phi = pm_atan2f(mi, mr);
            float g_hat = __fdiv_rn(cabsf_(make_float2(mr, mi)), 15.0f);
            pil[0] = make_float2(dphi, phi);
            pil[1] = make_float2(__fdiv_rn(1.0f, g_hat), 0.0f);
            *pll_dtheta = nco_constrain_dev(dphi);
            *pll_theta0 = nco_constrain_dev(__fadd_rn(phi, __fmul_rn(dphi, 231.0f)));
        }
    }
    wsync();
    // ---- derotate the 216 data symbols and slice QPSK: 4 symbols -> one byte
    {
        const float dphi = pil[0].x, phi = pil[0].y, g = pil[1].x;
        for (int byte = tid; byte < 54; byte += kThreads) {
            unsigned out = 0;
            for (int q = 0; q < 4; ++q) {
                int n = 4 * byte + q;
                int i = n + n / 15 + 1;                  // position in the pilot-bearing frame
                float ang = -__fadd_rn(__fmul_rn(dphi, (float)i), phi);
                float sn, cs;
                pm_sincosf(ang, &sn, &cs);
                float2 v = cmulf(hsym[i], make_float2(cs, sn));
                v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g);
                unsigned s = (v.x > 0.0f ? 0u : 1u) + (v.y > 0.0f ? 0u : 2u);
                out = (out << 2) | s;
            }
            sh.hbytes[byte] = (unsigned char)out;
        }
    }
    wsync();
    PROF_MARK(21);
    // ---- decode: deinterleave(54) -> Hamming(8,4) -> deinterleave(27) -> SECDED(72,64) -> unscramble -> CRC-32.
    // Each deinterleaver is one gather through the precomputed bit permutation (thread per output byte); the three
    // SECDED blocks decode on three threads; only the 20-byte CRC is serial.
    {
        unsigned char *e = sh.hbytes, *d = sh.hdec;
        unsigned char *e2 = reinterpret_cast<unsigned char *>(pil + 8);         // 54 deinterleaved bytes (pil[0..1] stay live)
        unsigned char *d27 = e2 + 64, *d27b = d27 + 32;
        if (tid < 54) {
            unsigned v = 0;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const unsigned src = T->hperm54[8 * tid + bit];
                v |= ((e[src >> 3] >> (src & 7u)) & 1u) << bit;
            }
            e2[tid] = (unsigned char)v;
        }
        wsync();
        if (tid < 27) d27[tid] = (unsigned char)((T->h84_dec[e2[2 * tid]] << 4) | T->h84_dec[e2[2 * tid + 1]]);
        wsync();
        if (tid < 27) {
            unsigned v = 0;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const unsigned src = T->hperm27[8 * tid + bit];
                v |= ((d27[src >> 3] >> (src & 7u)) & 1u) << bit;
            }
            d27b[tid] = (unsigned char)v;
        }
        wsync();
        if (tid < 3) {
            const int blk = tid;
            const unsigned char *src = d27b + 9 * blk;
            // liquid's Hsiao (72,64) code: syndrome = parity of the data XOR the received parity byte; a syndrome that
            // equals a column of P is that data bit (flipped back), anything else is left as received
            const unsigned char *col = T->secded_col[2];
            unsigned p = 0;
            unsigned char b8[8];
            for (int q = 0; q < 8; ++q) b8[q] = src[1 + q];
            for (int bit = 0; bit < 64; ++bit)
                if ((b8[bit >> 3] >> (7 - (bit & 7))) & 1u) p ^= col[bit];
            const unsigned syn = (p ^ src[0]) & 0xffu;
            if (syn)
                for (int bit = 0; bit < 64; ++bit)
                    if (col[bit] == syn) { b8[bit >> 3] ^= (unsigned char)(0x80u >> (bit & 7)); break; }
            const unsigned char mask[4] = { 0xb4, 0x6a, 0x8b, 0xc5 };
            for (int q = 0; q < 8; ++q) d[8 * blk + q] = b8[q] ^ mask[q & 3];      // 8 blk is a multiple of 4: unscramble in place
        }
        wsync();
        if (tid == 0) {
            unsigned key = 0xffffffffu;
            for (int i = 0; i < 20; ++i) key = (key >> 8) ^ T->crc_tab[6][(key ^ d[i]) & 0xffu];
            key = ~key;
            unsigned rx = ((unsigned)d[20] << 24) | ((unsigned)d[21] << 16) | ((unsigned)d[22] << 8) | d[23];
            sh.hv = (key == rx);
        }
    }
    wsync();
}

}  // namespace

// ------------------------------------------------------------------ the kernel
// 4 worker warps walk the stream's state machine; a fifth warp only issues tcgen05.mma for the
// pre-filter pipeline (fused mode) and otherwise idles until the end of the CTA.
#ifndef LQB_SEEK_CTAS_PER_SM
#define LQB_SEEK_CTAS_PER_SM 3
#endif
__global__ void __launch_bounds__(kCtaThreads, LQB_SEEK_CTAS_PER_SM)
k_seek(SeekParams P)
{
    __shared__ SeekShared sh;
    __shared__ __align__(16) StreamState st;
    __shared__ unsigned pll_theta0, pll_dtheta;
    const int tid = threadIdx.x;
    PROF_DECL;
    const DevTables *T = P.tables;
    // Time slices (P.slice_len != 0): the launch has one CTA per slot of the GPU, not one per stream; a CTA takes the
    // next READY stream from a queue, walks it for about slice_len samples, hands it back and takes the next one.  The
    // work items become short and alike, so a call with more streams than the GPU has CTA slots keeps every slot busy
    // to the end (no wave of leftovers).  queue[0]: next ticket, [1]: next free entry, [2]: finished streams,
    // [4 + t]: io index behind ticket t.  A stream's state passes from CTA to CTA through global memory (release by
    // the CTA that hands it back, acquire by the one that takes it).
    __shared__ unsigned s_item;                 // io entry this CTA works on
    __shared__ long long s_saved_stop;          // the stream's own stop_at while the slice boundary stands in for it

    extern __shared__ unsigned char dyn_smem[];
    unsigned char *Bsm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(dyn_smem) + 127) & ~uintptr_t(127));
    const bool fused = (P.coarse == 2);
    if (fused) {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.bmat);
        uint4 *dst = reinterpret_cast<uint4 *>(Bsm);
        for (int i = tid; i < tc::kBBytes / 16; i += kCtaThreads) dst[i] = src[i];
        if (tid == 0) {
            tc::mbar_init(&sh.z_full[0], kThreads); tc::mbar_init(&sh.z_full[1], kThreads);
            tc::mbar_init(&sh.acc_full, 1); tc::mbar_init(&sh.acc_empty, kThreads);
            sh.tc_cmd[0] = 0; sh.tc_cmd[1] = 0;
        }
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc::smem_u32(&sh.tmem_base)), "r"(128));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (tid == 0) sh.tables_dirty = 1;
    __syncthreads();
    if (fused) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid >= kThreads) {
        if (fused) mma_warp(sh, Bsm, tid & 31);
    } else {
    // ======================================================== worker warps
    // compact twiddles survive the strips; W itself is (re)loaded together with Sc on demand
    for (int i = tid; i < 240; i += kThreads) {
        int sidx = (i >= 112) ? 3 : (i >= 48) ? 2 : (i >= 16) ? 1 : 0;
        int j = i - 16 * ((1 << sidx) - 1);
        sh.Wc[i] = T->W512[j * (16 >> sidx)];
    }
    wsync();
    ScanCarry sc_;
    float2 tc_pre[4];
    long long tc_pre_a0;
    unsigned ph_full = 0u;
    int buf_w = 0;
    float rxy_q = 0.0f;
    unsigned n_windows = 0, n_aligns = 0, n_exact = 0, n_tc_tiles = 0, n_bins = 0;      // work counters (uniform across the workers)
    PROF_MARK(10);

    // make Sc / W valid again after a strip overwrote them
    auto restore_tables = [&]() {
        if (sh.tables_dirty) {
            wsync();
            for (int i = tid; i < 512; i += kThreads) sh.Sc[i] = T->Sc[i];
            for (int i = tid; i < 256; i += kThreads) sh.W[i] = T->W512[i];
            if (tid == 0) sh.tables_dirty = 0;
            wsync();
        }
    };

    for (;;) {       // one pass (a CTA per stream) or one pass per slice taken from the queue
#ifdef LQB_SEEK_TRACE
    const unsigned long long trace_t0 = gtime_ns();
    const unsigned trace_w0 = n_windows;
#endif
    if (tid == 0) {
        unsigned v = blockIdx.x;
        if (P.slice_len) {
            const unsigned t = atomicAdd(P.queue + 0, 1u);
            unsigned polls = 0;
            if (t >= P.queue_cap) { atomicAdd(&g_seek_stall[0], 1u << 16); v = kQueueEmpty; }       // cannot happen: see launch_seek
            else
            while ((v = ld_acquire_u32(P.queue + 4 + t)) == kQueueEmpty) {
                if (ld_acquire_u32(P.queue + 2) >= P.n_io) break;       // every stream is through
                __nanosleep(200);
                if (++polls > kQueueMaxPolls) {
                    if (atomicAdd(&g_seek_stall[0], 1u) == 0u) {
                        g_seek_stall[1] = t; g_seek_stall[2] = ld_acquire_u32(P.queue + 0); g_seek_stall[3] = ld_acquire_u32(P.queue + 1);
                        g_seek_stall[4] = ld_acquire_u32(P.queue + 2); g_seek_stall[5] = P.n_io; g_seek_stall[6] = P.grid; g_seek_stall[7] = blockIdx.x;
                    }
                    break;
                }
            }
        }
        s_item = v;
        if (v != kQueueEmpty) {
            if (P.slice_len) {
                // the state was written by another CTA of this launch: nothing of it may come from this SM's L1
                __threadfence();
                const uint4 *gs = reinterpret_cast<const uint4 *>(P.states + P.io[v].stream);
                uint4 *ss = reinterpret_cast<uint4 *>(&st);
                static_assert(sizeof(StreamState) % 16 == 0, "StreamState is copied in 16-byte words");
#pragma unroll
                for (unsigned k = 0; k < sizeof(StreamState) / 16; ++k) ss[k] = __ldcg(gs + k);
            } else
            st = P.states[P.io[v].stream];
            s_saved_stop = st.stop_at;
            if (P.slice_len) {
                // the slice ends at the first window start slice_len beyond where this CTA picks the stream up (only
                // the search stops there: a frame in progress is finished first)
                const long long lim = (st.mode == 0 ? st.wstart : st.F) + (long long)P.slice_len;
                if (lim < st.stop_at) st.stop_at = lim;
            }
        }
    }
    wsync();
    if (s_item == kQueueEmpty) break;
    const StreamIO io = P.io[s_item];
    sc_.valid = false; sc_.unsafe = false; sc_.ex_valid = false; sc_.w = 0; sc_.tail = 0.0f; sc_.tail_d = 0.0f; sc_.tail_e = 0.0f; sc_.half = 0.0f; sc_.ex = 0;
    tc_pre_a0 = -(1ll << 62);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_pre[k] = make_float2(0.0f, 0.0f);
    StreamView sv;
    sv.carry = P.carry[st.carry_sel] + (size_t)io.stream * P.carry_cap;
    sv.in = io.in;
    sv.base = st.base;
    sv.carry_len = st.carry_len;
    sv.end = st.base + (long long)st.carry_len + (long long)io.n_in;
    sv.G = st.G;
    if (tid == 0 && P.views) P.views[s_item] = sv;
    PROF_MARK(10);

    while (true) {
        // ---------------- SEEK
        if (st.mode == 0) {
            // (sharded search) the first window start at or beyond mark_at that the walk reaches, and the stop boundary
            const long long w_in = st.wstart;
            if (tid == 0 && st.mark_w == kNoMark && w_in >= st.mark_at) { st.mark_w = w_in; st.mark_G = st.G; }
            if (w_in + 512 > sv.end || w_in >= st.stop_at) break;
            if (fused) {
                const int hit = fused_scan(sh, sv, T, P.b_err, st, tid, sc_, tc_pre, tc_pre_a0, ph_full, buf_w, n_windows, n_tc_tiles, rxy_q PROF_PASS);
                // the scan walked the grid w_in, w_in + 256, .. up to st.wstart: if it passed mark_at, the mark is on that grid
                if (tid == 0 && st.mark_w == kNoMark && st.wstart >= st.mark_at) {
                    st.mark_w = w_in + 256ll * ((st.mark_at - w_in + 255ll) / 256ll);
                    st.mark_G = st.G;
                }
                if (!hit) break;
            } else {
                ++n_windows;
            }
            ++n_exact;
            PROF_MARK(7);
            // (uniform) which CFO bins the exact evaluation has to visit
            unsigned long long cand = (1ull << (2 * T->range + 1)) - 1ull;
            // Worth it only when the pre-filter saw a peak well above the threshold (a real preamble: 3 .. 9 bins
            // survive); a window that merely could not be ruled out has a noise-like spectrum in which nearly every bin
            // stays a candidate, and the scan would be pure overhead.  Measured (B200): flex_rx search 22.1 -> 20.9 ms;
            // the bare detector, where two of three exact windows are of the second kind, 87.3 -> 89.6 ms: off there.
#if LQB_SEEK_THR_PRUNE
            // Every window the pre-filter could not rule out: with the threshold in the candidate rule a noise-like window
            // (one in five in flex_rx, two in three in the bare detector) keeps the one or two bins that can still reach
            // it -- or none, which proves it cannot trigger -- instead of all 49.
            if (fused && LQB_SEEK_BINS) {
                cand = bin_candidates(sh, sv, T, P.b_err, st.wstart, tid, tc_pre, tc_pre_a0, ph_full, buf_w, n_tc_tiles PROF_PASS);
                sh.tables_dirty = 1;
                if (cand == 0ull) {                  // (uniform) proven: no bin of this window reaches the threshold
                    --n_exact;
                    if (tid == 0) st.wstart += 256;
                    wsync();
                    PROF_MARK(11);
                    continue;
                }
            }
#else
            if (fused && LQB_SEEK_BINS && !P.det_mode && rxy_q > T->threshold + 0.1f) {
                cand = bin_candidates(sh, sv, T, P.b_err, st.wstart, tid, tc_pre, tc_pre_a0, ph_full, buf_w, n_tc_tiles PROF_PASS);
                sh.tables_dirty = 1;
            }
#endif
            PROF_MARK(11);
            n_bins += (unsigned)__popcll(cand);
            restore_tables();
            load_window(sh, sv, st.wstart, tid);
            wsync();
            PROF_MARK(12);
            eval_window(sh, T, tid, cand PROF_PASS);
            PROF_MARK(8);
            if (!sh.trig) {
                if (tid == 0) st.wstart += 256;
                wsync();
                continue;
            }
            if (tid == 0) {
                st.mode = 1;
                st.F = st.wstart + sh.idx;
                st.det_idx = (unsigned)sh.idx;
                st.offset = sh.off;
                st.rxy = sh.rxy;
                st.need_until = st.F + 512;
            }
            wsync();
        }
        // ---------------- PENDING: frame start known (the hop grid restarts afterwards)
        sc_.valid = false;
        tc_pre_a0 = -(1ll << 62);          // prefetched samples were taken under the old zero boundary
        if (sv.end < st.need_until) break;
        PROF_MARK(7);
        restore_tables();
        const long long F = st.F;
        load_window(sh, sv, F, tid);
        if (tid == 0) sh.off = st.offset;
        wsync();
        align_frame(sh, T, tid PROF_PASS);
        ++n_aligns;
        PROF_MARK(13);

        if (P.det_mode) {
            // frame_detector_cc: report and re-phase the hop grid half a buffer later
            if (tid == 0) {
                unsigned slot = atomicAdd(P.n_out, 1u);
                if (slot < P.max_out) {
                    Detection d;
                    d.F = F; d.stream = io.stream; d.seq = st.seq;
                    d.tau = sh.tau; d.gamma = sh.gamma; d.dphi = sh.dphi; d.phi = sh.phi; d.rxy = st.rxy; d.pad = st.det_idx;
                    P.detections[slot] = d;
                }
                st.seq++;
                st.mode = 0;
                st.wstart = F + 256;
            }
            wsync();
            PROF_MARK(9);
            continue;
        }

        const long long hdr_last = F + 616 - (long long)sh.tau_neg;    // sample that completes header symbol 308
        if (hdr_last + 1 > sv.end) {
            if (tid == 0) st.need_until = hdr_last + 1;
            wsync();
            break;
        }
        decode_header(sh, T, sv, F, tid, &pll_theta0, &pll_dtheta PROF_PASS);
        PROF_MARK(14);

        // header fields (every thread reads the same shared bytes)
        const unsigned char *hd = sh.hdec;
        int hv = sh.hv;
        unsigned plen = ((unsigned)hd[15] << 8) | hd[16];
        unsigned ms = hd[17], check = (hd[18] >> 5) & 7u, fec0 = hd[18] & 0x1fu, fec1 = hd[19] & 0x1fu;
        if (hv) {
            if (hd[14] != 102) hv = 0;
            else if (!modem_supported_hd(ms)) hv = 0;
            else if (check == 0 || check >= 7) hv = 0;
            else if (!fec_supported_hd(fec0) || !fec_supported_hd(fec1)) hv = 0;
        }
        unsigned n_sym = 0, k0 = 0, n0 = 0, n1 = 0;
        long long last = hdr_last;
        if (hv) {
            k0 = plen + crc_len_hd(check);
            n0 = fec_enc_len_hd(fec0, k0);
            n1 = fec_enc_len_hd(fec1, n0);
            unsigned bps = modem_bps_hd(ms);
            n_sym = (8u * n1 + bps - 1u) / bps;
            last = F + 2ll * (308ll + (long long)n_sym) - (long long)sh.tau_neg;
            if (last + 1 > sv.end) {
                // payload not complete yet: either wait for more samples or drop an oversized frame
                long long need = last + 1 - (F > sv.G ? F : sv.G);
                if (need > (long long)P.carry_cap) {
                    // reported as a frame with flags bit 0 (header fields filled, no payload): the caller sees the drop
                    if (tid == 0) {
                        unsigned slot = atomicAdd(P.n_out, 1u);
                        if (slot < P.max_out) {
                            FrameDesc &d = P.frames[slot];
                            d = FrameDesc{};
                            d.F = F; d.G = sv.G;
                            d.stream = io.stream; d.seq = st.seq; d.io_index = s_item; d.flags = 1u; d.det_idx = st.det_idx;
                            d.tau = sh.tau; d.gamma = sh.gamma; d.dphi = sh.dphi; d.phi = sh.phi; d.rxy = st.rxy;
                            d.header_valid = 1; d.payload_len = plen; d.ms = ms; d.bps = modem_bps_hd(ms);
                            d.check = check; d.fec0 = fec0; d.fec1 = fec1;
                            d.rssi = __fmul_rn(20.0f, log10f(sh.gamma));
                            d.cfo = nco_get_frequency_dev(sh.dtheta);
                            for (int i = 0; i < 20; ++i) d.header[i] = hd[i];
                        }
                        st.seq++;
                        st.dropped++;
                        st.mode = 0; st.G = F + 512; st.wstart = st.G - 256;
                    }
                    wsync();
                    sv.G = st.G;
                    continue;
                }
                if (tid == 0) st.need_until = last + 1;
                wsync();
                break;
            }
        }
        if (tid == 0) {
            unsigned slot = atomicAdd(P.n_out, 1u);
            if (slot < P.max_out) {
                FrameDesc &d = P.frames[slot];
                d.F = F; d.G = sv.G;
                d.sym_off = 0; d.buf_off = 0; d.pay_off = 0; d.dec_off = 0;
                d.stream = io.stream; d.seq = st.seq; d.io_index = s_item; d.flags = 0; d.det_idx = st.det_idx;
                d.tau = sh.tau; d.gamma = sh.gamma; d.dphi = sh.dphi; d.phi = sh.phi; d.rxy = st.rxy;
                d.mf_scale = sh.mf_scale;
                d.mix_theta0 = sh.theta0; d.mix_dtheta = sh.dtheta;
                d.pll_theta0 = pll_theta0; d.pll_dtheta = pll_dtheta;
                d.pfb_index = sh.pfb; d.tau_neg = sh.tau_neg;
                d.header_valid = hv; d.payload_valid = 0;
                d.payload_len = hv ? plen : 0; d.ms = hv ? ms : 0; d.bps = hv ? modem_bps_hd(ms) : 0;
                d.check = hv ? check : 0; d.fec0 = hv ? fec0 : 0; d.fec1 = hv ? fec1 : 0;
                d.n_sym = n_sym; d.k0 = k0; d.n0 = n0; d.n1 = n1;
                d.buf_len = 0; d.ilv1_off = 0; d.ilv0_off = 0;
                d.evm = 0.0f; d.evm_acc = 0.0f;
                d.rssi = __fmul_rn(20.0f, log10f(sh.gamma));
                d.cfo = nco_get_frequency_dev(sh.dtheta);
                for (int i = 0; i < 20; ++i) d.header[i] = hd[i];
            }
            st.seq++;
            st.mode = 0;
            st.G = last + 1;
            st.wstart = st.G - 256;
        }
        wsync();
        sv.G = st.G;
        PROF_MARK(9);
    }
    PROF_MARK(7);
    if (tid == 0) {
        long long r = (st.mode == 0) ? st.wstart : st.F;
        if (r < st.G) r = st.G;
        if (r > sv.end) r = sv.end;
        if (r < sv.base) r = sv.base;
        st.resume = r;
        // (time slices) the stream goes back into the queue when this CTA stopped at its slice boundary only
        const bool again = P.slice_len && st.mode == 0 && st.wstart + 512 <= sv.end && st.wstart < s_saved_stop;
        st.stop_at = s_saved_stop;
        P.states[io.stream] = st;
        if (P.slice_len) {
            __threadfence();                                             // state and frame descriptors before the hand-over
            if (again) {
                const unsigned p_ = atomicAdd(P.queue + 1, 1u);
                if (p_ < P.queue_cap) st_release_u32(P.queue + 4 + p_, s_item);
                else { atomicAdd(&g_seek_stall[0], 1u << 24); atomicAdd(P.queue + 2, 1u); }            // cannot happen: see launch_seek
            } else atomicAdd(P.queue + 2, 1u);
        }
#ifdef LQB_SEEK_TRACE
        const unsigned slot_ = atomicAdd(&g_seek_trace_n, 1u) & 0xffffu;
        g_seek_trace[0][slot_] = trace_t0; g_seek_trace[1][slot_] = gtime_ns();
        g_seek_trace[2][slot_] = smid_of(); g_seek_trace[3][slot_] = ((unsigned long long)(n_windows - trace_w0) << 32) | io.stream;
#endif
    }
    if (!P.slice_len) break;
    wsync();                      // (st and s_item are rewritten by the next pass)
    }   // passes

    if (fused) {
        // tell the MMA warp to leave (the pipeline is empty here)
        if (tid == 0) sh.tc_cmd[buf_w] = 0;
        mbar_arrive(&sh.z_full[buf_w]);
    }
    if (tid == 0) {
        atomicAdd(P.n_out + 1, n_windows);
        atomicAdd(P.n_out + 2, n_aligns);
        atomicAdd(P.n_out + 3, n_exact);
        atomicAdd(P.n_out + 4, n_tc_tiles);
        atomicAdd(P.n_out + 5, n_bins);
#ifdef LQB_SEEK_PROF
        for (int i = 0; i < 24; ++i) atomicAdd(&g_seek_prof[i], (unsigned long long)prof_acc[i]);
#endif
    }
    }   // workers

    if (fused) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(sh.tmem_base), "r"(128));
    }
}

// copy the unconsumed tail [resume, end) of every fed stream into the other carry buffer
__global__ void k_carry(SeekParams P)
{
    const StreamIO io = P.io[blockIdx.x];
    StreamState *sp = P.states + io.stream;
    __shared__ StreamState st;
    if (threadIdx.x == 0) st = *sp;
    __syncthreads();
    const float2 *old = P.carry[st.carry_sel] + (size_t)io.stream * P.carry_cap;
    float2 *dst = P.carry[st.carry_sel ^ 1u] + (size_t)io.stream * P.carry_cap;
    const long long end = st.base + (long long)st.carry_len + (long long)io.n_in;
    long long n = end - st.resume;
    if (n > (long long)P.carry_cap) n = P.carry_cap;        // cannot happen (oversized frames are dropped)
    const long long first = end - n;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        long long a = first + i - st.base;
        dst[i] = (a < (long long)st.carry_len) ? old[a] : io.in[a - st.carry_len];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        sp->base = first;
        sp->carry_len = (unsigned)n;
        sp->carry_sel = st.carry_sel ^ 1u;
    }
}

__global__ void k_seek_queue_init(unsigned *q, unsigned n_io, unsigned cap)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) q[4 + i] = i < n_io ? i : kQueueEmpty;
    if (i == 0) { q[0] = 0; q[1] = n_io; q[2] = 0; q[3] = 0; }
}

void launch_seek(const SeekParams &P, unsigned n_io, cudaStream_t s)
{
    static std::atomic<unsigned long long> attr_seen{ 0 };
    const int dyn = tc::kBBytes + 256;       // three CTAs per SM: 3 x (34 KB static + 35 KB B) fits 227 KB
    if (first_launch_on_this_device(attr_seen)) cudaFuncSetAttribute(k_seek, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (P.slice_len) {
        // one CTA per slot of the GPU (three per SM); P.grid bounds the tickets they can draw: the slices the fed
        // streams can take plus one failed draw per CTA
        const unsigned ctas = std::min<unsigned>(3u * (unsigned)sm_count_of_this_device(), P.grid);
        SeekParams Q = P;
        Q.queue_cap = P.grid + ctas;
        k_seek_queue_init<<<(Q.queue_cap + 255) / 256, 256, 0, s>>>(Q.queue, n_io, Q.queue_cap);
        k_seek<<<ctas, kCtaThreads, P.coarse == 2 ? dyn : 0, s>>>(Q);
    } else {
        k_seek<<<n_io, kCtaThreads, P.coarse == 2 ? dyn : 0, s>>>(P);
    }
}
void launch_carry(const SeekParams &P, unsigned n_io, cudaStream_t s) { k_carry<<<n_io, 256, 0, s>>>(P); }

// host: B operand in the kernel's shared-memory layout (e4m3 bytes, values x 32).  s: 156 template samples; column
// b holds the real part of C[., b], column 56 + b the imaginary part, for component c of x (0: re, 1: im).
// Layout: [component][MMA j][K chunk q][column nn][16 k-values], k = 32 j + 16 q + e.
// Returns max_b || t_b - t_q,b || / || s || of the bytes it produced (the B term of the pre-filter's error bound).
float build_coarse_bmat(const float *s_re, const float *s_im, int range, std::vector<unsigned char> &out)
{
    using namespace tc;
    out.assign(kBBytes, 0);
    auto quant = [](double v, unsigned char &byte) {
        byte = (unsigned char)__nv_cvt_float_to_fp8((float)(v * (double)kBScale), __NV_SATFINITE, __NV_E4M3);
        const __half_raw hr = __nv_cvt_fp8_to_halfraw(byte, __NV_E4M3);
        return (double)__half2float(__half(hr)) / (double)kBScale;
    };
    double s2 = 0.0, worst = 0.0;
    for (int n = 0; n < 156; ++n) s2 += (double)s_re[n] * s_re[n] + (double)s_im[n] * s_im[n];
    for (int b = 0; b < kNBins; ++b) {
        double err2 = 0.0;
        for (int n = 0; n < kKPad; ++n) {
            if (n >= 156) continue;                          // padding rows stay zero
            const double ph = 2.0 * 3.14159265358979323846 * (double)(b - range) * (double)n / 512.0;
            const double tr = s_re[n] * cos(ph) - s_im[n] * sin(ph), ti = s_re[n] * sin(ph) + s_im[n] * cos(ph);
            const int j = n >> 5, q = (n >> 4) & 1, e = n & 15;
            unsigned char byte = 0;
            // C = sum (xr + j xi)(tr - j ti):  re = xr tr + xi ti,  im = xi tr - xr ti
            for (int c = 0; c < 2; ++c)
                for (int part = 0; part < 2; ++part) {
                    const double val = (part == 0) ? (c == 0 ? tr : ti) : (c == 0 ? -ti : tr);
                    const double back = quant(val, byte);
                    out[(size_t)((c * kMmaPerComp + j) * 2 + q) * kBChunkBytes + (size_t)(kImCol0 * part + b) * 16 + e] = byte;
                    if (c == 0) err2 += (val - back) * (val - back);       // tr and ti once each (c = 1 holds the same two numbers)
                }
        }
        worst = std::max(worst, err2);
    }
    return (float)(std::sqrt(worst / s2) * 1.0001);
}

// [0] draws that gave up since the last reset, then the first one's ticket, queue head / tail / finished count, streams, slice bound, CTA
extern "C" int lqb_dbg_seek_stall(unsigned *out8, int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out8, g_seek_stall, sizeof(unsigned) * 8) != cudaSuccess) return -5;
    if (reset) { unsigned z[8] = {}; cudaMemcpyToSymbol(g_seek_stall, z, sizeof z); }
    return 0;
}

#ifdef LQB_SEEK_TRACE
extern "C" int lqb_dbg_seek_trace(unsigned long long *out /* [4][cap] */, unsigned cap, unsigned *n, int reset)
{
    cudaDeviceSynchronize();
    unsigned cnt = 0;
    if (cudaMemcpyFromSymbol(&cnt, g_seek_trace_n, sizeof cnt) != cudaSuccess) return -5;
    if (cnt > 65536u) cnt = 65536u;
    if (cnt > cap) cnt = cap;
    for (int k = 0; k < 4; ++k)
        if (out && cudaMemcpyFromSymbol(out + (size_t)k * cap, g_seek_trace, sizeof(unsigned long long) * cnt, sizeof(unsigned long long) * 65536 * k) != cudaSuccess) return -5;
    if (n) *n = cnt;
    if (reset) { unsigned z = 0; cudaMemcpyToSymbol(g_seek_trace_n, &z, sizeof z); }
    return 0;
}
#endif

#ifdef LQB_SEEK_PROF
extern "C" int lqb_dbg_seek_prof(unsigned long long *out24, int reset)
{
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out24, g_seek_prof, sizeof(unsigned long long) * 24) != cudaSuccess) return -5;
    if (reset) { unsigned long long z[24] = {}; cudaMemcpyToSymbol(g_seek_prof, z, sizeof z); }
    return 0;
}
#endif

}  // namespace lqb
