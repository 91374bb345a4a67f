// lqb_rx_coarse.cu -- tensor-core preamble pre-filter (tcgen05 / TMEM).
//
// The exact detector (lqb_rx_seek.cu) spends 50 FFT-512 per 256-sample hop whether or not a
// preamble is there.  This kernel computes, for EVERY lag l of every stream and all 49 CFO bins,
// the correlation  C[l,b] = sum_n x[l+n] conj(s[n]) exp(-j 2 pi (b-24) n / 512)  as one dense
// [lags x 2*156] x [2*156 x 2*49] contraction on the 5th-generation tensor cores, in fp16 with
// fp32 accumulation, and keeps max_b |C|^2 per 8 lags plus the sample energy per 8 samples.
// The exact kernel then proves most windows cannot trigger (upper bound on rxy below threshold,
// with a margin that covers the fp16 rounding bound) and runs its FFTs only on the rest, so
// decisions stay identical to the specification.
//
// The A operand is never materialised: a window matrix A[l][n] = x[l+n] is Hankel, and a
// no-swizzle K-major UMMA shared-memory descriptor addresses "8 rows x 16 bytes" core matrices
// at arbitrary 16-byte strides.  Storing the samples 8-fold interleaved, Z[8 m + e] = x[m + e],
// makes row r of the core matrix at &Z[8 (l0 + n0)] equal to x[l0 + r + n0 .. +7]; consecutive
// row groups and consecutive K chunks are both 128 bytes further on, so one 9 KB buffer serves
// all 20 MMAs of a 128-lag tile.
#include "lqb_dev.cuh"
#include "lqb_kernels.h"
#include "lqb_tc.cuh"
#include <cmath>
#include <vector>

namespace lqb {

namespace {

using namespace tc;

struct CoarseShared {
    alignas(128) unsigned char B[kBBytes];
    alignas(128) unsigned char Z[2][2][kZBytes];      // [buffer][component]
    __half xs[2][kSampNeed + 8];                      // staged samples, re / im planes
    alignas(8) uint64_t bar[2];
    uint32_t tmem_base;
    float scale2[2];                                  // 2^(2 e) of the tile in each buffer
    float wmax[2][4];
};

struct CoarseTile { unsigned stream, io_index, lag0; unsigned long long out_off; };

// tile -> (stream, first lag): binary search in the per-stream tile prefix (n_io + 1 entries)
__device__ __forceinline__ CoarseTile locate_tile(const CoarseParams &P, unsigned tile)
{
    unsigned lo = 0, hi = P.n_io;
    while (hi - lo > 1) {
        const unsigned mid = (lo + hi) >> 1;
        if (P.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    CoarseTile ct;
    ct.io_index = lo;
    ct.stream = P.io[lo].stream;
    ct.lag0 = (tile - P.tile_prefix[lo]) * (unsigned)kTileLags;
    ct.out_off = (unsigned long long)tile * 16ull;
    return ct;
}

}  // namespace

// work item: 128 consecutive lags of one stream
__global__ void __launch_bounds__(kCoarseThreads, 2)
k_coarse(CoarseParams P)
{
    extern __shared__ unsigned char smem_raw[];
    CoarseShared &sh = *reinterpret_cast<CoarseShared *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: B matrix, barriers, tensor memory
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.bmat);
        uint4 *dst = reinterpret_cast<uint4 *>(sh.B);
        for (int i = tid; i < kBBytes / 16; i += kCoarseThreads) dst[i] = src[i];
    }
    if (tid == 0) { mbar_init(&sh.bar[0], 1); mbar_init(&sh.bar[1], 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sh.tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sh.tmem_base;
    // instruction descriptor: D = F32 (bit 4), A = B = F16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileLags >> 4) << 24);
    const uint32_t b_addr = smem_u32(sh.B);

    const unsigned n_tiles = P.n_tiles;
    unsigned phase[2] = { 0u, 0u };

    // builds Z for tile `tile` into buffer `buf`; also writes the per-8-sample energies of its 128 lags
    auto build = [&](unsigned tile, int buf) {
        const CoarseTile ct = locate_tile(P, tile);
        const StreamState &st = P.states[ct.stream];
        const StreamIO &io = P.io[ct.io_index];
        StreamView sv;
        sv.carry = P.carry[st.carry_sel] + (size_t)ct.stream * P.carry_cap;
        sv.in = io.in; sv.base = st.base; sv.carry_len = st.carry_len;
        sv.end = st.base + (long long)st.carry_len + (long long)io.n_in;
        sv.G = -(1ll << 62);                                   // no zero region: raw samples
        const long long l0 = st.base + (long long)ct.lag0;
        // stage: three samples per thread (296 needed), track max magnitude for the tile scale
        float2 v[3];
        float mx = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = tid + kCoarseThreads * k;
            v[k] = (i < kSampNeed) ? sv.at(l0 + i) : make_float2(0.0f, 0.0f);
            mx = fmaxf(mx, fmaxf(fabsf(v[k].x), fabsf(v[k].y)));
        }
        // energy of the tile's own 128 samples in blocks of 8 (exact fp32 inputs)
        {
            float e = __fmaf_rn(v[0].y, v[0].y, __fmul_rn(v[0].x, v[0].x));
            e += __shfl_xor_sync(0xffffffffu, e, 1);
            e += __shfl_xor_sync(0xffffffffu, e, 2);
            e += __shfl_xor_sync(0xffffffffu, e, 4);
            if ((lane & 7) == 0) P.e8[ct.out_off + (tid >> 3)] = e;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0) sh.wmax[buf][warp] = mx;
        __syncthreads();
        mx = fmaxf(fmaxf(sh.wmax[buf][0], sh.wmax[buf][1]), fmaxf(sh.wmax[buf][2], sh.wmax[buf][3]));
        // power-of-two scale bringing the largest component into [1, 2): exact, no rounding
        int ex = 0;
        if (mx > 0.0f) ex = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;
        ex = max(-100, min(100, ex));
        const float sc = __uint_as_float((uint32_t)(127 - ex) << 23);
        if (tid == 0) sh.scale2[buf] = __uint_as_float((uint32_t)(127 + 2 * ex) << 23);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = tid + kCoarseThreads * k;
            if (i < kSampNeed) {
                sh.xs[0][i] = __float2half_rn(v[k].x * sc);
                sh.xs[1][i] = __float2half_rn(v[k].y * sc);
            }
        }
        __syncthreads();
        // Z[c][8 m + e] = xs[c][m + e], m < 288: each thread writes rows tid, tid+128, tid+256
        for (int m = tid; m < kZRows; m += kCoarseThreads) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const __half *s = sh.xs[c] + m;
                __half2 h0 = __halves2half2(s[0], s[1]), h1 = __halves2half2(s[2], s[3]);
                __half2 h2 = __halves2half2(s[4], s[5]), h3 = __halves2half2(s[6], s[7]);
                uint4 w;
                w.x = *reinterpret_cast<uint32_t *>(&h0); w.y = *reinterpret_cast<uint32_t *>(&h1);
                w.z = *reinterpret_cast<uint32_t *>(&h2); w.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(sh.Z[buf][c] + 16 * m) = w;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
    };

    auto issue = [&](int buf) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem + (uint32_t)buf * 128u;
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t za = smem_u32(sh.Z[buf][c]);
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const uint64_t da = make_desc(za + 256u * j, 128u, 128u);
                    const uint64_t db = make_desc(b_addr + (uint32_t)((c * 10 + j) * 2) * kBChunkBytes, kBChunkBytes, 128u);
                    mma_f16(d, da, db, idesc, acc);
                    acc = 1;
                }
            }
            mma_commit(&sh.bar[buf]);
        }
    };

    auto epilogue = [&](unsigned tile, int buf) {
        mbar_wait(&sh.bar[buf], phase[buf]);
        phase[buf] ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const CoarseTile ct = locate_tile(P, tile);
        const uint32_t taddr = tmem + (uint32_t)buf * 128u + ((uint32_t)(warp * 32) << 16);
        float best = 0.0f;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            uint32_t r[16];
            tmem_ld16(taddr + 16u * q, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                if (16 * q + k < 2 * kNBins) {
                    const float re = __uint_as_float(r[k]), im = __uint_as_float(r[k + 1]);
                    best = fmaxf(best, fmaf(im, im, re * re));
                }
            }
        }
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 4));
        if ((lane & 7) == 0) P.m8[ct.out_off + (tid >> 3)] = best * sh.scale2[buf];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    // ---- software pipeline over this CTA's tiles: build(i+1) and epilogue(i) overlap MMA(i+1)
    unsigned first = blockIdx.x, i = 0;
    if (first < n_tiles) { build(first, 0); issue(0); }
    for (unsigned tile = first; tile < n_tiles; tile += gridDim.x, ++i) {
        const int buf = (int)(i & 1u);
        const unsigned next = tile + gridDim.x;
        if (next < n_tiles) { build(next, buf ^ 1); issue(buf ^ 1); }
        epilogue(tile, buf);
        __syncthreads();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

void launch_coarse(const CoarseParams &P, cudaStream_t s)
{
    if (!P.n_tiles) return;
    static bool attr_set = false;
    const int smem = (int)sizeof(CoarseShared) + 128;
    if (!attr_set) { cudaFuncSetAttribute(k_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
    const unsigned grid = P.n_tiles < 148u * 2u ? P.n_tiles : 148u * 2u;
    k_coarse<<<grid, kCoarseThreads, smem, s>>>(P);
}

// host: B operand in the kernel's shared-memory layout (fp16).  s: 156 template samples.
// chunk (c, j, q) holds k = 8 consecutive template taps n = 16 j + 8 q + e of sample component c
// (0: real part of x multiplies it, 1: imaginary part), for all 112 output columns.
void build_coarse_bmat(const float *s_re, const float *s_im, int range, std::vector<unsigned short> &out)
{
    out.assign(kBBytes / 2, 0);
    for (int c = 0; c < 2; ++c)
        for (int j = 0; j < 10; ++j)
            for (int q = 0; q < 2; ++q) {
                const size_t chunk = (size_t)((c * 10 + j) * 2 + q) * (kBChunkBytes / 2);
                for (int nn = 0; nn < 2 * kNBins; ++nn)
                    for (int e = 0; e < 8; ++e) {
                        const int n = 16 * j + 8 * q + e, b = nn >> 1, part = nn & 1;
                        double val = 0.0;
                        if (n < 156) {
                            const double ph = 2.0 * 3.14159265358979323846 * (double)(b - range) * (double)n / 512.0;
                            const double tr = s_re[n] * cos(ph) - s_im[n] * sin(ph), ti = s_re[n] * sin(ph) + s_im[n] * cos(ph);
                            // C = sum (xr + j xi)(tr - j ti):  re = xr tr + xi ti,  im = xi tr - xr ti
                            val = (part == 0) ? (c == 0 ? tr : ti) : (c == 0 ? -ti : tr);
                        }
                        __half hv = __float2half_rn((float)val);
                        out[chunk + (size_t)nn * 8 + e] = *reinterpret_cast<unsigned short *>(&hv);
                    }
            }
}

}  // namespace lqb
