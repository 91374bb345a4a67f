// lqb_rx_soft.cu -- soft decisions for the payload (opt-in, LQB_RX_SOFT).
//
// liquid-dsp's flexframesync decodes hard decisions by default and that is what the blocks of the reference use
// (lib/flex_rx_impl.cc:49: flexframesync_create only); this path is an extension for callers that want the ~2 dB a
// soft-input Viterbi decoder gives.  Its definition is OURS (oracle/lqo.h: lqo_modem_demodulate_soft,
// lqo_qpm_decode_soft) and the kernels below reproduce the oracle's arithmetic operation by operation:
//   k_soft_demod : per payload symbol (the PLL-corrected constellation point kept by k_pll_emit) and per bit of the
//                  symbol, soft = clamp(trunc(128 + G (d0 - d1)), 0, 255) with d0 / d1 the smallest squared distance to
//                  a constellation point whose bit is 0 / 1 and G = 64 / dmin^2
//   k_soft_deint : the byte interleaver as a permutation of coded bits (host-built, csrc/lqb_tables.cpp: ilv_bit_perm)
// The soft-input Viterbi decoder is k_viterbi<true> (csrc/lqb_rx_fec.cu).
#include "lqb_dev.cuh"
#include "lqb_kernels.h"

namespace lqb {

namespace {

constexpr int kSoftThreads = 128;
constexpr unsigned kSoftSymsPerCta = 1024;

__global__ void __launch_bounds__(kSoftThreads)
k_soft_demod(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list)
{
    __shared__ float2 map[256];
    __shared__ float red[kSoftThreads / 32];
    const unsigned fi = list[blockIdx.x];
    const FrameDesc &d = P.frames[fi];
    const SoftDesc sd = P.soft[fi];
    const unsigned bps = d.bps, M = 1u << bps;
    const unsigned s0 = blockIdx.y * kSoftSymsPerCta;
    if (s0 >= d.n_sym) return;
    const int tid = threadIdx.x;
    for (unsigned i = tid; i < M; i += kSoftThreads) map[i] = modem_point(P.tables->psk_map, d.ms, bps, i);
    __syncthreads();
    // G = 64 / (smallest squared distance between two points): minimum over all pairs (order does not matter)
    float dmin2 = 3.0e38f;
    for (unsigned idx = tid; idx < M * M; idx += kSoftThreads) {
        const unsigned a = idx / M, b = idx - a * M;
        if (b > a) {
            const float dr = __fsub_rn(map[a].x, map[b].x), di = __fsub_rn(map[a].y, map[b].y);
            dmin2 = fminf(dmin2, __fmaf_rn(di, di, __fmul_rn(dr, dr)));
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) dmin2 = fminf(dmin2, __shfl_xor_sync(0xffffffffu, dmin2, m));
    if ((tid & 31) == 0) red[tid >> 5] = dmin2;
    __syncthreads();
    dmin2 = fminf(fminf(red[0], red[1]), fminf(red[2], red[3]));
    const float G = __fdiv_rn(64.0f, dmin2);
    const float2 *syms = P.syms + d.sym_off;
    unsigned char *out = P.soft_raw + sd.raw_off;
    const unsigned s1 = min(d.n_sym, s0 + kSoftSymsPerCta);
    for (unsigned i = s0 + tid; i < s1; i += kSoftThreads) {
        const float2 x = syms[i];
        float d0[8], d1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { d0[k] = 3.0e38f; d1[k] = 3.0e38f; }
        for (unsigned s = 0; s < M; ++s) {
            const float dr = __fsub_rn(x.x, map[s].x), di = __fsub_rn(x.y, map[s].y);
            const float dist = __fmaf_rn(di, di, __fmul_rn(dr, dr));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < (int)bps) {
                    const bool one = (s >> (bps - 1 - k)) & 1u;
                    d1[k] = one ? fminf(d1[k], dist) : d1[k];
                    d0[k] = one ? d0[k] : fminf(d0[k], dist);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < (int)bps) {
                const float t = __fmaf_rn(G, __fsub_rn(d0[k], d1[k]), 128.0f);
                out[(size_t)i * bps + k] = (unsigned char)(t <= 0.0f ? 0 : t >= 255.0f ? 255 : (int)t);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_soft_deint(PayloadParams P, const unsigned *__restrict__ list, unsigned n_list)
{
    const unsigned fi = list[blockIdx.x];
    const FrameDesc &d = P.frames[fi];
    const SoftDesc sd = P.soft[fi];
    const unsigned nbits = 8u * (sd.stage == 1 ? d.n1 : d.n0);
    const unsigned char *raw = P.soft_raw + sd.raw_off;
    unsigned char *out = P.soft_d + sd.d_off;
    const unsigned *perm = P.bitperm + sd.perm_off;
    for (unsigned i = blockIdx.y * 4096u + threadIdx.x; i < min(nbits, (blockIdx.y + 1u) * 4096u); i += 256u) out[i] = raw[perm[i]];
}

}  // namespace

void launch_soft_demod(const PayloadParams &P, const unsigned *list, unsigned n, unsigned max_syms, unsigned max_bits, cudaStream_t s)
{
    if (!n) return;
    k_soft_demod<<<dim3(n, (max_syms + kSoftSymsPerCta - 1) / kSoftSymsPerCta), kSoftThreads, 0, s>>>(P, list, n);
    k_soft_deint<<<dim3(n, (max_bits + 4095u) / 4096u), 256, 0, s>>>(P, list, n);
}

}  // namespace lqb
