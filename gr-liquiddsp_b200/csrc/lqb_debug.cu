// lqb_debug.cu -- single-launch hooks that expose kernel building blocks to the unit tests
// (exported as lqb_dbg_*; not part of the public header).
#include "lqb_kernels.h"
#include "lqb_tables.h"
#include <vector>
#include <cstring>
#include <cstdio>

namespace lqb {

__global__ void k_dbg_fft512(const float2 *W, const float2 *in, float2 *out, int dir)
{
    __shared__ float2 sW[256];
    __shared__ float2 sWc[240];
    __shared__ float2 scr[544];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) sW[i] = W[i];
    __syncwarp();
    fft512_fill_compact(sWc, sW, lane, 32);
    __syncwarp();
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = in[fft512_in_index(lane, r)];
    if (dir > 0) fft512_warp<+1>(v, sW, sWc, scr, lane); else fft512_warp<-1>(v, sW, sWc, scr, lane);
#pragma unroll
    for (int r = 0; r < 16; ++r) out[fft512_out_index(lane, r)] = v[r];
}

// cycles of `reps` back-to-back transforms on one warp (a latency probe for the exact-evaluation path of k_seek)
__global__ void k_dbg_fft512_cycles(const float2 *W, const float2 *in, float2 *out, int reps, long long *cycles)
{
    __shared__ float2 sW[256];
    __shared__ float2 sWc[240];
    __shared__ float2 scr[4][544];
    __shared__ float2 buf[4][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sW[i] = W[i];
    for (int i = lane; i < 512; i += 32) buf[warp][i] = in[i];
    __syncthreads();
    fft512_fill_compact(sWc, sW, threadIdx.x, blockDim.x);
    __syncthreads();
    const long long t0 = clock64();
    float2 v[16];
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[warp][fft512_in_index(lane, r)];
        if (it & 1) fft512_warp<-1>(v, sW, sWc, scr[warp], lane); else fft512_warp<+1>(v, sW, sWc, scr[warp], lane);
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[warp][fft512_out_index(lane, r)] = v[r];
        __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    for (int i = lane; i < 512; i += 32) out[i] = buf[warp][i];
}

__global__ void k_dbg_pm(const float *y, const float *x, float *at, float *sn, float *cs, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    at[i] = pm_atan2f(y[i], x[i]);
    pm_sincosf(y[i], &sn[i], &cs[i]);
}

}  // namespace lqb

// pm_atan2f(y[i], x[i]) and pm_sincosf(y[i]) for n host values (the pinned arg / exp(j t) of the per-symbol loops)
extern "C" int lqb_dbg_pm(const float *y, const float *x, float *at, float *sn, float *cs, unsigned n)
{
    using namespace lqb;
    float *d[5] = {};
    for (auto &p : d) if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) return -19;
    cudaMemcpy(d[0], y, n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(d[1], x, n * sizeof(float), cudaMemcpyHostToDevice);
    k_dbg_pm<<<(n + 255) / 256, 256>>>(d[0], d[1], d[2], d[3], d[4], n);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(at, d[2], n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(sn, d[3], n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(cs, d[4], n * sizeof(float), cudaMemcpyDeviceToHost);
    for (auto &p : d) cudaFree(p);
    return e == cudaSuccess ? 0 : -5;
}

extern "C" int lqb_dbg_fft512_cycles(int warps, int ctas, int reps, long long *cycles_per_cta)
{
    using namespace lqb;
    auto W = twiddles(512);
    float2 *dW = nullptr, *din = nullptr, *dout = nullptr;
    long long *dc = nullptr;
    if (cudaMalloc(&dW, 256 * sizeof(float2)) != cudaSuccess) return -19;
    cudaMalloc(&din, 512 * sizeof(float2));
    cudaMalloc(&dout, 512 * sizeof(float2));
    cudaMalloc(&dc, ctas * sizeof(long long));
    cudaMemcpy(dW, W.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemset(din, 0, 512 * sizeof(float2));
    k_dbg_fft512_cycles<<<ctas, 32 * warps>>>(dW, din, dout, reps, dc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(cycles_per_cta, dc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dW); cudaFree(din); cudaFree(dout); cudaFree(dc);
    return e == cudaSuccess ? 0 : -5;
}

extern "C" int lqb_dbg_fft512(const float *in_host, float *out_host, int dir)
{
    using namespace lqb;
    auto W = twiddles(512);
    float2 *dW = nullptr, *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&dW, 256 * sizeof(float2)) != cudaSuccess) return -19;
    cudaMalloc(&din, 512 * sizeof(float2));
    cudaMalloc(&dout, 512 * sizeof(float2));
    cudaMemcpy(dW, W.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(din, in_host, 512 * sizeof(float2), cudaMemcpyHostToDevice);
    k_dbg_fft512<<<1, 32>>>(dW, din, dout, dir);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out_host, dout, 512 * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(dW); cudaFree(din); cudaFree(dout);
    return e == cudaSuccess ? 0 : -5;
}
