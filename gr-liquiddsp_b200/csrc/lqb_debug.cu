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

// one stream of n samples through the tensor-core pre-filter; returns m8 / e8 (n_tiles * 16 each)
extern "C" int lqb_dbg_coarse(const float *x_host, unsigned n, float beta, float *m8_host, float *e8_host)
{
    using namespace lqb;
    auto s = detector_template(beta);
    std::vector<float> sre(156), sim(156);
    for (int i = 0; i < 156; ++i) { sre[i] = s[i].re; sim[i] = s[i].im; }
    std::vector<unsigned short> bm;
    build_coarse_bmat(sre.data(), sim.data(), 24, bm);
    const unsigned n_tiles = (n + 127) / 128;
    StreamState st; std::memset(&st, 0, sizeof st);
    float2 *dx = nullptr, *dcarry = nullptr; void *dB = nullptr; StreamState *dst = nullptr; StreamIO *dio = nullptr;
    unsigned *dpre = nullptr; float *dm = nullptr, *de = nullptr;
    if (cudaMalloc(&dx, (size_t)n * sizeof(float2)) != cudaSuccess) return -19;
    cudaMalloc(&dcarry, 2048 * sizeof(float2)); cudaMalloc(&dB, bm.size() * 2); cudaMalloc(&dst, sizeof st);
    cudaMalloc(&dio, sizeof(StreamIO)); cudaMalloc(&dpre, 2 * sizeof(unsigned));
    cudaMalloc(&dm, (size_t)n_tiles * 16 * sizeof(float)); cudaMalloc(&de, (size_t)n_tiles * 16 * sizeof(float));
    StreamIO io; io.in = dx; io.n_in = n; io.stream = 0; io.pad = 0;
    unsigned pre[2] = { 0, n_tiles };
    cudaMemcpy(dx, x_host, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bm.data(), bm.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dst, &st, sizeof st, cudaMemcpyHostToDevice);
    cudaMemcpy(dio, &io, sizeof io, cudaMemcpyHostToDevice);
    cudaMemcpy(dpre, pre, sizeof pre, cudaMemcpyHostToDevice);
    CoarseParams P;
    P.states = dst; P.io = dio; P.carry[0] = dcarry; P.carry[1] = dcarry; P.carry_cap = 2048;
    P.tile_prefix = dpre; P.n_io = 1; P.n_tiles = n_tiles; P.bmat = dB; P.m8 = dm; P.e8 = de;
    launch_coarse(P, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaMemcpy(m8_host, dm, (size_t)n_tiles * 16 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(e8_host, de, (size_t)n_tiles * 16 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dcarry); cudaFree(dB); cudaFree(dst); cudaFree(dio); cudaFree(dpre); cudaFree(dm); cudaFree(de);
    if (e != cudaSuccess) { fprintf(stderr, "lqb_dbg_coarse: %s\n", cudaGetErrorString(e)); return -5; }
    return 0;
}
