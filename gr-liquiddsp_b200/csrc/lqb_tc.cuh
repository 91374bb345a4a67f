// lqb_tc.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers and the tile geometry of the tensor-core
// preamble correlation (the pre-filter fused into k_seek, lqb_rx_seek.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

#ifndef LQB_MBAR_HINT_NS
#define LQB_MBAR_HINT_NS 0
#endif

namespace lqb {
namespace tc {

// Geometry of the tensor-core preamble correlation (fp8 e4m3 operands, fp32 accumulation in TMEM):
//   D[lag, 2 b + part] = sum_k A[lag, k] B[k, 2 b + part],  A = implicit Hankel matrix of the samples (one component
//   plane per pass), B = template x 49 CFO rotations.  One tcgen05.mma (kind::f8f6f4) covers K = 32 template samples.
constexpr int kTileLags = 128;          // M of one MMA / one accumulator tile
constexpr int kNBins = 49;
constexpr int kN = 112;                 // output columns: re of bin b at b, im at kImCol0 + b (49 + 7 zero columns each)
constexpr int kImCol0 = 56;
constexpr int kKPad = 160;              // 156 template samples padded to 5 MMAs of K = 32
constexpr int kMmaPerComp = kKPad / 32; // 5
constexpr int kBChunkBytes = kN * 16;   // 1792: one K-chunk (16 k-values, one byte each) of B, all N rows
constexpr int kBBytes = 2 * kMmaPerComp * 2 * kBChunkBytes;   // comp x mma x kchunk = 35840
constexpr float kBScale = 32.0f;        // B holds template values x 32 (|t| <= 1.4: well inside e4m3's normal range)
constexpr int kBScaleLog2 = 5;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 in [0,14), LBO >> 4 in [16,30),
    // SBO >> 4 in [32,46), version = 1 in [46,48), layout_type = 0 (no swizzle) in [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[smem] B[smem], e4m3 x e4m3 -> f32, K = 32 per instruction
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
// Waits for the phase with the given parity to complete.  A wait that cannot end (a tensor-core instruction that was
// rejected, a lost arrival) would hang the GPU: after ~2^26 polls (tens of seconds) the kernel traps instead, which
// the host sees as a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t sleep_ns)
{
    uint32_t done = 0, polls = 0;
    const uint32_t addr = smem_u32(bar);
    while (true) {
#if LQB_MBAR_HINT_NS
        // suspend-time hint: the thread sleeps in hardware until the phase completes or the hint elapses, instead of
        // returning after the (short) default limit and taking issue slots from the warps it is waiting for
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"((uint32_t)LQB_MBAR_HINT_NS) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
#endif
        if (done) break;
        // back off: a polling warp takes issue slots from the warps it is waiting for (ncu: a third of all executed
        // instructions were this loop before the sleep)
        if (sleep_ns) __nanosleep(sleep_ns);
        if (++polls > (1u << 24)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ float max3f(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));        // FMNMX3
    return r;
}

}  // namespace tc
}  // namespace lqb
