// lqb_tc.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers and the tile geometry of the tensor-core
// preamble correlation (shared by lqb_rx_coarse.cu and the fused path in lqb_rx_seek.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace lqb {
namespace tc {

constexpr int kCoarseThreads = 128;
constexpr int kTileLags = 128;
constexpr int kNBins = 49;
constexpr int kN = 112;                 // 2 * 49 = 98 output columns, padded to a multiple of 16
constexpr int kKPad = 160;              // 156 template samples padded to 10 MMAs of K = 16
constexpr int kZRows = kTileLags + kKPad;       // 288 rows of 16 bytes per component
constexpr int kZBytes = kZRows * 16;            // 4608
constexpr int kBChunkBytes = kN * 16;           // 1792: one K-chunk (8 k-values) of B, all N rows
constexpr int kBBytes = 2 * 10 * 2 * kBChunkBytes;   // comp x mma x kchunk = 71680
constexpr int kSampNeed = kTileLags + kKPad + 8;     // 296 samples feed one tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 in [0,14), LBO >> 4 in [16,30),
    // SBO >> 4 in [32,46), version = 1 in [46,48), layout_type = 0 (no swizzle) in [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}


}  // namespace tc
}  // namespace lqb
