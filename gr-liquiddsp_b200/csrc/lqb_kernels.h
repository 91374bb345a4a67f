// lqb_kernels.h -- launch interface between the C-ABI layer (lqb_api.cu) and the kernels.
#pragma once
#include "lqb_dev.cuh"
#include <atomic>
#include <vector>

namespace lqb {

// Function attributes and device properties belong to a DEVICE, not to the process: a handle on a second GPU of the
// same process needs them again.  One bit per device ordinal; the first launch on a device (any thread) sees true.
// Doing the set-up twice in a race is harmless, skipping it is not.
inline bool first_launch_on_this_device(std::atomic<unsigned long long> &seen)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen.load(std::memory_order_relaxed) & bit) return false;
    seen.fetch_or(bit, std::memory_order_relaxed);
    return true;
}
inline int sm_count_of_this_device()
{
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = cache[dev & 63].load(std::memory_order_relaxed);
    if (v <= 0) {
        v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cache[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

struct SeekParams {
    const DevTables *tables;
    StreamState     *states;
    const StreamIO  *io;
    float2          *carry[2];      // [n_streams][carry_cap] each
    unsigned         carry_cap;
    int              det_mode;      // 0: flexframesync discovery, 1: bare qdetector
    FrameDesc       *frames;        // det_mode 0
    Detection       *detections;    // det_mode 1
    unsigned        *n_out;
    unsigned         max_out;
    StreamView      *views;         // [n_io] out: the sample view every fed stream was searched under (payload kernels read through it)
    // tensor-core pre-filter fused into k_seek
    int              coarse;        // 0: off (every window takes the exact 50-FFT evaluation), 2: on
    const void      *bmat;          // B operand in shared-memory layout (build_coarse_bmat)
    float            b_err;         // max_b ||t_b - e4m3(t_b)|| / ||s|| of those bytes (error bound of the pre-filter)
    // time slices (0: one CTA per stream for the whole call): see k_seek
    unsigned         slice_len = 0; // samples a CTA searches before it hands the stream back (multiple of 256)
    unsigned         n_io = 0;      // fed streams of this launch
    unsigned         grid = 0;      // bound on the slices of this launch: sum over streams of ceil(available / slice_len) + 1
    unsigned        *queue = nullptr; // [4 + grid + CTAs launched] device words
    unsigned         queue_cap = 0; // entries behind the 4 header words (launch_seek fills it in)
};

// CTA shapes of k_mf / k_pll_emit.  64-thread / 1024-symbol variants (14 KB / 13 KB of shared memory) were built to fit
// beside resident search CTAs; measured A/B on the pipelined bench they are no faster (profiles/r01_notes.md v21), and
// the larger shapes are a little faster on their own.
#ifndef LQB_MF_THREADS
#define LQB_MF_THREADS 128
#endif
#ifndef LQB_EMIT_SPAN
#define LQB_EMIT_SPAN 4096
#endif
constexpr unsigned kMfTileSyms = 8 * LQB_MF_THREADS;   // payload symbols per matched-filter tile (host plan and k_mf agree on it)

// work list entry for kernels that run per FEC stage
struct StageItem { unsigned frame; unsigned pad; };

// soft-decision frames (LQB_RX_SOFT): where a frame's soft bits live and which stage consumes them
struct SoftDesc {
    unsigned long long raw_off;     // soft bytes in transmission order, n_sym * bps of them
    unsigned long long d_off;       // deinterleaved soft bytes, 8 * enc_len of the soft stage
    unsigned perm_off;              // bit permutation of that length in the bitperm arena
    int      stage;                 // 1: fec1, 0: fec0 (fec1 is "none"), -1: hard decisions only
};

struct PayloadParams {
    const DevTables   *tables;
    const StreamView  *views;       // [n_io] written by k_seek: carry / input pointers and bounds as of this call
    FrameDesc         *frames;
    unsigned           n_frames;
    // matched filter tiling: tile_start[f] = first tile of frame f (exclusive prefix), n_tiles total
    const unsigned    *tile_start;
    unsigned           n_tiles;
    uint4             *tile_rec;    // [3 * n_tiles] per-tile matched-filter records (MfTileRec, filled on the device)
    float2            *syms;        // symbol arena
    unsigned char     *bufA, *bufB; // byte arenas
    unsigned char     *payload;     // payload output pool
    const unsigned    *ilv_maps;    // interleaver map arena
    unsigned long long *decisions;  // Viterbi decision arena
    void              *pll_ckpt;    // PLL checkpoints, 16 bytes per 32 payload symbols (FrameDesc::ck_off)
    // soft-decision path (null without LQB_RX_SOFT)
    const SoftDesc    *soft;        // [n_frames]
    unsigned char     *soft_raw, *soft_d;
    const unsigned    *bitperm;
};

void launch_seek(const SeekParams &P, unsigned n_io, cudaStream_t s);
// host: B operand of the pre-filter (template x CFO rotations) in the kernel's shared-memory layout
float build_coarse_bmat(const float *s_re, const float *s_im, int range, std::vector<unsigned char> &out);
void launch_carry(const SeekParams &P, unsigned n_io, cudaStream_t s);

// device <-> pinned-host copy of small control data by a kernel (never queues behind bulk DMA copies)
void launch_copy(void *dst, const void *src, size_t bytes, cudaStream_t s);
void launch_mf(const PayloadParams &P, cudaStream_t s);
constexpr unsigned kEmitSpan = LQB_EMIT_SPAN;     // payload symbols per CTA of the PLL emit pass (host plan and k_pll_emit agree on it)
// list: frames grouped by modulation; span_start: exclusive prefix (n + 1) of kEmitSpan-symbol spans over that list
void launch_pll(const PayloadParams &P, const unsigned *list, const unsigned *span_start, unsigned n, unsigned n_spans, cudaStream_t s);
// stage = 1: bufA(n1) -> bufB(n0) with fec1;  stage = 0: bufB(n0) -> bufA(k0) with fec0
void launch_deinterleave(const PayloadParams &P, const unsigned *list, unsigned n, int stage, cudaStream_t s);
void launch_blockfec(const PayloadParams &P, const unsigned *list, unsigned n, int stage, cudaStream_t s);
// punct: some frame in the list uses a punctured code (K = 7 only: selects the generic symbol fetch)
void launch_viterbi(const PayloadParams &P, const unsigned *list, unsigned n, int stage, unsigned K, bool punct, cudaStream_t s);
void launch_rs(const PayloadParams &P, const unsigned *blocks /* pairs (frame, block) */, unsigned n_blocks, int stage, cudaStream_t s);
void launch_crc(const PayloadParams &P, const unsigned *list, unsigned n, cudaStream_t s);
// soft decisions for the listed frames: demodulate the stored constellation points to soft bytes, deinterleave them
// through the bit permutation, then (launch_viterbi_soft, per stage) the soft-input Viterbi decoder
void launch_soft_demod(const PayloadParams &P, const unsigned *list, unsigned n, unsigned max_syms, unsigned max_bits, cudaStream_t s);
void launch_viterbi_soft(const PayloadParams &P, const unsigned *list, unsigned n, int stage, unsigned K, bool punct, cudaStream_t s);

// debug / unit-test hooks (single launches on tiny inputs)
void launch_dbg_fft512(const DevTables *T, const float2 *in, float2 *out, int dir, cudaStream_t s);

}  // namespace lqb
