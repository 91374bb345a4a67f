// lqb_api.cu -- C-ABI (include/lqb200.h) for the receiver and the detector: handle
// lifetime, device arenas, the per-call plan (which frames go through which kernels) and
// result gathering.  All arithmetic lives in the kernels; this file only moves data and
// decides launch shapes.  There is no CPU fallback: without a usable device create() fails.
#include "../../include/lqb200.h"
#include "lqb_kernels.h"
#include "lqb_tables.h"

#include <algorithm>
#include <chrono>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace lqb;

namespace {

thread_local std::string g_err;

// LQB_TRACE=1: host-side timeline of every execute on stderr (ms since the call started)
struct Trace {
    bool on = getenv("LQB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0;
    void start() { if (on) t0 = std::chrono::steady_clock::now(); }
    void mark(const char *what, unsigned lane) const
    {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[lqb %8.3f ms] lane %u %s\n", ms, lane, what);
    }
};
Trace g_trace;
int fail(int code, const char *fmt, const char *a = "")
{
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a);
    g_err = buf;
    return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(LQB_ECUDA, "CUDA: %s", cudaGetErrorString(e_)); return LQB_ECUDA; } } while (0)
#define CUP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail(LQB_ECUDA, "CUDA: %s", cudaGetErrorString(e_)); return nullptr; } } while (0)

// grow-only device / pinned buffers
template <typename T> struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    int reserve(size_t n, bool keep = false, cudaStream_t s = 0)
    {
        if (n <= cap) return 0;
        size_t ncap = std::max(n + n / 4, cap + cap / 2);      // headroom: call-to-call variation must not reallocate
        T *q = nullptr;
        if (cudaMalloc(&q, ncap * sizeof(T)) != cudaSuccess) return fail(LQB_ENOMEM, "cudaMalloc failed");
        if (keep && p && cap) { cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s); cudaStreamSynchronize(s); }
        if (p) cudaFree(p);
        p = q; cap = ncap;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T> struct PinBuf {
    T *p = nullptr; size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        size_t ncap = std::max(n + n / 4, cap + cap / 2);      // pinned allocations are slow (0.4 s per GB): leave headroom
        T *q = nullptr;
        if (cudaMallocHost(&q, ncap * sizeof(T)) != cudaSuccess) return fail(LQB_ENOMEM, "cudaMallocHost failed");
        if (p) cudaFreeHost(p);
        p = q; cap = ncap;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

int build_tables(DevTables &T, float beta, float threshold, float dphi_max)
{
    std::memset(&T, 0, sizeof T);
    auto w512 = twiddles(512), w32 = twiddles(32);
    std::memcpy(T.W512, w512.data(), sizeof T.W512);
    std::memcpy(T.W32, w32.data(), sizeof T.W32);
    auto s = detector_template(beta);
    std::vector<cf> buf(512, cf{ 0.0f, 0.0f }), S(512);
    std::copy(s.begin(), s.end(), buf.begin());
    host_fft(buf.data(), S.data(), 512, +1);
    for (unsigned i = 0; i < 512; ++i) T.Sc[i] = make_float2(S[i].re, -S[i].im);
    float s2 = 0.0f;
    for (unsigned i = 0; i < kSLen; ++i) {
        T.sconj[i] = make_float2(s[i].re, -s[i].im);
        s2 += std::fmaf(s[i].im, s[i].im, s[i].re * s[i].re);
    }
    T.s2_sum = s2;
    T.threshold = threshold;
    T.range = (int)(dphi_max * 512.0f / (2.0f * 3.14159274f));
    std::memcpy(T.sintab, nco_sintab(), sizeof T.sintab);
    auto banks = pfb_banks(kRxBeta);
    std::memcpy(T.banks, banks.data(), sizeof T.banks);
    cf pil[15];
    header_pilots(pil);
    for (unsigned i = 0; i < 15; ++i) T.pilots_conj[i] = make_float2(pil[i].re, -pil[i].im);
    auto maps = psk_maps();
    std::memcpy(T.psk_map, maps.data(), sizeof T.psk_map);
    for (unsigned c = 3; c <= 6; ++c) crc_table(c, T.crc_tab[c]);
    auto m54 = ilv_maps(54), m27 = ilv_maps(27);
    for (unsigned p = 0; p < 4; ++p) {
        for (unsigned i = 0; i < 27; ++i) T.ilv54[p][i] = (uint16_t)m54[p * 27 + i];
        for (unsigned i = 0; i < 13; ++i) T.ilv27[p][i] = (uint16_t)m27[p * 13 + i];
    }
    // the header's deinterleaver (passes 3, 2, 1, 0 of masked swaps between byte pairs) moves bits without changing
    // them: run it once on bit labels and keep the resulting permutation, so the device can gather every byte at once
    auto compose = [](const std::vector<uint32_t> &maps, unsigned n, uint16_t *perm) {
        const unsigned n2 = n / 2, masks[4] = { 0xffu, 0x0fu, 0x55u, 0x33u };
        std::vector<uint16_t> lab(8 * n);
        for (unsigned i = 0; i < 8 * n; ++i) lab[i] = (uint16_t)i;
        for (int pass = 3; pass >= 0; --pass)
            for (unsigned i = 0; i < n2; ++i) {
                const unsigned j = maps[(size_t)pass * n2 + i];
                for (unsigned b = 0; b < 8; ++b)
                    if ((masks[pass] >> b) & 1u) std::swap(lab[8 * (2 * j + 1) + b], lab[8 * (2 * i) + b]);
            }
        for (unsigned i = 0; i < 8 * n; ++i) perm[i] = lab[i];
    };
    compose(m54, 54, T.hperm54);
    compose(m27, 27, T.hperm27);
    hamming_dec_tables(T.h84_dec, T.h74_dec);
    secded_cols(T.secded_col);
    uint8_t gen[33];
    gf256_tables(T.gf_exp, T.gf_log, gen);
    std::memcpy(T.rs_gen, gen, 33);
    for (unsigned v = 0; v < 256; ++v)
        for (unsigned i = 0; i < 32; ++i) {
            uint32_t w = 0;
            for (unsigned k = 1; k <= 4; ++k)
                if (v) w |= (uint32_t)T.gf_exp[(T.gf_log[v] + k * (i + 1)) % 255] << (8 * (k - 1));
            T.rs_syn[v][i] = w;
        }
    return 0;
}

// interleaved int16 (re, im) -> complex64, value / 32768 (exact): two samples per thread and step
__global__ void k_sc16_to_c32(const int2 *__restrict__ in, float4 *__restrict__ out, size_t n_pairs)
{
    const float k = 1.0f / 32768.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (size_t)gridDim.x * blockDim.x) {
        const int2 v = in[i];
        out[i] = make_float4((float)(short)(v.x & 0xffff) * k, (float)(short)(v.x >> 16) * k,
                             (float)(short)(v.y & 0xffff) * k, (float)(short)(v.y >> 16) * k);
    }
}

bool is_conv(unsigned fs) { return fs == FEC_CONV_V27 || fs == FEC_CONV_V29 || (fs >= FEC_CONV_V27P23 && fs <= FEC_CONV_V29P78); }
unsigned conv_K(unsigned fs) { return (fs == FEC_CONV_V29 || fs >= FEC_CONV_V29P23) ? 9u : 7u; }

// common per-stream front end shared by rx and det handles
struct Front {
    int device = 0;
    unsigned n_streams = 0, carry_cap = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    DevTables *d_tables = nullptr;
    StreamState *d_states = nullptr;
    float2 *d_carry[2] = { nullptr, nullptr };
    // per-call I/O state exists twice so that a call can be staged while the previous one is still being
    // finished on the GPU (pipelined receiver); the detector only ever uses set 0
    struct IoSet {
        DevBuf<StreamIO> d_io;
        PinBuf<StreamIO> h_io;
        DevBuf<float2> d_stage;
        DevBuf<int> d_stage16;            // raw sc16 samples (one int = one complex sample) before conversion
        unsigned *d_count = nullptr;
        unsigned *h_count = nullptr;
        size_t conv_pairs = 0;            // sc16 samples staged by a caller-supplied copy stream, widening still to be launched (take_conversion)
    } io[2];
    unsigned cur = 0;
    std::vector<uint8_t> fed;
    // search work a stream is expected to cost, in sample equivalents (from the last planned call): k_seek runs one CTA
    // per stream and streams differ several-fold (a stream full of decodable frames skips most of its samples, one
    // below the header threshold searches all of them and aligns at every detection), so the longest are started
    // first (profiles/r01_notes.md v19).  Only the ORDER of the io entries changes; results do not depend on it.
    std::vector<int64_t> est_work;
    bool lpt = getenv("LQB_NO_LPT") == nullptr;
    uint64_t launches = 0;
    // tensor-core pre-filter fused into k_seek: the B operand (template x 49 CFO rotations) in shared-memory layout
    bool coarse_ok = false;
    void *d_bmat = nullptr;
    float b_err = 0.0f;                   // rounding error of B relative to ||s|| (enters the pre-filter's bound)
    StreamState *h_states = nullptr;      // pinned staging for reset()
    DevBuf<unsigned> d_queue;             // slice queue of k_seek (searches of one front never overlap: one buffer)

    // Time slices of the search (k_seek): LQB_SEEK_SLICE=<samples per slice> turns them on for calls whose streams are
    // at least four slices long; unset or 0: one CTA per stream.  Never changes a result (tests/test_gpu_slices.py).
    // Off by default -- measured on the bench workload (1024 streams, one lane): the search alone 20.3 -> 19.65 ms
    // (every CTA slot busy to the end), but one sliced lane gives up the four-lane pipeline that hides the host's
    // planning and the payload chains behind other lanes' searches: step 31.3 -> 33.1 ms (profiles/r02_notes.md v24).
    int set_slices(SeekParams &sp, uint32_t n, const StreamIO *h_io)
    {
        sp.slice_len = 0; sp.n_io = n; sp.grid = n; sp.queue = nullptr;
        const char *ev = getenv("LQB_SEEK_SLICE");
        const long conf = ev ? atol(ev) : 0;
        if (conf <= 0) return 0;
        const uint64_t Q = ((uint64_t)conf + 255u) & ~uint64_t(255);
        uint64_t grid = 0, mx = 0;
        for (uint32_t i = 0; i < n; ++i) {
            grid += ((uint64_t)carry_cap + h_io[i].n_in + Q - 1) / Q + 1;
            mx = std::max<uint64_t>(mx, h_io[i].n_in);
        }
        if (mx < 4 * Q || grid > 0x7fffff00ull) return 0;
        if (int e = d_queue.reserve((size_t)grid + 4 + 3u * (size_t)sm_count_of_this_device())) return e;
        sp.slice_len = (unsigned)Q; sp.grid = (unsigned)grid; sp.queue = d_queue.p;
        return 0;
    }

    int init(int dev, unsigned ns, unsigned cap, void *user_stream, const DevTables &T, bool low_priority = false)
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(LQB_ENODEV, "no CUDA device available (no CPU fallback exists)");
        if (dev < 0 || dev >= ndev) return fail(LQB_ENODEV, "device ordinal out of range");
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10) return fail(LQB_ENODEV, "device is not sm_100 class; this library carries sm_100a code only");
        device = dev; n_streams = ns; carry_cap = cap;
        CU(cudaSetDevice(dev));
        if (user_stream) stream = (cudaStream_t)user_stream;
        else {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            CU(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, low_priority ? lo : hi));
            own_stream = true;
        }
        CU(cudaMalloc(&d_tables, sizeof(DevTables)));
        CU(cudaMemcpy(d_tables, &T, sizeof(DevTables), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&d_states, (size_t)ns * sizeof(StreamState)));
        for (int k = 0; k < 2; ++k) CU(cudaMalloc(&d_carry[k], (size_t)ns * cap * sizeof(float2)));
        for (auto &x : io) {
            CU(cudaMalloc(&x.d_count, 8 * sizeof(unsigned)));
            CU(cudaMallocHost(&x.h_count, 8 * sizeof(unsigned)));
        }
        CU(cudaMallocHost(&h_states, (size_t)ns * sizeof(StreamState)));
        fed.assign(ns, 0);
        est_work.assign(ns, 0);
        if (T.range == 24 && !getenv("LQB_NO_COARSE")) {
            std::vector<float> sre(kSLen), sim(kSLen);
            for (unsigned i = 0; i < kSLen; ++i) { sre[i] = T.sconj[i].x; sim[i] = -T.sconj[i].y; }
            std::vector<unsigned char> bm;
            b_err = build_coarse_bmat(sre.data(), sim.data(), 24, bm);
            CU(cudaMalloc(&d_bmat, bm.size()));
            CU(cudaMemcpy(d_bmat, bm.data(), bm.size(), cudaMemcpyHostToDevice));
            coarse_ok = true;
        }
        return reset(-1);
    }
    int reset(int s)
    {
        CU(cudaSetDevice(device));
        StreamState z;
        std::memset(&z, 0, sizeof z);
        z.wstart = -256;           // first window = 256 zeros + first 256 samples (qdetector reset state)
        z.stop_at = kNoStop; z.mark_at = kNoStop; z.mark_w = kNoMark;
        if (s < 0) {
            for (unsigned i = 0; i < n_streams; ++i) h_states[i] = z;
            CU(cudaMemcpyAsync(d_states, h_states, (size_t)n_streams * sizeof(StreamState), cudaMemcpyHostToDevice, stream));
            CU(cudaStreamSynchronize(stream));
        } else {
            if ((unsigned)s >= n_streams) return fail(LQB_EINVAL, "stream index out of range");
            h_states[s] = z;
            CU(cudaMemcpyAsync(d_states + s, &z, sizeof z, cudaMemcpyHostToDevice, stream));
            CU(cudaStreamSynchronize(stream));
        }
        return 0;
    }
    // stage inputs; fills d_io; returns total samples via *total
    int feed(uint32_t n, const uint32_t *ids, const float *const *iq, const uint64_t *ns, int mem, uint64_t *total, uint64_t *max_n, cudaStream_t cs = nullptr)
    {
        cudaStream_t stream = cs ? cs : this->stream;     // (shadows the member: copies go where the caller says)
        if (!n) { *total = 0; *max_n = 0; return 0; }
        if (n > n_streams) return fail(LQB_EINVAL, "more entries than streams");
        DevBuf<StreamIO> &d_io = io[cur].d_io;
        PinBuf<StreamIO> &h_io = io[cur].h_io;
        DevBuf<float2> &d_stage = io[cur].d_stage;
        if (int e = h_io.reserve(n)) return e;
        if (int e = d_io.reserve(n)) return e;
        std::fill(fed.begin(), fed.end(), 0);
        uint64_t tot = 0, mx = 0;
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t s = ids ? ids[i] : i;
            if (s >= n_streams) return fail(LQB_EINVAL, "stream index out of range");
            if (fed[s]) return fail(LQB_EINVAL, "stream listed twice in one execute");
            fed[s] = 1;
            if (ns[i] && !iq[i]) return fail(LQB_EINVAL, "null sample pointer");
            tot += ns[i]; mx = std::max<uint64_t>(mx, ns[i]);
        }
        const bool sc16 = (mem == LQB_MEM_HOST_SC16 || mem == LQB_MEM_DEVICE_SC16);
        io[cur].conv_pairs = 0;
        if (sc16) {
            // int16 pairs travel as they are (half the bytes of complex64 over PCIe) and are widened on the device:
            // packed back to back in feed order, then one conversion kernel on the same stream
            DevBuf<int> &d_raw = io[cur].d_stage16;
            if (int e = d_stage.reserve(tot + 2)) return e;
            if (int e = d_raw.reserve(tot + 2)) return e;
            const cudaMemcpyKind kind = mem == LQB_MEM_HOST_SC16 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
            bool regular = n >= 2 && ns[0] > 0;
            const ptrdiff_t pitch = regular ? (reinterpret_cast<const int *>(iq[1]) - reinterpret_cast<const int *>(iq[0])) : 0;
            for (uint32_t i = 1; i < n && regular; ++i)
                regular = ns[i] == ns[0] && (reinterpret_cast<const int *>(iq[i]) - reinterpret_cast<const int *>(iq[i - 1])) == pitch;
            regular = regular && pitch >= (ptrdiff_t)ns[0];
            uint64_t off = 0;
            if (regular) {
                const size_t row = ns[0] * sizeof(int);
                CU(cudaMemcpy2DAsync(d_raw.p, row, iq[0], (size_t)pitch * sizeof(int), row, n, kind, stream));
            }
            for (uint32_t k = 0; k < n; ++k) {
                if (!regular && ns[k]) CU(cudaMemcpyAsync(d_raw.p + off, iq[k], ns[k] * sizeof(int), kind, stream));
                h_io.p[k].in = d_stage.p + off; h_io.p[k].n_in = ns[k]; h_io.p[k].stream = ids ? ids[k] : k; h_io.p[k].pad = 0;
                off += ns[k];
            }
            const size_t pairs = (size_t)((tot + 1) / 2);
            // With a separate copy stream (the pipelined receiver) the widening is NOT queued behind the copy: it would sit
            // there until an SM is free of search CTAs and hold up the next call's H2D copy, which queues on the same stream
            // (measured: the link ran at 37 GB/s instead of 52).  The search stream launches it right before its k_seek.
            if (cs && cs != this->stream) io[cur].conv_pairs = pairs;
            else { io[cur].conv_pairs = pairs; launch_conversion(cur, stream); }
        } else if (mem == LQB_MEM_HOST) {
            if (int e = d_stage.reserve(tot + 1)) return e;
            // equal-length streams at a constant pitch in host memory (the dense layout, also after the lane split):
            // one strided 2-D copy instead of one call per stream
            bool regular = n >= 2 && ns[0] > 0;
            const ptrdiff_t pitch = regular ? (iq[1] - iq[0]) : 0;
            for (uint32_t i = 1; i < n && regular; ++i) regular = ns[i] == ns[0] && (iq[i] - iq[i - 1]) == pitch;
            regular = regular && pitch >= (ptrdiff_t)(2 * ns[0]);
            if (regular) {
                const size_t row = ns[0] * sizeof(float2);
                CU(cudaMemcpy2DAsync(d_stage.p, row, iq[0], (size_t)pitch * sizeof(float), row, n, cudaMemcpyHostToDevice, stream));
                for (uint32_t k = 0; k < n; ++k) {
                    h_io.p[k].in = d_stage.p + (size_t)k * ns[0]; h_io.p[k].n_in = ns[k]; h_io.p[k].stream = ids ? ids[k] : k; h_io.p[k].pad = 0;
                }
            } else {
                uint64_t off = 0;
                uint32_t i = 0;
                while (i < n) {                         // merge runs that are contiguous in host memory into one copy
                    uint32_t j = i; uint64_t run = ns[i];
                    while (j + 1 < n && iq[j + 1] == iq[j] + 2 * ns[j]) { ++j; run += ns[j]; }
                    if (run) CU(cudaMemcpyAsync(d_stage.p + off, iq[i], run * sizeof(float2), cudaMemcpyHostToDevice, stream));
                    for (uint32_t k = i; k <= j; ++k) {
                        h_io.p[k].in = d_stage.p + off; h_io.p[k].n_in = ns[k]; h_io.p[k].stream = ids ? ids[k] : k; h_io.p[k].pad = 0;
                        off += ns[k];
                    }
                    i = j + 1;
                }
            }
        } else {
            for (uint32_t i = 0; i < n; ++i) {
                h_io.p[i].in = reinterpret_cast<const float2 *>(iq[i]); h_io.p[i].n_in = ns[i];
                h_io.p[i].stream = ids ? ids[i] : i; h_io.p[i].pad = 0;
            }
        }
        if (lpt && n > 1)
            std::stable_sort(h_io.p, h_io.p + n, [&](const StreamIO &a, const StreamIO &b) { return est_work[a.stream] > est_work[b.stream]; });
        CU(cudaMemcpyAsync(d_io.p, h_io.p, n * sizeof(StreamIO), cudaMemcpyHostToDevice, stream));
        *total = tot; *max_n = mx;
        return 0;
    }
    // widen the staged sc16 samples of I/O set `set` on stream s (no-op when there is nothing pending)
    void launch_conversion(unsigned set, cudaStream_t s)
    {
        const size_t pairs = io[set].conv_pairs;
        io[set].conv_pairs = 0;
        if (!pairs) return;
        const unsigned grid = (unsigned)std::min<size_t>((pairs + 255) / 256, 148u * 16u);
        k_sc16_to_c32<<<grid, 256, 0, s>>>(reinterpret_cast<const int2 *>(io[set].d_stage16.p), reinterpret_cast<float4 *>(io[set].d_stage.p), pairs);
        launches++;
    }
    // the pre-filter fields of sp (coarse == 0: exact FFT search only, LQB_NO_COARSE=1)
    void set_coarse(SeekParams &sp) const
    {
        sp.bmat = d_bmat;
        sp.b_err = b_err;
        sp.coarse = coarse_ok ? 2 : 0;
    }

    void destroy()
    {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (d_tables) cudaFree(d_tables);
        if (d_states) cudaFree(d_states);
        for (int k = 0; k < 2; ++k) if (d_carry[k]) cudaFree(d_carry[k]);
        for (auto &x : io) {
            x.d_io.release(); x.h_io.release(); x.d_stage.release(); x.d_stage16.release();
            if (x.d_count) cudaFree(x.d_count);
            if (x.h_count) cudaFreeHost(x.h_count);
        }
        if (d_bmat) cudaFree(d_bmat);
        d_queue.release();
        if (h_states) cudaFreeHost(h_states);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

}  // namespace

// =================================================================== RX handle
// The receiver is split into independent LANES: every lane owns a fixed subset of the streams, its own
// CUDA streams, stream states, carries and arenas.  One host thread drives all lanes as a software
// pipeline -- seek(l) is queued for every lane, then for each lane in turn the host reads the frame
// list, plans and queues the payload kernels -- so lane l's payload kernels, host planning and result
// copies run while the tensor-core search of lanes l+1.. is still in flight, and with host inputs the
// H2D copy of lane l+1 runs under lane l's search.  Results do not depend on the lane count
// (tests/test_gpu_parity.py::test_lane_count_does_not_change_results).
namespace {

// everything one execute call produces; two generations alternate so that the payload chain of call k runs
// while call k+1 is being searched
struct RxGen {
    DevBuf<FrameDesc> d_frames;
    PinBuf<FrameDesc> h_frames;
    DevBuf<float2> d_syms;
    PinBuf<float2> h_syms;
    DevBuf<unsigned char> d_bufA, d_bufB, d_payload;
    PinBuf<unsigned char> h_payload;
    DevBuf<unsigned long long> d_dec;
    DevBuf<uint4> d_ckpt;
    DevBuf<unsigned> d_lists, d_tilemap;
    PinBuf<unsigned> h_lists;
    DevBuf<StreamView> d_views;
    // soft-decision path (LQB_RX_SOFT): soft bytes in transmission order, deinterleaved soft bytes, per-frame descriptors
    DevBuf<unsigned char> d_soft_raw, d_soft_d;
    DevBuf<SoftDesc> d_softdesc;
    PinBuf<SoftDesc> h_softdesc;
    unsigned n_frames = 0, n_fed = 0;
    uint64_t n_valid = 0;
    cudaEvent_t ev[7] = {};
    cudaEvent_t mf_done = nullptr;        // the matched filter (the only payload kernel that reads input / carry) has run
    cudaEvent_t done = nullptr;           // results are on the host
    cudaEvent_t staged = nullptr;         // inputs and the I/O list are on the device
    cudaEvent_t seek_done = nullptr;      // the search has run and its counters are on the host
    cudaEvent_t carry_done = nullptr;     // k_carry (the last reader of the caller's input buffers) has run
    bool carry_pending = false;
    bool mf_pending = false;
    float ms[6] = {};
    uint64_t work[6] = {};
    uint64_t bins = 0;                    // CFO bins the exact window evaluations visited (of 49 each)
    SeekParams sp;
    size_t max_frames = 0;
    uint64_t total = 0;
    void release()
    {
        d_frames.release(); h_frames.release(); d_syms.release(); h_syms.release();
        d_bufA.release(); d_bufB.release(); d_payload.release(); h_payload.release();
        d_dec.release(); d_ckpt.release(); d_lists.release(); h_lists.release(); d_tilemap.release(); d_views.release();
        d_soft_raw.release(); d_soft_d.release(); d_softdesc.release(); h_softdesc.release();
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        if (mf_done) cudaEventDestroy(mf_done);
        if (done) cudaEventDestroy(done);
        if (staged) cudaEventDestroy(staged);
        if (seek_done) cudaEventDestroy(seek_done);
        if (carry_done) cudaEventDestroy(carry_done);
    }
};

struct RxLane {
    Front f;
    cudaStream_t pay = nullptr;           // payload + gather stream (higher priority than the search stream)
    cudaStream_t copy = nullptr;          // input staging (H2D) stream: the next call's samples arrive under this call's search
    bool own_pay = false;
    unsigned flags = 0;
    unsigned lane = 0, n_lanes = 1;
    RxGen g[2];
    // true while the other generation has no call in flight: buffers that grow are then grown for BOTH generations,
    // so a caller alternating execute() calls pays every (slow, synchronising) pinned / device allocation once, on
    // the first call of a size, not again on the second (bench: a 100 ms second step)
    bool twin_idle = false;
    DevBuf<unsigned> d_ilv;
    size_t ilv_used = 0;
    std::unordered_map<unsigned, size_t> ilv_cache;
    DevBuf<unsigned> d_bitperm;           // soft-decision path: bit permutations of the deinterleaver, by block length
    size_t bitperm_used = 0;
    std::unordered_map<unsigned, size_t> bitperm_cache;
    // per-call input lists (lane-local stream ids)
    std::vector<uint32_t> ids;
    std::vector<const float *> iq;
    std::vector<uint64_t> ns;

    void destroy()
    {
        cudaSetDevice(f.device);
        if (f.stream) cudaStreamSynchronize(f.stream);
        if (pay) cudaStreamSynchronize(pay);
        for (auto &x : g) x.release();
        d_ilv.release(); d_bitperm.release();
        if (own_pay && pay) cudaStreamDestroy(pay);
        if (copy) { cudaStreamSynchronize(copy); cudaStreamDestroy(copy); }
        f.destroy();
    }

    size_t ilv_offset(unsigned n)
    {
        auto it = ilv_cache.find(n);
        if (it != ilv_cache.end()) return it->second;
        std::vector<uint32_t> maps = ilv_maps(n);
        size_t off = ilv_used;
        if (off + maps.size() + 4 > d_ilv.cap) cudaStreamSynchronize(pay);   // growing frees the old arena: no chain may still read it
        if (d_ilv.reserve(off + maps.size() + 4, true, pay)) return (size_t)-1;
        if (!maps.empty())
            cudaMemcpyAsync(d_ilv.p + off, maps.data(), maps.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, pay);
        cudaStreamSynchronize(pay);         // maps is a local; finish the copy before it dies
        ilv_used = off + maps.size();
        ilv_cache.emplace(n, off);
        return off;
    }

    size_t bitperm_offset(unsigned n)
    {
        auto it = bitperm_cache.find(n);
        if (it != bitperm_cache.end()) return it->second;
        std::vector<uint32_t> perm = ilv_bit_perm(n);
        size_t off = bitperm_used;
        if (off + perm.size() + 4 > d_bitperm.cap) cudaStreamSynchronize(pay);   // growing frees the old arena: no chain may still read it
        if (d_bitperm.reserve(off + perm.size() + 4, true, pay)) return (size_t)-1;
        if (!perm.empty())
            cudaMemcpyAsync(d_bitperm.p + off, perm.data(), perm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, pay);
        cudaStreamSynchronize(pay);         // perm is a local; finish the copy before it dies
        bitperm_used = off + perm.size();
        bitperm_cache.emplace(n, off);
        return off;
    }

    // phase A1: stage the inputs (H2D on the copy stream, so they travel under the previous call's search)
    int phase_stage(int mem, unsigned gen)
    {
        RxGen &G = g[gen];
        f.cur = gen;
        const uint32_t n = (uint32_t)ids.size();
        G.n_fed = n;
        G.n_frames = 0; G.n_valid = 0;
        std::memset(G.ms, 0, sizeof G.ms);
        std::memset(G.work, 0, sizeof G.work);
        G.total = 0;
        if (!n) return 0;
        uint64_t max_n = 0;
        if (int e = f.feed(n, ids.data(), iq.data(), ns.data(), mem, &G.total, &max_n, copy)) return e;
        CU(cudaEventRecord(G.staged, copy));
        // upper bound on frames: a frame spans at least 618 samples
        G.max_frames = 0;
        for (uint32_t i = 0; i < n; ++i) G.max_frames += (size_t)((ns[i] + f.carry_cap) / 600 + 2);
        if (int e = G.d_frames.reserve(G.max_frames)) return e;
        if (int e = G.h_frames.reserve(G.max_frames)) return e;
        if (int e = G.d_views.reserve(n)) return e;
        if (twin_idle) {
            RxGen &O = g[gen ^ 1u];
            if (int e = O.d_frames.reserve(G.max_frames)) return e;
            if (int e = O.h_frames.reserve(G.max_frames)) return e;
            if (int e = O.d_views.reserve(n)) return e;
            Front::IoSet &oi = f.io[gen ^ 1u];
            if (int e = oi.h_io.reserve(n)) return e;
            if (int e = oi.d_io.reserve(n)) return e;
            if (mem != LQB_MEM_DEVICE) if (int e = oi.d_stage.reserve(G.total + 2)) return e;
            if (mem == LQB_MEM_HOST_SC16 || mem == LQB_MEM_DEVICE_SC16) if (int e = oi.d_stage16.reserve(G.total + 2)) return e;
        }
        return 0;
    }

    // phase A2: queue the search on the lane's search stream.  This happens BEFORE the previous call's payload chain is
    // planned, so that the GPU goes straight from one search into the next while the host plans (the old order left it
    // idle for the 4 ms that the wait / frame list / plan / queue sequence of both lanes takes, profiles/r01_notes.md v20).
    int phase_search(unsigned gen)
    {
        RxGen &G = g[gen];
        f.cur = gen;
        const uint32_t n = G.n_fed;
        if (!n) return 0;
        cudaStream_t st = f.stream;
        CU(cudaStreamWaitEvent(st, G.staged, 0));
        f.launch_conversion(f.cur, st);            // sc16 input: widened here, ahead of the search that reads it
        G.sp.views = G.d_views.p;
        G.sp.tables = f.d_tables; G.sp.states = f.d_states; G.sp.io = f.io[f.cur].d_io.p;
        G.sp.carry[0] = f.d_carry[0]; G.sp.carry[1] = f.d_carry[1]; G.sp.carry_cap = f.carry_cap;
        G.sp.det_mode = 0; G.sp.frames = G.d_frames.p; G.sp.detections = nullptr;
        G.sp.n_out = f.io[f.cur].d_count; G.sp.max_out = (unsigned)G.max_frames;
        CU(cudaEventRecord(G.ev[0], st));
        f.set_coarse(G.sp);
        if (int e = f.set_slices(G.sp, n, f.io[f.cur].h_io.p)) return e;
        CU(cudaMemsetAsync(f.io[f.cur].d_count, 0, 8 * sizeof(unsigned), st));
        launch_seek(G.sp, n, st); f.launches += G.sp.slice_len ? 2 : 1;
        CU(cudaEventRecord(G.ev[1], st));
        launch_copy(f.io[f.cur].h_count, f.io[f.cur].d_count, 8 * sizeof(unsigned), st);
        CU(cudaEventRecord(G.seek_done, st));
        return 0;
    }
    // phase A3: the carry update, after the previous call's payload chain has been queued.
    // The unconsumed tails move to the other carry buffer right away, so that the next call can be searched while this
    // call's payload chain is still running.  That buffer is the one the PREVIOUS call's matched filter reads: wait for it.
    int phase_carry(unsigned gen)
    {
        RxGen &G = g[gen];
        f.cur = gen;
        const uint32_t n = G.n_fed;
        if (!n) return 0;
        cudaStream_t st = f.stream;
        RxGen &prev = g[gen ^ 1u];
        if (prev.mf_pending) { CU(cudaStreamWaitEvent(st, prev.mf_done, 0)); prev.mf_pending = false; }
        launch_carry(G.sp, n, st); f.launches++;
        // k_carry reads the caller's input (or the staging copy of it) on the low-priority search stream, where nothing
        // else orders it before collect(): the caller may touch its buffers once collect() returns, so collect waits
        CU(cudaEventRecord(G.carry_done, st));
        G.carry_pending = true;
        return 0;
    }

    // phase B: read the frame list, plan, queue payload kernels + carry + result copies on the payload stream
    int phase_payload(unsigned gen)
    {
        RxGen &G = g[gen];
        const uint32_t n = G.n_fed;
        if (!n) return 0;
        f.cur = gen;
        cudaStream_t ps = pay;
        g_trace.mark("wait seek", lane);
        CU(cudaEventSynchronize(G.seek_done));          // not the stream: the next call's search may already be queued behind
        g_trace.mark("seek done", lane);
        unsigned nf = std::min<unsigned>(f.io[f.cur].h_count[0], (unsigned)G.max_frames);
        G.work[0] = f.io[f.cur].h_count[1]; G.work[1] = f.io[f.cur].h_count[2]; G.work[2] = 0; G.work[3] = G.total; G.work[4] = f.io[f.cur].h_count[3];
        G.work[5] = f.io[f.cur].h_count[4];
        G.bins = f.io[f.cur].h_count[5];
        FrameDesc *fr = G.h_frames.p;
        if (nf) {
            // on the payload stream (idle: the chain of this generation's previous use has been collected), by kernel:
            // it fits beside the resident search CTAs of the next call
            launch_copy(fr, G.d_frames.p, nf * sizeof(FrameDesc), ps);
            g_trace.mark("frame list copy queued", lane);
            CU(cudaStreamSynchronize(ps));
        }

        g_trace.mark("frame list on host", lane);
        // ---------------- plan
        size_t sym_total = 0, buf_total = 0, pay_total = 0, dec_total = 0, n_tiles = 0, ck_total = 0;
        std::vector<unsigned> tile_start(nf + 1, 0), valid, deint[2], blk[2], vit[2], vit9[2], rsb[2], softl, vsoft7[2], vsoft9[2];
        const bool soft_on = (flags & LQB_RX_SOFT) != 0;
        size_t soft_raw_total = 0, soft_d_total = 0;
        unsigned soft_max_syms = 0, soft_max_bits = 0;
        SoftDesc *sdesc = nullptr;
        if (soft_on && nf) {
            if (int e = G.h_softdesc.reserve(nf)) return e;
            sdesc = G.h_softdesc.p;
            for (unsigned i = 0; i < nf; ++i) { sdesc[i].raw_off = 0; sdesc[i].d_off = 0; sdesc[i].perm_off = 0; sdesc[i].stage = -1; }
        }
        size_t tmax7[2] = { 0, 0 }, tmax_s7[2] = { 0, 0 };
        bool punct7[2] = { false, false }, punct_s7[2] = { false, false };
        std::fill(f.est_work.begin(), f.est_work.end(), 0);
        for (unsigned i = 0; i < nf; ++i) {
            FrameDesc &d = fr[i];
            tile_start[i] = (unsigned)n_tiles;
            // every detection costs an exact window and an alignment (about 3000 samples' worth of search); a frame with
            // a valid header lets the search skip its payload
            if (d.stream < f.est_work.size()) f.est_work[d.stream] += 3000 - (d.header_valid ? 2ll * d.n_sym : 0ll);
            if (!d.header_valid || (d.flags & 1u)) continue;       // flags bit 0: dropped (longer than the carry), header only
            valid.push_back(i);
            d.sym_off = sym_total; sym_total += (d.n_sym + 1u) & ~1u;      // even: 16-byte aligned symbol rows
            unsigned bl = std::max(std::max(d.n1, d.n0), d.k0) + 16;
            bl = (bl + 15u) & ~15u;
            d.buf_len = bl; d.buf_off = buf_total; buf_total += bl;
            d.pay_off = pay_total; pay_total += (d.payload_len + 3u) & ~3u;
            n_tiles += (d.n_sym + kMfTileSyms - 1) / kMfTileSyms;
            d.ck_off = (unsigned)ck_total; ck_total += (d.n_sym + 31) / 32;
            const unsigned fs[2] = { d.fec0, d.fec1 }, enc[2] = { d.n0, d.n1 }, dl[2] = { d.k0, d.n0 };
            size_t need_dec = 0;
            // soft decisions: the stage nearest the channel, when it is convolutional and the modem has a soft demodulator
            int sstage = -1;
            if (soft_on && !(d.ms >= 9 && d.ms <= 16)) {
                const int st = (d.fec1 != FEC_NONE) ? 1 : 0;
                if (is_conv(fs[st])) sstage = st;
            }
            if (sstage >= 0) {
                const size_t off = bitperm_offset(enc[sstage]);
                if (off == (size_t)-1) return LQB_ENOMEM;
                SoftDesc &sd = sdesc[i];
                sd.stage = sstage; sd.perm_off = (unsigned)off;
                sd.raw_off = soft_raw_total; soft_raw_total += ((size_t)d.n_sym * d.bps + 31u) & ~(size_t)15u;
                sd.d_off = soft_d_total; soft_d_total += ((size_t)8 * enc[sstage] + 31u) & ~(size_t)15u;
                soft_max_syms = std::max(soft_max_syms, d.n_sym); soft_max_bits = std::max(soft_max_bits, 8u * enc[sstage]);
                softl.push_back(i);
                const size_t T = (size_t)8 * dl[sstage] + conv_K(fs[sstage]) - 1;
                if (conv_K(fs[sstage]) == 7) {                // the packed four-lane decoder, soft input: shared [step][thread] arena
                    vsoft7[sstage].push_back(i); tmax_s7[sstage] = std::max(tmax_s7[sstage], T);
                    punct_s7[sstage] = punct_s7[sstage] || fs[sstage] != FEC_CONV_V27;
                } else { vsoft9[sstage].push_back(i); need_dec = std::max(need_dec, T * 4); }            // 8 words per step
            }
            for (int stg = 1; stg >= 0; --stg) {
                if (stg == sstage) continue;              // deinterleaved and decoded from soft bytes instead
                if (fs[stg] != FEC_NONE) {
                    size_t off = ilv_offset(enc[stg]);
                    if (off == (size_t)-1) return LQB_ENOMEM;
                    (stg ? d.ilv1_off : d.ilv0_off) = (unsigned)off;
                    deint[stg].push_back(i);
                }
                if (is_conv(fs[stg])) {
                    size_t T = (size_t)8 * dl[stg] + conv_K(fs[stg]) - 1;
                    if (conv_K(fs[stg]) == 7) { vit[stg].push_back(i); tmax7[stg] = std::max(tmax7[stg], T); punct7[stg] = punct7[stg] || fs[stg] != FEC_CONV_V27; }   // [step][thread] arena
                    else { vit9[stg].push_back(i); need_dec = std::max(need_dec, T * 4); }                      // 8 words per step
                } else if (fs[stg] == FEC_RS_M8) {
                    unsigned blocks = (dl[stg] + 222) / 223;
                    for (unsigned b = 0; b < blocks; ++b) { rsb[stg].push_back(i); rsb[stg].push_back(b); }
                } else {
                    blk[stg].push_back(i);
                }
            }
            d.dec_off = dec_total; dec_total += need_dec;
        }
        tile_start[nf] = (unsigned)n_tiles;
        // K=7 frames share one [step][thread] decision arena at the front; K=9 frames follow with private slices
        const size_t dec7 = std::max(std::max(tmax7[0] * vit[0].size(), tmax7[1] * vit[1].size()), std::max(tmax_s7[0] * vsoft7[0].size(), tmax_s7[1] * vsoft7[1].size()));
        for (unsigned i = 0; i < nf; ++i) fr[i].dec_off += dec7;
        dec_total += dec7;
        G.work[2] = sym_total;
        // group the PLL work list by modulation so warps diverge less
        std::vector<unsigned> pll = valid;
        std::stable_sort(pll.begin(), pll.end(), [&](unsigned a, unsigned b) { return fr[a].ms < fr[b].ms; });
        std::vector<unsigned> span_start(pll.size() + 1, 0);          // kEmitSpan-symbol spans of the emit pass, over the pll order
        for (size_t k = 0; k < pll.size(); ++k) span_start[k + 1] = span_start[k] + (fr[pll[k]].n_sym + kEmitSpan - 1) / kEmitSpan;

        g_trace.mark("planned", lane);
        if (nf && !valid.empty()) {
            if (int e = G.d_syms.reserve(sym_total + 1)) return e;
            if (int e = G.d_bufA.reserve(buf_total + 16)) return e;
            if (int e = G.d_bufB.reserve(buf_total + 16)) return e;
            if (int e = G.d_payload.reserve(pay_total + 16)) return e;
            if (int e = G.d_dec.reserve(dec_total + 1)) return e;
            if (int e = G.d_tilemap.reserve(12 * (n_tiles + 1))) return e;      // 48-byte records
            if (int e = G.d_ckpt.reserve(ck_total + 1)) return e;
            // one list arena: tile_start | pll | valid | deint1 | blk1 | vit1 | rs1 | deint0 | blk0 | vit0 | rs0
            std::vector<const std::vector<unsigned> *> parts = { &tile_start, &pll, &valid, &deint[1], &blk[1], &vit[1], &rsb[1],
                                                                 &deint[0], &blk[0], &vit[0], &rsb[0], &vit9[1], &vit9[0], &span_start,
                                                                 &softl, &vsoft7[1], &vsoft7[0], &vsoft9[1], &vsoft9[0] };
            size_t ltot = 0;
            std::vector<size_t> loff;
            for (auto p : parts) { loff.push_back(ltot); ltot += p->size(); }
            if (int e = G.h_lists.reserve(ltot + 1)) return e;
            if (int e = G.d_lists.reserve(ltot + 1)) return e;
            if (twin_idle) {
                RxGen &O = g[gen ^ 1u];
                if (int e = O.d_syms.reserve(sym_total + 1)) return e;
                if (int e = O.d_bufA.reserve(buf_total + 16)) return e;
                if (int e = O.d_bufB.reserve(buf_total + 16)) return e;
                if (int e = O.d_payload.reserve(pay_total + 16)) return e;
                if (int e = O.d_dec.reserve(dec_total + 1)) return e;
                if (int e = O.d_tilemap.reserve(12 * (n_tiles + 1))) return e;
                if (int e = O.d_ckpt.reserve(ck_total + 1)) return e;
                if (int e = O.h_lists.reserve(ltot + 1)) return e;
                if (int e = O.d_lists.reserve(ltot + 1)) return e;
                if (!(flags & LQB_RX_DEVICE_RESULTS)) {
                    if (int e = O.h_payload.reserve(pay_total + 16)) return e;
                    if (!(flags & LQB_RX_NO_FRAMESYMS)) if (int e = O.h_syms.reserve(sym_total + 1)) return e;
                }
            }
            for (size_t k = 0; k < parts.size(); ++k)
                if (!parts[k]->empty()) std::memcpy(G.h_lists.p + loff[k], parts[k]->data(), parts[k]->size() * sizeof(unsigned));
            launch_copy(G.d_lists.p, G.h_lists.p, ltot * sizeof(unsigned), ps);
            launch_copy(G.d_frames.p, fr, nf * sizeof(FrameDesc), ps);

            PayloadParams pp;
            pp.tables = f.d_tables; pp.views = G.d_views.p;
            pp.frames = G.d_frames.p; pp.n_frames = nf;
            pp.tile_start = G.d_lists.p + loff[0]; pp.n_tiles = (unsigned)n_tiles; pp.tile_rec = reinterpret_cast<uint4 *>(G.d_tilemap.p);
            pp.syms = G.d_syms.p; pp.bufA = G.d_bufA.p; pp.bufB = G.d_bufB.p; pp.payload = G.d_payload.p;
            pp.ilv_maps = d_ilv.p; pp.decisions = G.d_dec.p; pp.pll_ckpt = G.d_ckpt.p;
            pp.soft = nullptr; pp.soft_raw = nullptr; pp.soft_d = nullptr; pp.bitperm = d_bitperm.p;
            if (!softl.empty()) {
                if (int e = G.d_soft_raw.reserve(soft_raw_total + 64)) return e;
                if (int e = G.d_soft_d.reserve(soft_d_total + 64)) return e;
                if (int e = G.d_softdesc.reserve(nf)) return e;
                launch_copy(G.d_softdesc.p, sdesc, nf * sizeof(SoftDesc), ps);
                pp.soft = G.d_softdesc.p; pp.soft_raw = G.d_soft_raw.p; pp.soft_d = G.d_soft_d.p;
            }

            CU(cudaEventRecord(G.ev[2], ps));
            launch_mf(pp, ps); f.launches += n_tiles ? 2 : 0;
            CU(cudaEventRecord(G.mf_done, ps));
            G.mf_pending = true;
            CU(cudaEventRecord(G.ev[3], ps));
            launch_pll(pp, G.d_lists.p + loff[1], G.d_lists.p + loff[13], (unsigned)pll.size(), span_start.back(), ps);
            f.launches += getenv("LQB_PLL_FUSED") ? 1 : 2;          // tracker + emitter kernels (one fused kernel on request)
            CU(cudaEventRecord(G.ev[4], ps));
            if (!softl.empty()) { launch_soft_demod(pp, G.d_lists.p + loff[14], (unsigned)softl.size(), soft_max_syms, soft_max_bits, ps); f.launches += 2; }
            for (int stg = 1; stg >= 0; --stg) {
                const size_t base = stg ? 3 : 7;
                if (!vsoft7[stg].empty()) { launch_viterbi_soft(pp, G.d_lists.p + loff[stg ? 15 : 16], (unsigned)vsoft7[stg].size(), stg, 7, punct_s7[stg], ps); f.launches++; }
                if (!vsoft9[stg].empty()) { launch_viterbi_soft(pp, G.d_lists.p + loff[stg ? 17 : 18], (unsigned)vsoft9[stg].size(), stg, 9, true, ps); f.launches++; }
                if (!deint[stg].empty()) { launch_deinterleave(pp, G.d_lists.p + loff[base], (unsigned)deint[stg].size(), stg, ps); f.launches++; }
                if (!blk[stg].empty()) { launch_blockfec(pp, G.d_lists.p + loff[base + 1], (unsigned)blk[stg].size(), stg, ps); f.launches++; }
                if (!vit[stg].empty()) { launch_viterbi(pp, G.d_lists.p + loff[base + 2], (unsigned)vit[stg].size(), stg, 7, punct7[stg], ps); f.launches++; }
                if (!vit9[stg].empty()) { launch_viterbi(pp, G.d_lists.p + loff[stg ? 11 : 12], (unsigned)vit9[stg].size(), stg, 9, false, ps); f.launches++; }
                if (!rsb[stg].empty()) { launch_rs(pp, G.d_lists.p + loff[base + 3], (unsigned)(rsb[stg].size() / 2), stg, ps); f.launches++; }
            }
            launch_crc(pp, G.d_lists.p + loff[2], (unsigned)valid.size(), ps); f.launches++;
            CU(cudaEventRecord(G.ev[5], ps));
        } else {
            for (int k = 2; k <= 5; ++k) CU(cudaEventRecord(G.ev[k], ps));
        }
        CU(cudaEventRecord(G.ev[6], ps));

        // ---------------- gather
        if (nf && !valid.empty()) {
            CU(cudaMemcpyAsync(fr, G.d_frames.p, nf * sizeof(FrameDesc), cudaMemcpyDeviceToHost, ps));
            if (!(flags & LQB_RX_DEVICE_RESULTS)) {
                if (int e = G.h_payload.reserve(pay_total + 16)) return e;
                if (pay_total) CU(cudaMemcpyAsync(G.h_payload.p, G.d_payload.p, pay_total, cudaMemcpyDeviceToHost, ps));
                if (!(flags & LQB_RX_NO_FRAMESYMS)) {
                    if (int e = G.h_syms.reserve(sym_total + 1)) return e;
                    if (sym_total) CU(cudaMemcpyAsync(G.h_syms.p, G.d_syms.p, sym_total * sizeof(float2), cudaMemcpyDeviceToHost, ps));
                }
            }
        }
        CU(cudaEventRecord(G.done, ps));
        G.n_frames = nf;
        g_trace.mark("payload queued", lane);
        return 0;
    }

    // phase C: wait for the lane, read the event times
    int phase_finish(unsigned gen)
    {
        RxGen &G = g[gen];
        if (!G.n_fed) return 0;
        CU(cudaEventSynchronize(G.done));
        if (G.carry_pending) { CU(cudaEventSynchronize(G.carry_done)); G.carry_pending = false; }
        g_trace.mark("lane complete", lane);
        CU(cudaGetLastError());
        cudaEventElapsedTime(&G.ms[0], G.ev[0], G.ev[1]);
        for (int k = 1; k < 4; ++k) cudaEventElapsedTime(&G.ms[k], G.ev[k + 1], G.ev[k + 2]);
        cudaEventElapsedTime(&G.ms[4], G.ev[0], G.ev[6]);
        G.ms[5] = 0.0f;
        const FrameDesc *fr = G.h_frames.p;
        for (unsigned i = 0; i < G.n_frames; ++i) G.n_valid += fr[i].payload_valid ? 1 : 0;
        return 0;
    }
};

}  // namespace

struct lqb_rx_s {
    int device = 0;
    unsigned n_streams = 0, flags = 0;
    std::vector<RxLane *> lanes;
    cudaStream_t user_stream = nullptr;
    cudaEvent_t ev_in = nullptr;
    std::vector<cudaEvent_t> ev_out;
    // two calls may be in flight: `submit_gen` is the generation the next submit uses, `pending` how many
    // submits have not been collected, `cur_gen` the generation whose results poll / counts / timing report
    unsigned submit_gen = 0, pending = 0, cur_gen = 0;
    bool planned[2] = { false, false };      // the payload chain of that generation has been planned and queued
    std::vector<std::pair<unsigned, unsigned>> order;    // (lane, frame index) sorted by (stream, seq)
    std::vector<unsigned> stream_count;
    unsigned n_frames = 0;
    uint64_t n_valid = 0;
    float ms[6] = {};
    uint64_t work[6] = {};
    uint64_t bins = 0;
    // stream -> (lane, index inside the lane) and back (a fixed interleaved partition)
    std::vector<unsigned> lane_of, local_of;
    std::vector<std::vector<unsigned>> global_of;
    unsigned global_stream(unsigned lane, unsigned local) const { return global_of[lane][local]; }
    // time-sharded decoding of one capture (lqb_rx_execute_sharded): the accepted frames, owned by the handle
    struct OwnedFrame {
        lqb_frame_result r;
        long long trig_w;                 // start of the window that triggered
        long long after_w, after_G;       // the walk's state once this frame is through (time-sharded decoding)
        int pool;                         // which copy of a lane's result arenas (shard_pools) holds its bytes / points
        size_t pay_off, sym_off;
        bool has_payload, has_syms;
    };
    std::vector<OwnedFrame> merged;
    // one bulk copy of every lane's payload pool / symbol arena per execute of the sharded call (a copy per frame cost
    // 0.5 us each: 16 ms of the 59 ms a 31 000-frame capture took)
    struct ShardPool { std::vector<unsigned char> pay; std::vector<float2> syms; };
    std::vector<std::unique_ptr<ShardPool>> shard_pools;
    bool use_merged = false;
    uint64_t merged_valid = 0;
    uint64_t shard[4] = {};               // segments, segment runs in all, rounds, execute calls
    DevBuf<float2> d_capture;
    // stream states of the listed streams <- z[i] (host staging of each lane, then one copy per lane)
    int put_states(unsigned n, const StreamState *z)
    {
        for (unsigned i = 0; i < n; ++i) lanes[lane_of[i]]->f.h_states[local_of[i]] = z[i];
        for (auto *l : lanes) {
            CU(cudaMemcpyAsync(l->f.d_states, l->f.h_states, (size_t)l->f.n_streams * sizeof(StreamState), cudaMemcpyHostToDevice, l->f.stream));
            CU(cudaStreamSynchronize(l->f.stream));
        }
        return 0;
    }
    int get_states()
    {
        for (auto *l : lanes) {
            CU(cudaStreamSynchronize(l->f.stream));
            CU(cudaMemcpyAsync(l->f.h_states, l->f.d_states, (size_t)l->f.n_streams * sizeof(StreamState), cudaMemcpyDeviceToHost, l->f.stream));
            CU(cudaStreamSynchronize(l->f.stream));
        }
        return 0;
    }
    void sync_all()
    {
        for (auto *l : lanes) {
            if (l->copy) cudaStreamSynchronize(l->copy);
            if (l->f.stream) cudaStreamSynchronize(l->f.stream);
            if (l->pay) cudaStreamSynchronize(l->pay);
        }
    }
    // wait for the search of generation `gen`, plan its payload work on the host and queue it
    int plan(unsigned gen)
    {
        if (planned[gen]) return 0;
        int rc = 0;
        for (auto *l : lanes) if ((rc = l->phase_payload(gen))) break;
        planned[gen] = true;
        return rc;
    }
};

extern "C" {

const char *lqb_last_error(void) { return g_err.c_str(); }
void lqb_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
int lqb_version(void) { return LQB_VERSION; }
int lqb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void lqb_rx_destroy(lqb_rx h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    h->sync_all();
    for (auto *l : h->lanes) { l->destroy(); delete l; }
    h->d_capture.release();
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    for (auto &e : h->ev_out) if (e) cudaEventDestroy(e);
    delete h;
}

lqb_rx lqb_rx_create(const lqb_rx_opts *o)
{
    if (!o || !o->n_streams) { fail(LQB_EINVAL, "bad options"); return nullptr; }
    DevTables *T = new DevTables;
    build_tables(*T, kRxBeta, 0.5f, 0.3f);
    lqb_rx h = new lqb_rx_s;
    h->device = o->device; h->n_streams = o->n_streams; h->flags = o->flags;
    h->user_stream = (cudaStream_t)o->cuda_stream;
    unsigned cap = o->max_frame_samples ? o->max_frame_samples : 65536u;
    if (cap < 2048) cap = 2048;
    // lane count: explicit option, else LQB_RX_LANES, else one lane per 256 streams up to 4 (measured on the bench
    // workload, 1024 streams: 1 lane 42.0, 2 lanes 40.1, 3 or 4 lanes 39.2, 6 lanes 41.0, 8 lanes 42.3 ms per step)
    unsigned L = o->n_lanes;
    if (!L) { const char *e = getenv("LQB_RX_LANES"); if (e) L = (unsigned)atoi(e); }
    if (!L) L = std::min(4u, std::max(1u, o->n_streams / 256u));
    L = std::max(1u, std::min(L, std::min(o->n_streams, 64u)));
    {
        // equal lanes by default; LQB_LANE_WEIGHTS="40,30,20,10" makes them unequal (measured: no gain on B200)
        std::vector<double> w(L, 1.0);
        const char *we = getenv("LQB_LANE_WEIGHTS");
        if (we) { unsigned l = 0; for (const char *p = we; *p && l < L; ++l) { w[l] = std::max(1e-3, atof(p)); while (*p && *p != ',') ++p; if (*p) ++p; } }
        for (int attempt = 0; attempt < 2; ++attempt) {
            double wsum = 0.0;
            for (double x : w) wsum += x;
            h->lane_of.assign(o->n_streams, 0); h->local_of.assign(o->n_streams, 0); h->global_of.assign(L, {});
            // largest-deficit assignment: deterministic, interleaves the lanes over the stream index
            for (unsigned s2 = 0; s2 < o->n_streams; ++s2) {
                unsigned best = 0; double bd = -1e300;
                for (unsigned l = 0; l < L; ++l) {
                    const double dfc = w[l] / wsum * (double)(s2 + 1) - (double)h->global_of[l].size();
                    if (dfc > bd) { bd = dfc; best = l; }
                }
                h->lane_of[s2] = best; h->local_of[s2] = (unsigned)h->global_of[best].size();
                h->global_of[best].push_back(s2);
            }
            bool empty = false;
            for (unsigned l = 0; l < L; ++l) empty = empty || h->global_of[l].empty();
            if (!empty) break;
            std::fill(w.begin(), w.end(), 1.0);         // weights left a lane without streams: fall back to equal lanes
        }
    }
    int lo = 0, hi = 0;
    bool ok = true;
    for (unsigned l = 0; l < L && ok; ++l) {
        RxLane *ln = new RxLane;
        h->lanes.push_back(ln);
        ln->lane = l; ln->n_lanes = L; ln->flags = o->flags;
        const unsigned ns = (unsigned)h->global_of[l].size();
        // every lane owns a low-priority search stream and a high-priority payload stream; a caller's stream is
        // ordered against them with events (inputs before the search, the caller's later work after the results)
        if (ln->f.init(o->device, ns, cap, nullptr, *T, /*low_priority=*/true)) { ok = false; break; }
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ln->pay, cudaStreamNonBlocking, hi) != cudaSuccess) { fail(LQB_ECUDA, "cudaStreamCreate failed"); ok = false; break; }
        ln->own_pay = true;
        if (cudaStreamCreateWithPriority(&ln->copy, cudaStreamNonBlocking, hi) != cudaSuccess) { fail(LQB_ECUDA, "cudaStreamCreate failed"); ok = false; break; }
        for (auto &G : ln->g) {
            for (auto &ev : G.ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&G.mf_done, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&G.done, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&G.staged, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&G.seek_done, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&G.carry_done, cudaEventDisableTiming) == cudaSuccess;
        }
    }
    delete T;
    if (ok && h->user_stream) {
        ok = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) == cudaSuccess;
        h->ev_out.assign(L, nullptr);
        for (auto &e : h->ev_out) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { lqb_rx_destroy(h); return nullptr; }
    return h;
}

// wait for the oldest submitted call and make its results current
int lqb_rx_collect(lqb_rx h)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (!h->pending) return fail(LQB_EINVAL, "nothing submitted");
    CU(cudaSetDevice(h->device));
    const unsigned L = (unsigned)h->lanes.size();
    const unsigned gen = (h->pending == 2) ? h->submit_gen : (h->submit_gen ^ 1u);
    h->pending--;
    h->cur_gen = gen;
    h->n_frames = 0; h->n_valid = 0; h->order.clear();
    std::memset(h->ms, 0, sizeof h->ms);
    std::memset(h->work, 0, sizeof h->work);
    h->bins = 0;
    for (auto *l : h->lanes) l->twin_idle = (h->pending == 0);       // (already decremented: nothing else in flight)
    int rc = h->plan(gen);
    if (!rc) for (auto *l : h->lanes) if ((rc = l->phase_finish(gen))) break;
    if (rc) {
        const std::string keep = g_err;
        h->sync_all(); cudaGetLastError();
        for (auto *l : h->lanes) for (auto &G : l->g) { G.mf_pending = false; G.carry_pending = false; G.n_fed = 0; G.n_frames = 0; }
        h->pending = 0;
        g_err = keep;
        return rc;
    }
    if (h->ev_in) {
        // the caller's stream continues after everything this call queued
        for (unsigned l = 0; l < L; ++l) {
            CU(cudaEventRecord(h->ev_out[l], h->lanes[l]->pay));
            CU(cudaStreamWaitEvent(h->user_stream, h->ev_out[l], 0));
            CU(cudaStreamWaitEvent(h->user_stream, h->lanes[l]->g[gen].carry_done, 0));    // last reader of the inputs
        }
    }
    std::vector<unsigned> &cnt = h->stream_count;        // frames per global stream, then running offsets
    cnt.assign(h->n_streams + 1, 0);
    for (unsigned l = 0; l < L; ++l) {
        const RxGen &G = h->lanes[l]->g[gen];
        for (int k = 0; k < 6; ++k) h->work[k] += G.work[k];
        h->bins += G.bins;
        for (int k = 0; k < 6; ++k) if (k != 4) h->ms[k] += G.ms[k];
        h->ms[4] = std::max(h->ms[4], G.ms[4]);
        h->n_frames += G.n_frames; h->n_valid += G.n_valid;
        const FrameDesc *fr = G.h_frames.p;
        for (unsigned i = 0; i < G.n_frames; ++i) cnt[h->global_stream(l, fr[i].stream) + 1]++;
    }
    // Order by (stream, seq).  A stream is walked by one CTA that emits its frames in time order, so within a
    // lane's list the frames of a stream already appear by increasing seq: a counting sort on the stream is enough.
    for (unsigned s2 = 0; s2 < h->n_streams; ++s2) cnt[s2 + 1] += cnt[s2];
    h->order.resize(h->n_frames);
    for (unsigned l = 0; l < L; ++l) {
        const RxGen &G = h->lanes[l]->g[gen];
        const FrameDesc *fr = G.h_frames.p;
        for (unsigned i = 0; i < G.n_frames; ++i) h->order[cnt[h->global_stream(l, fr[i].stream)]++] = std::make_pair(l, i);
    }
    g_trace.mark("results ordered", 0);
    return 0;
}

int lqb_rx_reset(lqb_rx h, int stream)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    while (h->pending) if (int e = lqb_rx_collect(h)) return e;      // nothing may be in flight while states are rewritten
    h->sync_all();
    if (stream < 0) {
        for (auto *l : h->lanes) if (int e = l->f.reset(-1)) return e;
        return 0;
    }
    if ((unsigned)stream >= h->n_streams) return fail(LQB_EINVAL, "stream index out of range");
    return h->lanes[h->lane_of[(unsigned)stream]]->f.reset((int)h->local_of[(unsigned)stream]);
}

// Search the new samples and queue the payload work; returns when the search has finished (the payload kernels,
// result copies and CRC checks of this call keep running on the GPU until lqb_rx_collect).
int lqb_rx_submit(lqb_rx h, uint32_t n, const uint32_t *ids, const float *const *iq, const uint64_t *ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (h->pending >= 2) return fail(LQB_EBUSY, "two submitted calls are waiting for lqb_rx_collect");
    CU(cudaSetDevice(h->device));
    h->use_merged = false;
    if (n > h->n_streams) return fail(LQB_EINVAL, "more entries than streams");
    for (auto *l : h->lanes) { l->ids.clear(); l->iq.clear(); l->ns.clear(); }
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = ids ? ids[i] : i;
        if (s >= h->n_streams) return fail(LQB_EINVAL, "stream index out of range");
        RxLane *l = h->lanes[h->lane_of[s]];
        l->ids.push_back(h->local_of[s]); l->iq.push_back(iq[i]); l->ns.push_back(ns[i]);
    }
    const unsigned gen = h->submit_gen;
    // inputs produced on the caller's stream must be complete before any lane reads them
    if (h->ev_in && n) {
        CU(cudaEventRecord(h->ev_in, h->user_stream));
        for (auto *l : h->lanes) { CU(cudaStreamWaitEvent(l->copy, h->ev_in, 0)); CU(cudaStreamWaitEvent(l->f.stream, h->ev_in, 0)); }
    }
    int rc = 0;
    g_trace.start();
    for (auto *l : h->lanes) l->twin_idle = (h->pending == 0);
    // 1. the new samples start travelling (copy streams) while the previous call is still being searched;
    // 2. this call's search is queued behind its own copies (and behind the previous search, same stream);
    // 3. the previous call's search is awaited, its payload work planned and queued;
    // 4. this call's carry update is queued behind the previous call's matched filter.
    for (auto *l : h->lanes) if ((rc = l->phase_stage(mem, gen))) break;
    g_trace.mark("inputs queued (all lanes)", 0);
    if (!rc) for (auto *l : h->lanes) if ((rc = l->phase_search(gen))) break;
    g_trace.mark("search queued (all lanes)", 0);
    if (!rc && h->pending == 1) rc = h->plan(gen ^ 1u);
    if (!rc) for (auto *l : h->lanes) if ((rc = l->phase_carry(gen))) break;
    if (rc) {
        const std::string keep = g_err;
        h->sync_all(); cudaGetLastError();
        for (auto *l : h->lanes) for (auto &G : l->g) { G.mf_pending = false; G.carry_pending = false; G.n_fed = 0; G.n_frames = 0; }
        h->pending = 0;
        g_err = keep;
        return rc;
    }
    h->planned[gen] = false;
    h->submit_gen ^= 1u;
    h->pending++;
    return 0;
}

int lqb_rx_execute(lqb_rx h, uint32_t n, const uint32_t *ids, const float *const *iq, const uint64_t *ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    while (h->pending) if (int e = lqb_rx_collect(h)) return e;      // results of uncollected submits are dropped
    if (int e = lqb_rx_submit(h, n, ids, iq, ns, mem)) return e;
    return lqb_rx_collect(h);
}

int lqb_rx_execute_dense(lqb_rx h, const float *iq, uint64_t stride, uint64_t ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    unsigned n = h->n_streams;
    std::vector<const float *> ptr(n);
    std::vector<uint64_t> len(n, ns);
    for (unsigned s = 0; s < n; ++s) ptr[s] = iq + (mem >= LQB_MEM_HOST_SC16 ? 1 : 2) * (size_t)s * stride;   // a sample is 2 floats, or 1 float-sized int16 pair
    return lqb_rx_execute(h, n, nullptr, ptr.data(), len.data(), mem);
}

int lqb_rx_submit_dense(lqb_rx h, const float *iq, uint64_t stride, uint64_t ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    unsigned n = h->n_streams;
    std::vector<const float *> ptr(n);
    std::vector<uint64_t> len(n, ns);
    for (unsigned s = 0; s < n; ++s) ptr[s] = iq + (mem >= LQB_MEM_HOST_SC16 ? 1 : 2) * (size_t)s * stride;   // a sample is 2 floats, or 1 float-sized int16 pair
    return lqb_rx_submit(h, n, nullptr, ptr.data(), len.data(), mem);
}

// One capture, decoded as ONE flexframesync would decode it from its reset state (reference call site
// lib/flex_rx_impl.cc:213), but cut in TIME into segments that run side by side as the streams of this handle.  The seam
// rule is the detector's (lqb_det_execute_sharded): a segment is first run speculatively from `preroll` samples before
// its boundary; the state of the walk is (next window start, zero boundary G); a run is accepted only if the first
// state at or beyond its boundary equals the state the accepted run before it stopped in, else the segment is run again
// from exactly that state.  A frame belongs to the segment in which the window that triggered it starts; a run is fed
// max_frame_samples beyond its segment so that such a frame is decoded in full.  Every accepted run is a piece of the
// sequential walk, so the frames (bytes, flags, estimates) are the sequential receiver's whatever the cut.
// Results: lqb_rx_poll / lqb_rx_counts (stream 0, seq = order; payload / framesyms owned by the handle until the next
// call).  Consumes the stream states: lqb_rx_reset before going back to lqb_rx_execute.
int lqb_rx_execute_sharded(lqb_rx h, const float *iq, uint64_t n_samples, int mem, uint32_t seg_len, uint32_t preroll)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (mem != LQB_MEM_HOST && mem != LQB_MEM_DEVICE) return fail(LQB_EINVAL, "complex64 input only");
    if (h->flags & LQB_RX_DEVICE_RESULTS) return fail(LQB_EINVAL, "the sharded call returns host results");
    if (!iq && n_samples) return fail(LQB_EINVAL, "null sample pointer");
    while (h->pending) if (int e = lqb_rx_collect(h)) return e;
    CU(cudaSetDevice(h->device));
    const unsigned cap = h->lanes[0]->f.carry_cap;
    if (!seg_len) seg_len = 1u << 20;
    seg_len = std::max(4096u, (seg_len + 255u) & ~255u);
    preroll = std::min((preroll + 255u) & ~255u, seg_len);
    h->merged.clear(); h->merged_valid = 0;
    size_t pools_used = 0;                             // (the pools of the previous call are recycled: their pages are already mapped)
    std::memset(h->shard, 0, sizeof h->shard);
    if (n_samples < 512) { h->use_merged = true; return 0; }
    const float2 *x = reinterpret_cast<const float2 *>(iq);
    if (mem == LQB_MEM_HOST) {
        if (int e = h->d_capture.reserve(n_samples + 2)) return e;
        CU(cudaMemcpy(h->d_capture.p, iq, n_samples * sizeof(float2), cudaMemcpyHostToDevice));
        x = h->d_capture.p;
    } else if (h->user_stream) CU(cudaStreamSynchronize(h->user_stream));      // the capture is complete before any lane reads it
    const long long N = (long long)n_samples, L = seg_len;
    const unsigned K = (unsigned)((N + L - 1) / L);
    struct Seg {
        long long start_w = 0, start_G = 0;     // state the current / next run begins in
        long long entry_w = 0, entry_G = 0, exit_w = 0, exit_G = 0;
        bool ran = false, fin = false, todo = true;
        std::vector<lqb_rx_s::OwnedFrame> fr;
        // a re-run that only goes as far as the first frame of the run before it (old[splice]) and is spliced onto the
        // rest of that run when it leaves the frame in the same state
        std::vector<lqb_rx_s::OwnedFrame> old;
        long long old_exit_w = 0, old_exit_G = 0, partial_stop = 0;
        int splice = -1;
    };
    // G only matters while it lies above the window start
    auto norm_G = [](long long w, long long G) { return G > w ? G : w; };
    std::vector<Seg> seg(K);
    for (unsigned k = 0; k < K; ++k) {
        seg[k].start_w = k ? std::max<long long>(0, (long long)k * L - (long long)preroll) : -256;
        seg[k].start_G = k ? seg[k].start_w : 0;
    }
    h->shard[0] = K;
    const unsigned W = h->n_streams;
    std::vector<StreamState> zs;
    std::vector<const float *> ptr;
    std::vector<uint64_t> len;
    std::vector<uint32_t> ids;
    const bool shard_trace = getenv("LQB_SHARD_TRACE") != nullptr;
    double t_exec = 0.0, t_copy = 0.0, t_state = 0.0;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    while (true) {
        std::vector<unsigned> todo;
        for (unsigned k = 0; k < K; ++k) if (seg[k].todo) todo.push_back(k);
        if (todo.empty()) break;
        h->shard[2]++;
        for (size_t t0 = 0; t0 < todo.size(); t0 += W) {
            const unsigned nb = (unsigned)std::min<size_t>(W, todo.size() - t0);
            zs.assign(nb, StreamState{}); ptr.resize(nb); len.resize(nb); ids.resize(nb);
            for (unsigned i = 0; i < nb; ++i) {
                const unsigned k = todo[t0 + i];
                Seg &sg = seg[k];
                const long long base = std::max<long long>(0, std::min(sg.start_w, sg.start_G) - 256);
                const long long end = std::min<long long>(N, (long long)(k + 1) * L + (long long)cap + 1024);
                StreamState &z = zs[i];
                std::memset(&z, 0, sizeof z);
                z.base = base; z.G = std::max(sg.start_G > sg.start_w ? sg.start_G : base, k ? base : 0ll); z.wstart = sg.start_w;
                if (!k && sg.start_w == -256) { z.base = 0; z.G = 0; }
                z.stop_at = (k + 1 < K) ? (long long)(k + 1) * L : kNoStop;
                if (sg.splice >= 0) z.stop_at = std::min(z.stop_at, sg.partial_stop);
                z.mark_at = k ? (long long)k * L : -256;
                z.mark_w = kNoMark; z.mark_G = 0;
                ptr[i] = reinterpret_cast<const float *>(x + z.base);
                len[i] = (uint64_t)std::max<long long>(0, end - z.base);
                ids[i] = i;
            }
            double ta = now();
            if (int e = h->put_states(nb, zs.data())) return e;
            double tb = now(); t_state += tb - ta;
            if (int e = lqb_rx_execute(h, nb, ids.data(), ptr.data(), len.data(), LQB_MEM_DEVICE)) return e;
            h->shard[3]++;
            ta = now(); t_exec += ta - tb;
            if (int e = h->get_states()) return e;
            tb = now(); t_state += tb - ta;
            for (unsigned i = 0; i < nb; ++i) {
                Seg &sg = seg[todo[t0 + i]];
                const StreamState &z = h->lanes[h->lane_of[i]]->f.h_states[h->local_of[i]];
                sg.ran = true; sg.todo = false; sg.fr.clear();
                // a walk that ends inside a frame it cannot finish (the capture ends there) has no further state
                if (z.mode == 0) { sg.exit_w = z.wstart; sg.exit_G = norm_G(z.wstart, z.G); } else { sg.exit_w = kNoStop; sg.exit_G = kNoStop; }
                if (z.mark_w == kNoMark) { sg.entry_w = kNoStop; sg.entry_G = kNoStop; }
                else { sg.entry_w = z.mark_w; sg.entry_G = norm_G(z.mark_w, z.mark_G); }
                h->shard[1]++;
            }
            // keep this call's frames (the receiver's result arenas are reused by the next call)
            const double tc0 = now();
            // this execute's result arenas, lane by lane, in one copy each (extent = what its frames reference)
            const int pool0 = (int)pools_used;
            {
                std::vector<size_t> pay_ext(h->lanes.size(), 0), sym_ext(h->lanes.size(), 0);
                for (unsigned kf = 0; kf < h->n_frames; ++kf) {
                    const unsigned li = h->order[kf].first;
                    const FrameDesc &d = h->lanes[li]->g[h->cur_gen].h_frames.p[h->order[kf].second];
                    if (!d.header_valid || (d.flags & 1u)) continue;
                    pay_ext[li] = std::max(pay_ext[li], (size_t)d.pay_off + d.payload_len);
                    sym_ext[li] = std::max(sym_ext[li], (size_t)d.sym_off + d.n_sym);
                }
                for (size_t li = 0; li < h->lanes.size(); ++li) {
                    const RxGen &G = h->lanes[li]->g[h->cur_gen];
                    if (pools_used == h->shard_pools.size()) h->shard_pools.emplace_back(new lqb_rx_s::ShardPool);
                    lqb_rx_s::ShardPool *sp = h->shard_pools[pools_used++].get();
                    sp->pay.assign(G.h_payload.p, G.h_payload.p + pay_ext[li]);
                    if (!(h->flags & LQB_RX_NO_FRAMESYMS)) sp->syms.assign(G.h_syms.p, G.h_syms.p + sym_ext[li]); else sp->syms.clear();
                }
            }
            for (unsigned kf = 0; kf < h->n_frames; ++kf) {
                const RxLane *ln = h->lanes[h->order[kf].first];
                const RxGen &G = ln->g[h->cur_gen];
                const FrameDesc &d = G.h_frames.p[h->order[kf].second];
                const unsigned gs = h->global_stream(ln->lane, d.stream);
                if (gs >= nb) continue;
                lqb_rx_s::OwnedFrame of;
                of.pool = pool0 + (int)h->order[kf].first; of.pay_off = 0; of.sym_off = 0; of.has_payload = false; of.has_syms = false;
                // (the conversion lqb_rx_poll does, with the payload / points left as offsets into the pool)
                {
                    lqb_frame_result &r = of.r;
                    std::memset(&r, 0, sizeof r);
                    r.stream = 0; r.seq = d.seq; r.sample_index = d.F;
                    std::memcpy(r.header, d.header, 20);
                    r.header_valid = d.header_valid; r.payload_valid = d.payload_valid; r.payload_len = d.payload_len;
                    if (d.header_valid && !(d.flags & 1u)) {
                        of.pay_off = d.pay_off; of.has_payload = true;
                        if (!(h->flags & LQB_RX_NO_FRAMESYMS)) { of.sym_off = d.sym_off; of.has_syms = true; }
                        r.num_framesyms = d.n_sym;
                    }
                    r.mod_scheme = d.ms; r.mod_bps = d.bps; r.check = d.check; r.fec0 = d.fec0; r.fec1 = d.fec1;
                    r.evm = d.evm; r.rssi = d.rssi; r.cfo = d.cfo;
                    r.tau_hat = d.tau; r.gamma_hat = d.gamma; r.dphi_hat = d.dphi; r.phi_hat = d.phi; r.rxy = d.rxy;
                    r.flags = d.flags;
                }
                of.trig_w = d.F - (long long)d.det_idx;
                // where k_seek leaves the walk behind this frame (lqb_rx_seek.cu: st.G = last + 1, st.wstart = st.G - 256)
                if (d.flags & 1u) of.after_G = d.F + 512;
                else if (d.header_valid) of.after_G = d.F + 2ll * (308ll + (long long)d.n_sym) - (long long)d.tau_neg + 1;
                else of.after_G = d.F + 616 - (long long)d.tau_neg + 1;
                of.after_w = of.after_G - 256;
                seg[todo[t0 + gs]].fr.push_back(std::move(of));
            }
            t_copy += now() - tc0;
        }
        // partial re-runs: spliced onto the rest of the old run if they left its frame in the same state, else run in full
        for (unsigned k = 1; k < K; ++k) {
            Seg &sg = seg[k];
            if (sg.splice < 0 || sg.todo) continue;
            std::sort(sg.fr.begin(), sg.fr.end(), [](const lqb_rx_s::OwnedFrame &a, const lqb_rx_s::OwnedFrame &b) { return a.r.seq < b.r.seq; });
            const lqb_rx_s::OwnedFrame &jf = sg.old[(size_t)sg.splice];
            if (sg.exit_w == jf.after_w && sg.exit_G == norm_G(jf.after_w, jf.after_G)) {
                for (size_t q = (size_t)sg.splice + 1; q < sg.old.size(); ++q) sg.fr.push_back(std::move(sg.old[q]));
                for (size_t q = 0; q < sg.fr.size(); ++q) sg.fr[q].r.seq = (uint32_t)q;
                sg.exit_w = sg.old_exit_w; sg.exit_G = sg.old_exit_G;
                sg.splice = -1; sg.old.clear();
            } else {
                sg.splice = -1; sg.old.clear();
                sg.ran = false;                              // its frames stop short: the next run of this segment is a full one
            }
        }
        // accept runs whose entry state is the accepted exit state before them; schedule the others from that state
        bool chain = true;
        for (unsigned k = 1; k < K; ++k) {
            Seg &sg = seg[k];
            seg[0].fin = true;
            const long long want_w = seg[k - 1].exit_w, want_G = seg[k - 1].exit_G;
            if (!seg[k - 1].fin) chain = false;
            if (want_w == kNoStop || want_w < (long long)k * L) {
                // the walk before never reaches this segment (the capture ends first): nothing of k belongs to the result
                if (chain) { sg.fin = true; sg.todo = false; sg.fr.clear(); sg.exit_w = kNoStop; sg.exit_G = kNoStop; }
                continue;
            }
            if (sg.ran && sg.entry_w == want_w && sg.entry_G == want_G) { if (chain) sg.fin = true; }
            else if (!sg.fin) {
                sg.start_w = want_w; sg.start_G = want_G; sg.todo = true;
                // the old run's first frame the true walk can still reach: go that far only, then try to splice
                if (sg.splice == -1 && sg.ran) {
                    std::sort(sg.fr.begin(), sg.fr.end(), [](const lqb_rx_s::OwnedFrame &a, const lqb_rx_s::OwnedFrame &b) { return a.r.seq < b.r.seq; });
                    int j = -1;
                    for (size_t q = 0; q < sg.fr.size(); ++q)
                        if (sg.fr[q].trig_w >= (long long)k * L && (long long)sg.fr[q].r.sample_index > std::max(want_w, want_G) + 512) { j = (int)q; break; }
                    if (j >= 0) {
                        sg.old = std::move(sg.fr); sg.fr.clear();
                        sg.old_exit_w = sg.exit_w; sg.old_exit_G = sg.exit_G;
                        sg.splice = j; sg.partial_stop = (long long)sg.old[(size_t)j].r.sample_index + 1;
                    }
                }
            }
        }
        if (K == 1) seg[0].fin = true;
    }
    // the sequential list: every accepted run's frames whose triggering window starts in the run's own segment
    for (unsigned k = 0; k < K; ++k) {
        Seg &sg = seg[k];
        std::sort(sg.fr.begin(), sg.fr.end(), [](const lqb_rx_s::OwnedFrame &a, const lqb_rx_s::OwnedFrame &b) { return a.r.seq < b.r.seq; });
        const long long lo = k ? (long long)k * L : -256;
        for (auto &of : sg.fr) {
            if (of.trig_w < lo) continue;                    // found during the pre-roll: the segment before owns it
            of.r.stream = 0; of.r.seq = (uint32_t)h->merged.size();
            h->merged_valid += (of.r.header_valid && of.r.payload_valid) ? 1u : 0u;
            h->merged.push_back(std::move(of));
        }
    }
    for (auto &of : h->merged) {                             // (pointers into the vectors as they finally lie)
        const lqb_rx_s::ShardPool &sp = *h->shard_pools[(size_t)of.pool];
        of.r.payload = of.has_payload ? sp.pay.data() + of.pay_off : nullptr;
        of.r.framesyms = of.has_syms ? reinterpret_cast<const float *>(sp.syms.data() + of.sym_off) : nullptr;
    }
    h->use_merged = true;
    if (shard_trace) fprintf(stderr, "[lqb sharded] total %.1f ms: execute %.1f, states %.1f, frame copies %.1f, rest (verify, merge) %.1f\n",
                             now() - t_begin, t_exec, t_state, t_copy, now() - t_begin - t_exec - t_state - t_copy);
    return 0;
}

int lqb_rx_last_shard_info(lqb_rx h, uint64_t out[4])
{
    if (!h || !out) return fail(LQB_EINVAL, "null handle");
    std::memcpy(out, h->shard, sizeof h->shard);
    return 0;
}

int lqb_rx_poll(lqb_rx h, lqb_frame_result *out, uint32_t max_out, uint32_t *n_out)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (h->use_merged) {
        const unsigned nm = std::min<unsigned>((unsigned)h->merged.size(), max_out);
        for (unsigned k = 0; k < nm && out; ++k) out[k] = h->merged[k].r;
        if (n_out) *n_out = (uint32_t)h->merged.size();
        return 0;
    }
    unsigned n = std::min<unsigned>(h->n_frames, max_out);
    const bool host_res = !(h->flags & LQB_RX_DEVICE_RESULTS);
    for (unsigned k = 0; k < n && out; ++k) {
        const RxLane *ln = h->lanes[h->order[k].first];
        const RxGen &G = ln->g[h->cur_gen];
        const FrameDesc &d = G.h_frames.p[h->order[k].second];
        lqb_frame_result &r = out[k];
        std::memset(&r, 0, sizeof r);
        r.stream = h->global_stream(ln->lane, d.stream); r.seq = d.seq; r.sample_index = d.F;
        std::memcpy(r.header, d.header, 20);
        r.header_valid = d.header_valid; r.payload_valid = d.payload_valid; r.payload_len = d.payload_len;
        if (d.header_valid && !(d.flags & 1u)) {
            r.payload = host_res ? G.h_payload.p + d.pay_off : G.d_payload.p + d.pay_off;
            if (h->flags & LQB_RX_DEVICE_RESULTS) r.framesyms = reinterpret_cast<const float *>(G.d_syms.p + d.sym_off);
            else if (!(h->flags & LQB_RX_NO_FRAMESYMS)) r.framesyms = reinterpret_cast<const float *>(G.h_syms.p + d.sym_off);
            r.num_framesyms = d.n_sym;
        }
        r.mod_scheme = d.ms; r.mod_bps = d.bps; r.check = d.check; r.fec0 = d.fec0; r.fec1 = d.fec1;
        r.evm = d.evm; r.rssi = d.rssi; r.cfo = d.cfo;
        r.tau_hat = d.tau; r.gamma_hat = d.gamma; r.dphi_hat = d.dphi; r.phi_hat = d.phi; r.rxy = d.rxy;
        r.flags = d.flags;
    }
    if (n_out) *n_out = h->n_frames;
    return 0;
}

int lqb_rx_counts(lqb_rx h, uint64_t *frames, uint64_t *valid)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (frames) *frames = h->use_merged ? h->merged.size() : h->n_frames;
    if (valid) *valid = h->use_merged ? h->merged_valid : h->n_valid;
    return 0;
}
int lqb_rx_last_timing(lqb_rx h, float ms[6])
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    std::memcpy(ms, h->ms, sizeof h->ms);
    return 0;
}
int lqb_rx_last_work(lqb_rx h, uint64_t w[6])
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    std::memcpy(w, h->work, sizeof h->work);
    return 0;
}
int lqb_rx_last_search_bins(lqb_rx h, uint64_t *bins)
{
    if (!h || !bins) return fail(LQB_EINVAL, "null handle");
    *bins = h->bins;
    return 0;
}
int lqb_rx_launch_count(lqb_rx h, uint64_t *l)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    uint64_t t = 0;
    for (auto *ln : h->lanes) t += ln->f.launches;
    *l = t;
    return 0;
}
int lqb_rx_lane_count(lqb_rx h) { return h ? (int)h->lanes.size() : fail(LQB_EINVAL, "null handle"); }

}  // extern "C"

// =================================================================== detector handle
struct lqb_det_s {
    Front f;
    DevBuf<Detection> d_det;
    PinBuf<Detection> h_det;
    std::vector<unsigned> order;
    unsigned n_det = 0;
    uint64_t windows = 0;
    uint64_t search[4] = {};              // windows, alignments, exact window evaluations, CFO bins those visited
    cudaEvent_t ev[2] = {};
    float ms = 0.0f;
    // time-sharded search of one capture (lqb_det_execute_sharded)
    DevBuf<float2> d_capture;             // host captures are staged here
    std::vector<Detection> merged;        // the sequential detector's list, in order
    bool use_merged = false;
    uint64_t shard[4] = {};               // segments, segment runs in all, rounds, launches
};

extern "C" {

lqb_det lqb_det_create(const lqb_det_opts *o)
{
    if (!o || !o->n_streams) { fail(LQB_EINVAL, "bad options"); return nullptr; }
    DevTables *T = new DevTables;
    build_tables(*T, o->beta > 0.0f ? o->beta : 0.3f, o->threshold > 0.0f ? o->threshold : 0.45f, o->dphi_max > 0.0f ? o->dphi_max : 0.3f);
    lqb_det h = new lqb_det_s;
    int e = h->f.init(o->device, o->n_streams, 2048, o->cuda_stream, *T);
    delete T;
    if (e) { h->f.destroy(); delete h; return nullptr; }
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    return h;
}
void lqb_det_destroy(lqb_det h)
{
    if (!h) return;
    cudaSetDevice(h->f.device);
    if (h->f.stream) cudaStreamSynchronize(h->f.stream);
    h->d_det.release(); h->h_det.release(); h->d_capture.release();
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    h->f.destroy();
    delete h;
}
int lqb_det_reset(lqb_det h, int s) { return h ? h->f.reset(s) : fail(LQB_EINVAL, "null handle"); }

int lqb_det_execute(lqb_det h, uint32_t n, const uint32_t *ids, const float *const *iq, const uint64_t *ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    Front &f = h->f;
    CU(cudaSetDevice(f.device));
    cudaStream_t st = f.stream;
    h->n_det = 0; h->order.clear(); h->use_merged = false;
    uint64_t total = 0, max_n = 0;
    if (int e = f.feed(n, ids, iq, ns, mem, &total, &max_n)) return e;
    if (!n) return 0;
    size_t max_det = 0;
    for (uint32_t i = 0; i < n; ++i) max_det += (size_t)((ns[i] + f.carry_cap) / 256 + 2);
    if (int e = h->d_det.reserve(max_det)) return e;
    if (int e = h->h_det.reserve(max_det)) return e;
    SeekParams sp;
    sp.tables = f.d_tables; sp.states = f.d_states; sp.io = f.io[0].d_io.p;
    sp.carry[0] = f.d_carry[0]; sp.carry[1] = f.d_carry[1]; sp.carry_cap = f.carry_cap;
    sp.det_mode = 1; sp.frames = nullptr; sp.detections = h->d_det.p; sp.views = nullptr;
    sp.n_out = f.io[0].d_count; sp.max_out = (unsigned)max_det;
    CU(cudaEventRecord(h->ev[0], st));
    f.set_coarse(sp);
    if (int e = f.set_slices(sp, n, f.io[0].h_io.p)) return e;
    CU(cudaMemsetAsync(f.io[0].d_count, 0, 8 * sizeof(unsigned), st));
    launch_seek(sp, n, st); f.launches += sp.slice_len ? 2 : 1;
    launch_carry(sp, n, st); f.launches++;
    CU(cudaEventRecord(h->ev[1], st));
    CU(cudaMemcpyAsync(f.io[0].h_count, f.io[0].d_count, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    h->windows = f.io[0].h_count[1];
    h->search[0] = f.io[0].h_count[1]; h->search[1] = f.io[0].h_count[2]; h->search[2] = f.io[0].h_count[3]; h->search[3] = f.io[0].h_count[5];
    unsigned nd = std::min<unsigned>(f.io[0].h_count[0], (unsigned)max_det);
    if (nd) {
        CU(cudaMemcpyAsync(h->h_det.p, h->d_det.p, nd * sizeof(Detection), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    CU(cudaGetLastError());
    cudaEventElapsedTime(&h->ms, h->ev[0], h->ev[1]);
    h->n_det = nd;
    h->order.resize(nd);
    for (unsigned i = 0; i < nd; ++i) h->order[i] = i;
    const Detection *dd = h->h_det.p;
    std::sort(h->order.begin(), h->order.end(), [&](unsigned a, unsigned b) {
        return dd[a].stream != dd[b].stream ? dd[a].stream < dd[b].stream : dd[a].seq < dd[b].seq;
    });
    return 0;
}

int lqb_det_execute_dense(lqb_det h, const float *iq, uint64_t stride, uint64_t ns, int mem)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    unsigned n = h->f.n_streams;
    std::vector<const float *> ptr(n);
    std::vector<uint64_t> len(n, ns);
    for (unsigned s = 0; s < n; ++s) ptr[s] = iq + (mem >= LQB_MEM_HOST_SC16 ? 1 : 2) * (size_t)s * stride;   // a sample is 2 floats, or 1 float-sized int16 pair
    return lqb_det_execute(h, n, nullptr, ptr.data(), len.data(), mem);
}

// One capture, searched as the sequential detector would search it from its reset state, but cut in TIME into segments
// that run side by side as streams of this handle (the reference's qdetector is strictly serial: its hop grid re-phases
// at every detection, lib/frame_detector_cc_impl.cc:77).  Exactness does not rest on the cut:
//   * a segment k >= 1 is first run SPECULATIVELY from `preroll` samples before its boundary b_k = k seg_len, on the
//     grid that is right when nothing was detected so far; any two walks that detect the same frame are in the same
//     state from there on, so after a pre-roll that holds a frame the walk usually IS the sequential one;
//   * every run records the first window start >= b_k it reaches (its ENTRY) and the first one >= b_(k+1) (its EXIT,
//     where it stops).  Segment 0 starts from the true reset state.  A run is accepted only if its entry equals the
//     accepted exit of the segment before it; otherwise the segment is run again from exactly that state.  The walk is a
//     function of its state and of the samples at or after it, so every accepted run is a piece of the sequential walk.
// Results (lqb_det_poll): stream 0, seq = order of detection, sample_index absolute -- the list qdetector_cccf returns
// for the whole capture (windows that need samples beyond its end are not evaluated, as in a stream that has not ended).
// The handle's stream states are consumed: reset before going back to lqb_det_execute.
int lqb_det_execute_sharded(lqb_det h, const float *iq, uint64_t n_samples, int mem, uint32_t seg_len, uint32_t preroll)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    if (mem != LQB_MEM_HOST && mem != LQB_MEM_DEVICE) return fail(LQB_EINVAL, "complex64 input only");
    if (!iq && n_samples) return fail(LQB_EINVAL, "null sample pointer");
    Front &f = h->f;
    CU(cudaSetDevice(f.device));
    cudaStream_t st = f.stream;
    if (!seg_len) seg_len = 1u << 18;
    seg_len = std::max(2048u, (seg_len + 255u) & ~255u);
    preroll = std::min((preroll + 255u) & ~255u, seg_len);
    h->n_det = 0; h->order.clear(); h->merged.clear(); h->use_merged = true;
    std::memset(h->shard, 0, sizeof h->shard);
    h->windows = 0; std::memset(h->search, 0, sizeof h->search);
    if (n_samples < 512) return 0;
    const float2 *x = reinterpret_cast<const float2 *>(iq);
    CU(cudaEventRecord(h->ev[0], st));
    if (mem == LQB_MEM_HOST) {
        if (int e = h->d_capture.reserve(n_samples + 2)) return e;
        CU(cudaMemcpyAsync(h->d_capture.p, iq, n_samples * sizeof(float2), cudaMemcpyHostToDevice, st));
        x = h->d_capture.p;
    }
    const long long N = (long long)n_samples, L = seg_len;
    const unsigned K = (unsigned)((N + L - 1) / L);
    struct Seg {
        long long start = 0;        // window start the current / next run begins at
        long long entry = 0, exit = 0;
        bool ran = false, fin = false, todo = true;
        std::vector<Detection> det;
        // a re-run that only goes as far as the first detection of the run before it (old[splice]) and is spliced onto
        // the rest of that run when it makes the same detection (the walk's state behind a detection at F is F + 256)
        std::vector<Detection> old;
        long long old_exit = 0, partial_stop = 0;
        int splice = -1;
    };
    std::vector<Seg> seg(K);
    for (unsigned k = 0; k < K; ++k) seg[k].start = k ? std::max<long long>(0, (long long)k * L - (long long)preroll) : -256;
    h->shard[0] = K;
    const unsigned W = f.n_streams;
    std::vector<unsigned> batch;
    std::vector<const float *> ptr;
    std::vector<uint64_t> len;
    while (true) {
        // ---- run every segment that needs it, W at a time
        std::vector<unsigned> todo;
        for (unsigned k = 0; k < K; ++k) if (seg[k].todo) todo.push_back(k);
        if (todo.empty()) break;
        h->shard[2]++;
        for (size_t t0 = 0; t0 < todo.size(); t0 += W) {
            const unsigned nb = (unsigned)std::min<size_t>(W, todo.size() - t0);
            ptr.resize(nb); len.resize(nb);
            size_t max_det = 0;
            for (unsigned i = 0; i < nb; ++i) {
                const unsigned k = todo[t0 + i];
                Seg &sg = seg[k];
                // samples from 256 before the first window (what the sequential walk has behind it there) up to what a
                // frame found by the segment's last window needs: start + 355 + 512 < b_(k+1) + 1024
                const long long base = std::max<long long>(0, sg.start - 256);
                const long long end = std::min<long long>(N, (long long)(k + 1) * L + 1024);
                StreamState z;
                std::memset(&z, 0, sizeof z);
                z.base = base; z.G = base; z.wstart = sg.start;
                z.stop_at = (k + 1 < K) ? (long long)(k + 1) * L : kNoStop;
                if (sg.splice >= 0) z.stop_at = std::min(z.stop_at, sg.partial_stop);
                z.mark_at = k ? (long long)k * L : -256;
                z.mark_w = kNoMark;
                f.h_states[i] = z;
                ptr[i] = reinterpret_cast<const float *>(x + base);
                len[i] = (uint64_t)std::max<long long>(0, end - base);
                max_det += (size_t)(len[i] / 256 + 4);
            }
            CU(cudaMemcpyAsync(f.d_states, f.h_states, nb * sizeof(StreamState), cudaMemcpyHostToDevice, st));
            uint64_t total = 0, max_n = 0;
            const bool lpt = f.lpt;
            f.lpt = false;                                   // (segments cost about the same; keep the list in order)
            const int fe = f.feed(nb, nullptr, ptr.data(), len.data(), LQB_MEM_DEVICE, &total, &max_n);
            f.lpt = lpt;
            if (fe) return fe;
            if (int e = h->d_det.reserve(max_det)) return e;
            if (int e = h->h_det.reserve(max_det)) return e;
            SeekParams sp;
            sp.tables = f.d_tables; sp.states = f.d_states; sp.io = f.io[0].d_io.p;
            sp.carry[0] = f.d_carry[0]; sp.carry[1] = f.d_carry[1]; sp.carry_cap = f.carry_cap;
            sp.det_mode = 1; sp.frames = nullptr; sp.detections = h->d_det.p; sp.views = nullptr;
            sp.n_out = f.io[0].d_count; sp.max_out = (unsigned)max_det;
            f.set_coarse(sp);
            CU(cudaMemsetAsync(f.io[0].d_count, 0, 8 * sizeof(unsigned), st));
            launch_seek(sp, nb, st); f.launches++; h->shard[3]++;
            CU(cudaMemcpyAsync(f.io[0].h_count, f.io[0].d_count, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(f.h_states, f.d_states, nb * sizeof(StreamState), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const unsigned nd = std::min<unsigned>(f.io[0].h_count[0], (unsigned)max_det);
            if (nd) {
                CU(cudaMemcpyAsync(h->h_det.p, h->d_det.p, nd * sizeof(Detection), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
            }
            CU(cudaGetLastError());
            h->windows += f.io[0].h_count[1];
            h->search[0] += f.io[0].h_count[1]; h->search[1] += f.io[0].h_count[2]; h->search[2] += f.io[0].h_count[3]; h->search[3] += f.io[0].h_count[5];
            for (unsigned i = 0; i < nb; ++i) {
                Seg &sg = seg[todo[t0 + i]];
                const StreamState &z = f.h_states[i];
                sg.ran = true; sg.todo = false; sg.det.clear();
                // a walk that ends with a frame start it cannot align yet (the capture ends there) has no further state
                sg.exit = (z.mode == 0) ? z.wstart : kNoStop;
                sg.entry = (z.mark_w == kNoMark) ? kNoStop : z.mark_w;        // (never reached its own boundary: the capture ended)
                h->shard[1]++;
            }
            for (unsigned j = 0; j < nd; ++j) {
                const Detection &d = h->h_det.p[j];
                if (d.stream < nb) seg[todo[t0 + d.stream]].det.push_back(d);
            }
        }
        // ---- partial re-runs: spliced onto the rest of the old run if they made its detection, else run in full
        auto by_seq = [](const Detection &a, const Detection &b) { return a.seq < b.seq; };
        for (unsigned k = 1; k < K; ++k) {
            Seg &sg = seg[k];
            if (sg.splice < 0 || sg.todo) continue;
            std::sort(sg.det.begin(), sg.det.end(), by_seq);
            if (sg.exit == sg.old[(size_t)sg.splice].F + 256) {
                for (size_t q = (size_t)sg.splice + 1; q < sg.old.size(); ++q) sg.det.push_back(sg.old[q]);
                for (size_t q = 0; q < sg.det.size(); ++q) sg.det[q].seq = (unsigned)q;
                sg.exit = sg.old_exit;
            } else sg.ran = false;                           // its list stops short: the next run of this segment is a full one
            sg.splice = -1; sg.old.clear();
        }
        // ---- accept runs whose entry is the accepted exit before them; schedule the others from that exit
        bool chain = true;                                   // everything before k is accepted
        for (unsigned k = 0; k < K; ++k) {
            Seg &sg = seg[k];
            if (k == 0) { sg.fin = true; continue; }
            const long long want = seg[k - 1].exit;          // accepted (chain) or merely the best knowledge so far
            if (want == kNoStop || want < (long long)k * L) {
                // the walk before never reaches this segment (it ran out of samples, or waits for the rest of a frame,
                // at the end of the capture): nothing of k belongs to the result
                if (chain) { sg.fin = true; sg.todo = false; sg.det.clear(); sg.exit = kNoStop; }
                else chain = false;
                continue;
            }
            if (sg.ran && sg.entry == want) { if (chain) sg.fin = true; }
            else {
                // the run has to start exactly in that state (a window start at or beyond the boundary)
                if (!sg.fin) {
                    sg.start = want; sg.todo = true;
                    // the old run's first detection the true walk can still make: go that far only, then try to splice
                    if (sg.ran) {
                        std::sort(sg.det.begin(), sg.det.end(), by_seq);
                        int j = -1;
                        for (size_t q = 0; q < sg.det.size(); ++q)
                            if (sg.det[q].F - (long long)sg.det[q].pad >= (long long)k * L && sg.det[q].F > want + 512) { j = (int)q; break; }
                        if (j >= 0) {
                            sg.old = std::move(sg.det); sg.det.clear();
                            sg.old_exit = sg.exit; sg.splice = j; sg.partial_stop = sg.old[(size_t)j].F + 1;
                        }
                    }
                }
                chain = false;
            }
            if (!sg.fin) chain = false;
        }
    }
    // ---- the sequential list: each accepted run's detections from windows at or beyond its entry, in walk order
    for (unsigned k = 0; k < K; ++k) {
        Seg &sg = seg[k];
        std::sort(sg.det.begin(), sg.det.end(), [](const Detection &a, const Detection &b) { return a.seq < b.seq; });
        const long long lo = k ? (long long)k * L : -256;
        for (const Detection &d : sg.det) {
            if (d.F - (long long)d.pad < lo) continue;       // found during the pre-roll: the segment before owns it
            Detection o = d;
            o.stream = 0; o.seq = (unsigned)h->merged.size(); o.pad = 0;
            h->merged.push_back(o);
        }
    }
    CU(cudaEventRecord(h->ev[1], st));
    CU(cudaEventSynchronize(h->ev[1]));
    cudaEventElapsedTime(&h->ms, h->ev[0], h->ev[1]);
    h->n_det = (unsigned)h->merged.size();
    return 0;
}

int lqb_det_last_shard_info(lqb_det h, uint64_t out[4])
{
    if (!h || !out) return fail(LQB_EINVAL, "null handle");
    std::memcpy(out, h->shard, sizeof h->shard);
    return 0;
}

int lqb_det_poll(lqb_det h, lqb_detection *out, uint32_t max_out, uint32_t *n_out)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    unsigned n = std::min<unsigned>(h->n_det, max_out);
    for (unsigned k = 0; k < n && out; ++k) {
        const Detection &d = h->use_merged ? h->merged[k] : h->h_det.p[h->order[k]];
        out[k].stream = d.stream; out[k].seq = d.seq; out[k].sample_index = d.F;
        out[k].tau_hat = d.tau; out[k].gamma_hat = d.gamma; out[k].dphi_hat = d.dphi; out[k].phi_hat = d.phi; out[k].rxy = d.rxy;
    }
    if (n_out) *n_out = h->n_det;
    return 0;
}
int lqb_det_last_timing(lqb_det h, float *ms)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    *ms = h->ms;
    return 0;
}
int lqb_det_last_search(lqb_det h, uint64_t out[4])
{
    if (!h || !out) return fail(LQB_EINVAL, "null handle");
    std::memcpy(out, h->search, sizeof h->search);
    return 0;
}
int lqb_det_last_work(lqb_det h, uint64_t *windows)
{
    if (!h) return fail(LQB_EINVAL, "null handle");
    *windows = h->windows;
    return 0;
}

// =================================================================== host-side tables (no GPU)
int lqb_tab_interp_taps(float beta, float *h30) { auto h = interp_taps(beta); std::memcpy(h30, h.data(), 30 * sizeof(float)); return 0; }
int lqb_tab_pfb_banks(float beta, float *b) { auto v = pfb_banks(beta); std::memcpy(b, v.data(), v.size() * sizeof(float)); return 0; }
int lqb_tab_detector_template(float beta, float *s) { auto v = detector_template(beta); std::memcpy(s, v.data(), v.size() * sizeof(cf)); return 0; }
int lqb_tab_nco_sintab(float *t) { std::memcpy(t, nco_sintab(), 1024 * sizeof(float)); return 0; }
int lqb_tab_secded_columns(uint32_t data_bytes, uint8_t *col)
{
    if (!col || (data_bytes != 2 && data_bytes != 4 && data_bytes != 8)) return fail(LQB_EINVAL, "data_bytes is 2, 4 or 8");
    uint8_t c[3][64];
    secded_cols(c);
    std::memcpy(col, c[data_bytes == 2 ? 0 : data_bytes == 4 ? 1 : 2], 8 * data_bytes);
    return 0;
}
int lqb_tab_ilv_bit_perm(uint32_t n, uint32_t *perm)
{
    if (!perm) return fail(LQB_EINVAL, "null pointer");
    auto v = ilv_bit_perm(n);
    std::memcpy(perm, v.data(), v.size() * sizeof(uint32_t));
    return 0;
}
int lqb_tab_packet_len(uint32_t n, uint32_t check, uint32_t fec0, uint32_t fec1, uint32_t ms, uint32_t *enc, uint32_t *nsym)
{
    if (!modem_supported(ms) || !fec_supported(fec0) || !fec_supported(fec1) || check == 0 || check >= CRC_NUM) return fail(LQB_EINVAL, "unsupported scheme");
    if (enc) *enc = packetizer_enc_len(n, check, fec0, fec1);
    if (nsym) *nsym = qpm_frame_len(n, check, fec0, fec1, ms);
    return 0;
}

}  // extern "C"
