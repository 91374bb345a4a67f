/*
 * lqb200.h -- C-ABI of the B200-native packet PHY (liblqb200.so).
 *
 * This is the drop-in boundary for the one hot path of gvanhoy/gr-liquiddsp:
 * the liquid-dsp calls made by its three GNU Radio blocks.  Each entry point
 * below names the reference call site it replaces (paths relative to the
 * reference repository root):
 *
 *   lqb_rx_*   : flexframesync_create / _execute / _destroy and the
 *                framesync_callback + framesyncstats_s contract
 *                lib/flex_rx_impl.cc:49 (create), :71 (destroy), :213 (execute),
 *                :182-201 (callback), lib/flex_rx_impl.h:27-37 (packet_info)
 *   lqb_tx_*   : flexframegenprops_init_default / flexframegen_create /
 *                _setprops / _assemble / _getframelen / _write_samples / _destroy
 *                lib/flex_tx_impl.cc:51-56, :72, :188, :198-201
 *   lqb_det_*  : msequence_create/advance/destroy, qdetector_cccf_create_linear /
 *                _set_threshold / _execute / _destroy (and the commented-out getters)
 *                lib/frame_detector_cc_impl.cc:47-55, :63, :77, :90-93
 *
 * Conventions
 *   - plain C, opaque handles, no exceptions, never exit(): create() returns NULL on
 *     failure (see lqb_last_error()); every other call returns 0 or a negative LQB_E* code.
 *   - samples are interleaved complex64 (re, im float pairs) = gr_complex.
 *   - a handle is single-threaded (GNU Radio's thread-per-block rule); distinct handles
 *     may be driven from distinct host threads and distinct GPUs.
 *   - "mem" says where caller buffers live: LQB_MEM_HOST (pageable or pinned host
 *     memory; the library stages through pinned buffers) or LQB_MEM_DEVICE (device
 *     pointers on the handle's GPU; zero-copy, work is enqueued on the handle's stream).
 *   - result buffers returned by *_poll are owned by the handle and stay valid until
 *     the next *_execute / *_poll / *_destroy on that handle (the lifetime rule the
 *     reference relies on at lib/flex_rx_impl.cc:192-198).
 *   - there is no CPU fallback: every compute call fails with LQB_ENODEV when no
 *     sm_100 device is usable.
 */
#ifndef LQB200_H
#define LQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQB_VERSION 100

/* error codes */
#define LQB_OK        0
#define LQB_EINVAL   (-22)
#define LQB_ENOMEM   (-12)
#define LQB_ENODEV   (-19)
#define LQB_ECUDA    (-5)
#define LQB_ERANGE   (-34)
#define LQB_EBUSY    (-16)

#define LQB_MEM_HOST   0
#define LQB_MEM_DEVICE 1
/* Additive input formats for the receiver and the detector (the reference's blocks only know complex64): the iq
 * pointers address interleaved int16 pairs (re, im) -- what SDR front ends deliver natively -- and a sample is
 * value / 32768.  Half the bytes of complex64 over PCIe; widened on the device, results identical to feeding the
 * widened floats. */
#define LQB_MEM_HOST_SC16   2
#define LQB_MEM_DEVICE_SC16 3

/* liquid-dsp wire enums carried in the frame header (only the values the blocks use are named) */
#define LQB_MODEM_PSK2 1
#define LQB_MODEM_PSK4 2
#define LQB_MODEM_PSK8 3
#define LQB_MODEM_PSK16 4
#define LQB_MODEM_DPSK2 9
#define LQB_MODEM_DPSK4 10
#define LQB_MODEM_DPSK8 11
#define LQB_MODEM_ASK4 18
#define LQB_MODEM_QAM16 27
#define LQB_MODEM_QAM32 28
#define LQB_MODEM_QAM64 29
#define LQB_MODEM_QAM128 30
#define LQB_MODEM_QAM256 31
#define LQB_MODEM_BPSK 39
#define LQB_MODEM_QPSK 40
#define LQB_FEC_NONE 1
#define LQB_FEC_REP3 2
#define LQB_FEC_REP5 3
#define LQB_FEC_HAMMING74 4
#define LQB_FEC_HAMMING84 5
#define LQB_FEC_HAMMING128 6
#define LQB_FEC_GOLAY2412 7
#define LQB_FEC_SECDED2216 8
#define LQB_FEC_SECDED3932 9
#define LQB_FEC_SECDED7264 10
#define LQB_FEC_CONV_V27 11
#define LQB_FEC_CONV_V29 12
#define LQB_FEC_CONV_V27P23 15
#define LQB_FEC_CONV_V27P34 16
#define LQB_FEC_CONV_V27P45 17
#define LQB_FEC_CONV_V27P56 18
#define LQB_FEC_CONV_V27P67 19
#define LQB_FEC_CONV_V27P78 20
#define LQB_FEC_CONV_V29P23 21
#define LQB_FEC_CONV_V29P34 22
#define LQB_FEC_CONV_V29P45 23
#define LQB_FEC_CONV_V29P56 24
#define LQB_FEC_CONV_V29P67 25
#define LQB_FEC_CONV_V29P78 26
#define LQB_FEC_RS_M8 27
#define LQB_CRC_NONE 1
#define LQB_CRC_CHECKSUM 2
#define LQB_CRC_8 3
#define LQB_CRC_16 4
#define LQB_CRC_24 5
#define LQB_CRC_32 6

const char *lqb_last_error(void);          /* thread-local message of the last failure */
int  lqb_device_count(void);               /* number of usable CUDA devices (0 without a GPU) */
int  lqb_version(void);

/* ------------------------------------------------------------------ RX (flex_rx / flexframesync) */
typedef struct lqb_rx_s *lqb_rx;

#define LQB_RX_NO_FRAMESYMS  1u   /* do not copy payload constellation points back to the host */
#define LQB_RX_DEVICE_RESULTS 2u  /* keep payload bytes on the device too (descriptors only are copied) */
#define LQB_RX_SOFT 4u            /* opt-in extension (not liquid-dsp's default, which the blocks use): frames whose coding
                                     stage nearest the channel (fec1, or fec0 when fec1 is "none") is convolutional are
                                     decoded from SOFT decisions -- per coded bit clamp(128 + G (d0 - d1)) from the squared
                                     distances to the nearest constellation points with that bit 0 / 1, G = 64 / dmin^2 --
                                     through a soft-input Viterbi decoder (about 2 dB at the same PER).  DPSK frames and
                                     all other coding chains are decoded as without the flag.  Header decoding, estimates
                                     and constellation points do not change. */

typedef struct {
    int      device;             /* CUDA device ordinal */
    uint32_t n_streams;          /* independent channel streams in this batch */
    uint32_t max_frame_samples;  /* per-stream carry capacity; 0 = 65536. Frames longer than this are dropped */
    uint32_t flags;              /* LQB_RX_* */
    void    *cuda_stream;        /* cudaStream_t the caller works on (inputs are ordered after it, it is ordered
                                    after the results); NULL = library-owned streams only */
    uint32_t n_lanes;            /* streams are split over this many independent pipeline lanes (a fixed, interleaved
                                    partition) whose search, payload kernels and copies overlap;
                                    0 = automatic (env LQB_RX_LANES, else one per 256 streams, at most 4).  Results
                                    do not depend on it. */
} lqb_rx_opts;

/* mirrors framesync_callback's arguments + framesyncstats_s (lib/flex_rx_impl.cc:182-201),
 * plus the estimates liquid keeps private (tau, gamma, rxy) and the frame's position */
typedef struct {
    uint32_t stream;             /* which channel */
    uint32_t seq;                /* frame ordinal within that stream since create/reset */
    int64_t  sample_index;       /* absolute index (per stream) of the first sample of the frame */
    uint8_t  header[20];         /* 14 user bytes + 6 protocol bytes, as decoded */
    int32_t  header_valid;
    int32_t  payload_valid;
    uint32_t payload_len;
    const uint8_t *payload;      /* payload_len bytes (NULL when header invalid) */
    const float   *framesyms;    /* num_framesyms complex points (NULL with LQB_RX_NO_FRAMESYMS / header invalid) */
    uint32_t num_framesyms;
    uint32_t mod_scheme, mod_bps, check, fec0, fec1;
    float    evm;                /* dB */
    float    rssi;               /* dB */
    float    cfo;                /* rad/sample */
    float    tau_hat, gamma_hat, dphi_hat, phi_hat, rxy;
    uint32_t flags;              /* bit0: frame dropped -- its header decoded but the frame is longer than
                                    max_frame_samples: header / scheme fields are valid, payload and framesyms are NULL,
                                    payload_valid is 0 and the search resumed 512 samples after the frame start */
} lqb_frame_result;

lqb_rx lqb_rx_create(const lqb_rx_opts *opts);
void   lqb_rx_destroy(lqb_rx h);
int    lqb_rx_reset(lqb_rx h, int stream /* -1 = all */);
/* Feed n_samples[i] new samples to stream stream_ids[i] (each stream at most once per call).
 * iq[i] points at interleaved complex64 (interleaved int16 pairs cast to float* for the *_SC16 kinds).  Runs the whole receive chain for every frame that
 * completes inside the data seen so far; partial frames are carried to the next call. */
int    lqb_rx_execute(lqb_rx h, uint32_t n, const uint32_t *stream_ids,
                      const float *const *iq, const uint64_t *n_samples, int mem);
/* Dense form: all n_streams streams, stream s at iq + 2*s*stride_samples floats, n_samples each */
int    lqb_rx_execute_dense(lqb_rx h, const float *iq, uint64_t stride_samples, uint64_t n_samples, int mem);
/* Pipelined form of execute for callers that stream batch after batch.  submit() queues the input copies and the
 * search of the new samples and returns without waiting for them; collect() waits for the oldest submitted call and
 * makes its frames current for poll / counts / timing / work.  Up to two calls may be in flight:
 *     submit(k); submit(k+1); collect() -> k; poll...; submit(k+2); collect() -> k+1; ...
 * so the H2D copy of call k+1 runs under the search of call k, and the payload work of call k (matched filter, PLL,
 * FEC, CRC, result copies) under the search of call k+1.  execute() == submit() + collect(); results are identical.
 * Host input buffers may be reused as soon as submit returns only if they are pageable; pinned host buffers and
 * device buffers must stay untouched until the matching collect() returns (collect waits for the last reader).  Result
 * buffers of a collected call stay valid until the next submit / execute on the handle (with two calls in flight the
 * next submit reuses the collected call's buffers).  A third submit without a collect returns LQB_EBUSY. */
int    lqb_rx_submit(lqb_rx h, uint32_t n, const uint32_t *stream_ids,
                     const float *const *iq, const uint64_t *n_samples, int mem);
int    lqb_rx_submit_dense(lqb_rx h, const float *iq, uint64_t stride_samples, uint64_t n_samples, int mem);
int    lqb_rx_collect(lqb_rx h);
/* ONE capture decoded as one flexframesync would decode it from its reset state (reference call site
 * lib/flex_rx_impl.cc:213), cut in time into segments of seg_len samples that run side by side as the streams of this
 * handle (n_streams at a time): the single-stream case at batch speed.  Seam rule as lqb_det_execute_sharded (speculative
 * start `preroll` samples early; a run is accepted only if it entered its segment in exactly the state -- next window
 * start and zero boundary -- the accepted run before it stopped in, otherwise the segment is run again from that
 * state), so frames, bytes, flags and estimates do not depend on the cut.  preroll should span the longest frame plus a
 * gap.  mem: LQB_MEM_HOST or LQB_MEM_DEVICE (complex64).  Results through lqb_rx_poll / lqb_rx_counts: stream 0, seq =
 * order, payload / framesyms owned by the handle until its next call.  Not with LQB_RX_DEVICE_RESULTS.  Consumes the
 * stream states: lqb_rx_reset before going back to lqb_rx_execute.  0 for seg_len: 1048576. */
int lqb_rx_execute_sharded(lqb_rx h, const float *iq, uint64_t n_samples, int mem, uint32_t seg_len, uint32_t preroll);
int lqb_rx_last_shard_info(lqb_rx h, uint64_t out[4]);   /* segments, segment runs in all, rounds, execute calls */
/* Frames completed by the last execute / collect, ordered by (stream, seq). */
int    lqb_rx_poll(lqb_rx h, lqb_frame_result *out, uint32_t max_out, uint32_t *n_out);
/* number of frames completed by the last execute / payloads with a passing check */
int    lqb_rx_counts(lqb_rx h, uint64_t *frames, uint64_t *valid_payloads);
/* per-kernel device times (ms) of the last execute, measured with CUDA events on the streams the
 * kernels are launched on: [0]=pre-filter+seek/align/header [1]=matched filter [2]=PLL+demod
 * [3]=FEC+CRC [4]=total [5]=tensor-core pre-filter alone (included in [0]).  With more than one lane
 * [0..3],[5] are sums over lanes of intervals that overlap each other on the GPU (use n_lanes = 1
 * for an undisturbed per-kernel breakdown) and [4] is the longest lane. */
int    lqb_rx_last_timing(lqb_rx h, float ms[6]);
int    lqb_rx_launch_count(lqb_rx h, uint64_t *launches);   /* kernels launched since create */
int    lqb_rx_lane_count(lqb_rx h);                         /* pipeline lanes of this handle */
/* work done by the last execute: [0] 512-sample detector windows visited, [1] frame alignments,
 * [2] payload symbols matched-filtered/demodulated, [3] input samples consumed,
 * [4] windows that needed the exact 50-FFT evaluation, [5] 128-lag pre-filter tiles */
int    lqb_rx_last_work(lqb_rx h, uint64_t work[6]);
/* CFO bins the exact window evaluations of the last execute visited (49 per window without the tensor-core bin filter) */
int    lqb_rx_last_search_bins(lqb_rx h, uint64_t *bins);

/* ------------------------------------------------------------------ TX (flex_tx / flexframegen) */
typedef struct lqb_tx_s *lqb_tx;
typedef struct {
    int      device;
    uint32_t flags;
    void    *cuda_stream;
} lqb_tx_opts;
typedef struct { uint32_t check, fec0, fec1, mod_scheme; } lqb_tx_props;   /* flexframegenprops_s */

lqb_tx lqb_tx_create(const lqb_tx_opts *opts);
void   lqb_tx_destroy(lqb_tx h);
void   lqb_tx_props_init_default(lqb_tx_props *p);   /* CRC-16, no FEC, QPSK */
/* samples in a frame: 2*(64 + 231 + payload symbols + 14)   (flexframegen_getframelen) */
int    lqb_tx_frame_len(const lqb_tx_props *p, uint32_t payload_len, uint32_t *n_samples);
/* Assemble + write n frames in one launch.  headers[i]: 14 user bytes (NULL = zeros);
 * payloads[i]: payload_lens[i] bytes; out[i]: room for that frame's samples (complex64). */
int    lqb_tx_assemble(lqb_tx h, uint32_t n, const lqb_tx_props *props,
                       const uint8_t *const *headers, const uint8_t *const *payloads,
                       const uint32_t *payload_lens, float *const *out, int mem);
/* Asynchronous form: submit() queues the batch on the handle's stream and returns; with LQB_MEM_DEVICE the frames are
 * complete in stream order (on the caller's stream when the handle was created on one) and further submits may follow
 * without waiting; with LQB_MEM_HOST the frames reach `out` in collect().  collect() waits for everything submitted.
 * assemble() == submit() + collect().  The argument arrays may be reused as soon as submit returns. */
int    lqb_tx_submit(lqb_tx h, uint32_t n, const lqb_tx_props *props,
                     const uint8_t *const *headers, const uint8_t *const *payloads,
                     const uint32_t *payload_lens, float *const *out, int mem);
int    lqb_tx_collect(lqb_tx h);
/* device time (ms) of the frame generator kernel of the last collected batch (CUDA events on the handle's stream) */
int    lqb_tx_last_timing(lqb_tx h, float *kernel_ms);

/* ------------------------------------------------------------------ detector (frame_detector_cc / qdetector_cccf) */
typedef struct lqb_det_s *lqb_det;
typedef struct {
    int      device;
    uint32_t n_streams;
    float    beta;        /* 0 = 0.3   (lib/frame_detector_cc_impl.h:36) */
    float    threshold;   /* 0 = 0.45  (lib/frame_detector_cc_impl.cc:55) */
    float    dphi_max;    /* 0 = 0.3   (qdetector default range) */
    void    *cuda_stream;
} lqb_det_opts;
typedef struct {
    uint32_t stream, seq;
    int64_t  sample_index;
    float    tau_hat, gamma_hat, dphi_hat, phi_hat, rxy;
} lqb_detection;

lqb_det lqb_det_create(const lqb_det_opts *opts);
void    lqb_det_destroy(lqb_det h);
int     lqb_det_reset(lqb_det h, int stream);
int     lqb_det_execute(lqb_det h, uint32_t n, const uint32_t *stream_ids,
                        const float *const *iq, const uint64_t *n_samples, int mem);
int     lqb_det_execute_dense(lqb_det h, const float *iq, uint64_t stride_samples, uint64_t n_samples, int mem);
/* ONE capture searched as the sequential detector would search it from its reset state (the list qdetector_cccf returns
 * for the whole capture; reference call site lib/frame_detector_cc_impl.cc:77), but cut in time into segments of seg_len
 * samples that run side by side as the streams of this handle (n_streams at a time).  A segment is first searched
 * speculatively from `preroll` samples before its boundary; a run is accepted only if it entered its segment in exactly
 * the state the accepted run before it left in, otherwise the segment is searched again from that state -- so the
 * result does not depend on seg_len / preroll, only the time does (preroll should span a frame period).  mem:
 * LQB_MEM_HOST or LQB_MEM_DEVICE (complex64).  Results: lqb_det_poll (stream 0, seq = order).  Consumes the handle's
 * stream states: lqb_det_reset before going back to lqb_det_execute.  0 for seg_len: 262144. */
int     lqb_det_execute_sharded(lqb_det h, const float *iq, uint64_t n_samples, int mem, uint32_t seg_len, uint32_t preroll);
int     lqb_det_last_shard_info(lqb_det h, uint64_t out[4]);   /* segments, segment runs in all, rounds, launches */
int     lqb_det_poll(lqb_det h, lqb_detection *out, uint32_t max_out, uint32_t *n_out);
int     lqb_det_last_timing(lqb_det h, float *ms);
int     lqb_det_last_work(lqb_det h, uint64_t *windows);   /* detector windows evaluated by the last execute */
/* search work of the last execute: [0] windows, [1] alignments (= detections), [2] windows that needed the exact
 * evaluation, [3] CFO bins those evaluations visited (49 per window without the tensor-core bin filter) */
int     lqb_det_last_search(lqb_det h, uint64_t out[4]);

/* ------------------------------------------------------------------ host-side tables (no GPU needed) */
/* exposed so tests can check the product's own filter/table design against the oracle */
int lqb_tab_interp_taps(float beta, float *h30);
int lqb_tab_pfb_banks(float beta, float *banks32x28);
int lqb_tab_detector_template(float beta, float *s156_complex);
int lqb_tab_nco_sintab(float *tab1024);
int lqb_tab_secded_columns(uint32_t data_bytes /* 2, 4, 8 */, uint8_t *col /* 8 * data_bytes: parity-byte contribution of every data bit */);
int lqb_tab_ilv_bit_perm(uint32_t n_bytes, uint32_t *perm /* 8 * n_bytes */);   /* deinterleaved bit i = interleaved bit perm[i] */
int lqb_tab_packet_len(uint32_t payload_len, uint32_t check, uint32_t fec0, uint32_t fec1,
                       uint32_t mod_scheme, uint32_t *enc_bytes, uint32_t *n_symbols);

#ifdef __cplusplus
}
#endif
#endif /* LQB200_H */
