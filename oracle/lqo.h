/*
 * lqo.h -- CPU ORACLE for the gr-liquiddsp packet-PHY hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is shipped or linked into
 * the product library (liblqb200.so); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference (gvanhoy/gr-liquiddsp) holds none of this
 * arithmetic; it calls liquid-dsp (un-vendored, un-pinned `-lliquid`,
 * /root/reference/lib/CMakeLists.txt:33) at
 *   lib/flex_rx_impl.cc:49,71,213          flexframesync_create/destroy/execute
 *   lib/flex_tx_impl.cc:51,56,188,198-201  flexframegen_* (props, assemble, write)
 *   lib/frame_detector_cc_impl.cc:47-55,77 msequence_*, qdetector_cccf_*
 * and its own tests hold no golden vectors (python/qa_*.py:34-37 are empty).
 * liquid-dsp is not installed here and there is no network, so this file is a
 * restatement of liquid-dsp >= 1.3.1's published algorithms (SURVEY.md
 * Appendix A) plus Phil Karn's libfec (viterbi27/29 "port" butterflies,
 * rs_char) -- written from the algorithm descriptions, not from source.
 * Wherever the recollection was uncertain the choice made here is stated in
 * docs/FRAME_FORMAT.md and is authoritative for this project.
 *
 * Float discipline: this library is compiled with -ffp-contract=off and uses
 * fmaf() explicitly where a fused multiply-add is intended, so that the CUDA
 * kernels can reproduce the same roundings with __fmaf_rn/__fmul_rn/__fadd_rn.
 */
#ifndef LQO_H
#define LQO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } lqo_cf;

/* ---- wire enums (values travel in the frame header) ---- */
enum {
    LQ_MODEM_UNKNOWN = 0,
    LQ_MODEM_PSK2 = 1, LQ_MODEM_PSK4, LQ_MODEM_PSK8, LQ_MODEM_PSK16,
    LQ_MODEM_PSK32, LQ_MODEM_PSK64, LQ_MODEM_PSK128, LQ_MODEM_PSK256,
    LQ_MODEM_DPSK2 = 9, LQ_MODEM_DPSK4, LQ_MODEM_DPSK8, LQ_MODEM_DPSK16,
    LQ_MODEM_DPSK32, LQ_MODEM_DPSK64, LQ_MODEM_DPSK128, LQ_MODEM_DPSK256,
    LQ_MODEM_ASK2 = 17, LQ_MODEM_ASK4, LQ_MODEM_ASK8, LQ_MODEM_ASK16,
    LQ_MODEM_ASK32, LQ_MODEM_ASK64, LQ_MODEM_ASK128, LQ_MODEM_ASK256,
    LQ_MODEM_QAM4 = 25, LQ_MODEM_QAM8, LQ_MODEM_QAM16, LQ_MODEM_QAM32,
    LQ_MODEM_QAM64, LQ_MODEM_QAM128, LQ_MODEM_QAM256,
    LQ_MODEM_APSK4 = 32, /* ... APSK256 = 38: not implemented */
    LQ_MODEM_BPSK = 39, LQ_MODEM_QPSK = 40, LQ_MODEM_OOK = 41,
    LQ_MODEM_NUM_SCHEMES = 52
};
enum {
    LQ_FEC_UNKNOWN = 0, LQ_FEC_NONE, LQ_FEC_REP3, LQ_FEC_REP5, LQ_FEC_HAMMING74,
    LQ_FEC_HAMMING84, LQ_FEC_HAMMING128, LQ_FEC_GOLAY2412, LQ_FEC_SECDED2216,
    LQ_FEC_SECDED3932, LQ_FEC_SECDED7264, LQ_FEC_CONV_V27, LQ_FEC_CONV_V29,
    LQ_FEC_CONV_V39, LQ_FEC_CONV_V615, LQ_FEC_CONV_V27P23, LQ_FEC_CONV_V27P34,
    LQ_FEC_CONV_V27P45, LQ_FEC_CONV_V27P56, LQ_FEC_CONV_V27P67, LQ_FEC_CONV_V27P78,
    LQ_FEC_CONV_V29P23, LQ_FEC_CONV_V29P34, LQ_FEC_CONV_V29P45, LQ_FEC_CONV_V29P56,
    LQ_FEC_CONV_V29P67, LQ_FEC_CONV_V29P78, LQ_FEC_RS_M8, LQ_FEC_NUM_SCHEMES
};
enum {
    LQ_CRC_UNKNOWN = 0, LQ_CRC_NONE, LQ_CRC_CHECKSUM, LQ_CRC_8, LQ_CRC_16,
    LQ_CRC_24, LQ_CRC_32, LQ_CRC_NUM_SCHEMES
};
enum { LQ_FIRFILT_ARKAISER = 9 };

/* ---- m-sequence (A.2) ---- */
typedef struct { unsigned m, g, a, n, v, b; } lqo_mseq;
void     lqo_mseq_init(lqo_mseq *q, unsigned m, unsigned g, unsigned a);
void     lqo_mseq_init_default(lqo_mseq *q, unsigned m);
unsigned lqo_mseq_advance(lqo_mseq *q);
unsigned lqo_mseq_symbol(lqo_mseq *q, unsigned bps);

/* ---- filter design / polyphase banks (A.4) ---- */
void lqo_firdes_arkaiser(unsigned k, unsigned m, float beta, float dt, float *h /* 2km+1 */);
/* 2-phase interpolator taps, padded to 30 (A.4, firinterp_crcf_create_prototype) */
void lqo_interp_taps(unsigned k, unsigned m, float beta, float *h30);
/* 32-bank matched filter: bank[i][n], n = 0..27 = taps applied oldest->newest */
void lqo_pfb_rnyquist(unsigned npfb, unsigned k, unsigned m, float beta, float *banks /* npfb*2km */);

/* ---- FFT (radix-2 DIT, unnormalised both ways; n power of two <= 512) ---- */
#define LQO_FFT_FORWARD  (+1)
#define LQO_FFT_BACKWARD (-1)
void lqo_fft(const lqo_cf *in, lqo_cf *out, unsigned n, int dir);

/* ---- NCO with 32-bit phase accumulator and 1024-entry sine table (A.5, liquid >= 1.3.1) ---- */
typedef struct { uint32_t theta, d_theta; float alpha, beta; } lqo_nco;
void     lqo_nco_reset(lqo_nco *q);
uint32_t lqo_nco_constrain(float theta);
void     lqo_nco_set_frequency(lqo_nco *q, float dtheta);
void     lqo_nco_set_phase(lqo_nco *q, float theta);
float    lqo_nco_get_frequency(const lqo_nco *q);
void     lqo_nco_pll_set_bandwidth(lqo_nco *q, float bw);
void     lqo_nco_pll_step(lqo_nco *q, float dphi);
void     lqo_nco_step(lqo_nco *q);
lqo_cf   lqo_nco_mix_down(const lqo_nco *q, lqo_cf x);
const float *lqo_nco_sintab(void); /* 1024 floats */

/* ---- CRC / scrambler / interleaver (A.7) ---- */
unsigned lqo_crc_len(int scheme);
unsigned lqo_crc_key(int scheme, const uint8_t *msg, unsigned n);
void     lqo_scramble(uint8_t *x, unsigned n);
void     lqo_interleave(uint8_t *x, unsigned n, int depth);   /* in place */
void     lqo_deinterleave(uint8_t *x, unsigned n, int depth); /* in place */

/* ---- FEC ---- */
unsigned lqo_fec_enc_len(int fs, unsigned dec_len);
void     lqo_fec_encode(int fs, unsigned dec_len, const uint8_t *dec, uint8_t *enc);
void     lqo_fec_decode(int fs, unsigned dec_len, const uint8_t *enc, uint8_t *dec);
/* soft-input Viterbi for the convolutional schemes: soft[] holds one byte per coded bit (8 * enc_len of them, in
 * transmission order), same trellis, metrics and traceback as the hard decoder; returns 0 for other schemes */
int      lqo_fec_decode_soft(int fs, unsigned dec_len, const uint8_t *soft, uint8_t *dec);
int      lqo_fec_is_conv(int fs);
/* the byte interleaver of an n-byte block as a permutation of its 8 n coded bits (MSB-first positions):
 * deinterleaved[i] = interleaved[perm[i]] */
void     lqo_deinterleave_bit_perm(unsigned n, uint32_t *perm /* 8 n */);
/* exposed for KATs */
void     lqo_secded_columns(unsigned nb /* 2, 4 or 8 data bytes */, uint8_t *col /* 8 nb: parity-byte contribution of every data bit */);
int      lqo_rs_decode_block(uint8_t *block /* 255-pad */, unsigned pad); /* returns #corrected or -1 */
void     lqo_rs_encode_block(const uint8_t *data, unsigned pad, uint8_t *parity32);

/* ---- packetizer ---- */
unsigned lqo_packetizer_enc_len(unsigned n, int check, int fec0, int fec1);
void     lqo_packetizer_encode(unsigned n, int check, int fec0, int fec1, const uint8_t *msg, uint8_t *pkt);
int      lqo_packetizer_decode(unsigned n, int check, int fec0, int fec1, const uint8_t *pkt, uint8_t *msg);

/* ---- modem (A.6) ---- */
typedef struct {
    int scheme; unsigned bps, M;
    unsigned m_i, m_q;          /* QAM split */
    float alpha, d_phi, ref[8];
    float dpsk_phi;             /* DPSK memory */
    lqo_cf map[256];            /* symbol map (modulate table) */
    lqo_cf x_hat, r;            /* demod state */
} lqo_modem;
int      lqo_modem_supported(int scheme);
unsigned lqo_modem_bps(int scheme);
int      lqo_modem_init(lqo_modem *q, int scheme);
void     lqo_modem_reset(lqo_modem *q);
lqo_cf   lqo_modem_modulate(lqo_modem *q, unsigned sym);
unsigned lqo_modem_demodulate(lqo_modem *q, lqo_cf x);
float    lqo_modem_phase_error(const lqo_modem *q);
/* Soft decisions (opt-in extension; OUR definition, liquid's own soft demodulator is not
 * restated): for bit k (MSB first) of the symbol, with d0 / d1 the smallest squared distance from x to a constellation
 * point whose bit k is 0 / 1 and G = 64 / (smallest squared distance between two points),
 *     soft[k] = clamp(trunc(128 + G (d0 - d1)), 0, 255)        -- 0 = surely 0, 255 = surely 1 (libfec's convention).
 * Not defined for DPSK (returns 0: the caller falls back to hard decisions). */
int      lqo_modem_demodulate_soft(const lqo_modem *q, lqo_cf x, uint8_t *soft /* bps */);
float    lqo_modem_evm(const lqo_modem *q);
/* the pinned arg() / exp(j t) of the per-symbol loops (see lqo_modem.c) */
float    lqo_pm_atan2f(float y, float x);
void     lqo_pm_sincosf(float t, float *sn, float *cs);

/* ---- qpacketmodem ---- */
unsigned lqo_qpm_frame_len(unsigned payload_len, int check, int fec0, int fec1, int ms);
void     lqo_qpm_encode(unsigned payload_len, int check, int fec0, int fec1, int ms,
                        const uint8_t *payload, lqo_cf *frame);
/* soft-decision variant: the coding stage nearest the channel (fec1, or fec0 when fec1 is "none") is decoded from soft
 * bits when it is convolutional and the modem has a soft demodulator; everything else as lqo_qpm_decode */
int      lqo_qpm_decode_soft(unsigned payload_len, int check, int fec0, int fec1, int ms,
                             const lqo_cf *frame, uint8_t *payload);
int      lqo_qpm_decode(unsigned payload_len, int check, int fec0, int fec1, int ms,
                        const lqo_cf *frame, uint8_t *payload);

/* ---- qpilotgen / qpilotsync (A.8/A.9) ---- */
unsigned lqo_qpilot_num_pilots(unsigned payload_len, unsigned spacing);
unsigned lqo_qpilot_frame_len(unsigned payload_len, unsigned spacing);
void     lqo_qpilotgen(unsigned payload_len, unsigned spacing, const lqo_cf *payload, lqo_cf *frame);
void     lqo_qpilotsync(unsigned payload_len, unsigned spacing, const lqo_cf *frame, lqo_cf *payload,
                        float *dphi_hat, float *phi_hat, float *g_hat);

/* ---- qdetector_cccf (A.3) ---- */
typedef struct lqo_qdetector_s *lqo_qdetector;
lqo_qdetector lqo_qdetector_create_linear(const lqo_cf *seq, unsigned seq_len, int ftype,
                                          unsigned k, unsigned m, float beta);
void     lqo_qdetector_destroy(lqo_qdetector q);
void     lqo_qdetector_reset(lqo_qdetector q);
void     lqo_qdetector_set_threshold(lqo_qdetector q, float thr);
void     lqo_qdetector_set_range(lqo_qdetector q, float dphi_max);
/* returns pointer to nfft aligned samples on detection else NULL */
const lqo_cf *lqo_qdetector_execute(lqo_qdetector q, lqo_cf x);
float    lqo_qdetector_get_tau(lqo_qdetector q);
float    lqo_qdetector_get_gamma(lqo_qdetector q);
float    lqo_qdetector_get_dphi(lqo_qdetector q);
float    lqo_qdetector_get_phi(lqo_qdetector q);
float    lqo_qdetector_get_rxy(lqo_qdetector q);
unsigned lqo_qdetector_get_buf_len(lqo_qdetector q);
unsigned lqo_qdetector_get_seq_len(lqo_qdetector q);
const lqo_cf *lqo_qdetector_get_template(lqo_qdetector q);

/* ---- flexframegen (A.8) ---- */
typedef struct { int check, fec0, fec1, mod_scheme; } lqo_fgprops;
typedef struct lqo_flexframegen_s *lqo_flexframegen;
void     lqo_fgprops_init_default(lqo_fgprops *p);
lqo_flexframegen lqo_flexframegen_create(const lqo_fgprops *p);
void     lqo_flexframegen_destroy(lqo_flexframegen q);
void     lqo_flexframegen_setprops(lqo_flexframegen q, const lqo_fgprops *p);
void     lqo_flexframegen_assemble(lqo_flexframegen q, const uint8_t *header14,
                                   const uint8_t *payload, unsigned payload_len);
unsigned lqo_flexframegen_getframelen(lqo_flexframegen q);
int      lqo_flexframegen_write_samples(lqo_flexframegen q, lqo_cf *buf, unsigned n);

/* ---- flexframesync (A.9) ---- */
typedef struct {
    float evm, rssi, cfo;
    const lqo_cf *framesyms; unsigned num_framesyms;
    unsigned mod_scheme, mod_bps, check, fec0, fec1;
    /* extensions beyond liquid's framesyncstats_s */
    float tau_hat, gamma_hat, dphi_hat, phi_hat, rxy;
    uint64_t sample_index;      /* absolute index of x[F], the first frame sample */
} lqo_framesyncstats;
typedef int (*lqo_framesync_callback)(const uint8_t *header, int header_valid,
                                      const uint8_t *payload, unsigned payload_len,
                                      int payload_valid, lqo_framesyncstats stats, void *userdata);
typedef struct lqo_flexframesync_s *lqo_flexframesync;
lqo_flexframesync lqo_flexframesync_create(lqo_framesync_callback cb, void *userdata);
void     lqo_flexframesync_destroy(lqo_flexframesync q);
void     lqo_flexframesync_reset(lqo_flexframesync q);
void     lqo_flexframesync_execute(lqo_flexframesync q, const lqo_cf *x, unsigned n);
void     lqo_flexframesync_set_soft(lqo_flexframesync q, int soft);   /* payloads through lqo_qpm_decode_soft */

/* ---- convenience collectors for ctypes-driven tests / bench ---- */
typedef struct {
    uint64_t sample_index;
    int header_valid, payload_valid;
    unsigned payload_len, num_framesyms;
    unsigned mod_scheme, mod_bps, check, fec0, fec1;
    float evm, rssi, cfo, tau_hat, gamma_hat, dphi_hat, phi_hat, rxy;
    uint8_t header[20];
    uint64_t payload_off;   /* offset into payload byte pool */
    uint64_t syms_off;      /* offset (in complex samples) into symbol pool */
} lqo_frame_record;
/* run a whole capture through one flexframesync; returns number of frames found
 * (records beyond max_frames / pool capacity are counted but not stored) */
unsigned lqo_rx_capture(const lqo_cf *x, uint64_t n, unsigned chunk,
                        lqo_frame_record *recs, unsigned max_frames,
                        uint8_t *payload_pool, uint64_t payload_cap,
                        lqo_cf *sym_pool, uint64_t sym_cap);
unsigned lqo_rx_capture_soft(const lqo_cf *x, uint64_t n, unsigned chunk,
                             lqo_frame_record *recs, unsigned max_frames,
                             uint8_t *payload_pool, uint64_t payload_cap,
                             lqo_cf *sym_pool, uint64_t sym_cap);
/* multi-threaded: n_streams captures of n samples each (stride in samples); returns total frames
 * and number of valid payloads through *n_valid; nothing stored (used for CPU timing) */
uint64_t lqo_rx_many(const lqo_cf *x, unsigned n_streams, uint64_t stride, uint64_t n,
                     unsigned n_threads, uint64_t *n_valid);
/* frame_detector_cc equivalent: returns number of detections, fills sample indices / estimates */
typedef struct { uint64_t sample_index; float tau_hat, gamma_hat, dphi_hat, phi_hat, rxy; } lqo_detection;
unsigned lqo_detect_capture(const lqo_cf *x, uint64_t n, float beta, float threshold,
                            lqo_detection *out, unsigned max_out);
/* assemble one frame to samples (cfg-1 style TX) */
unsigned lqo_tx_frame(int ms, int check, int fec0, int fec1, const uint8_t *header14,
                      const uint8_t *payload, unsigned payload_len, lqo_cf *out, unsigned out_cap);

#ifdef __cplusplus
}
#endif
#endif /* LQO_H */
