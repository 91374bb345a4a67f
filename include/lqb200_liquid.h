/*
 * lqb200_liquid.h -- single-stream compatibility veneer with liquid-dsp's own signatures.
 *
 * Exactly the liquid-dsp names, argument orders and struct layouts that the reference's three
 * block implementations use, so those .cc files build against liblqb200.so instead of -lliquid
 * (reference: lib/CMakeLists.txt:33) with only `#include <liquid/liquid.h>` redirected here:
 *
 *   lib/flex_rx_impl.cc:49,71,213 + :182-201   flexframesync_create/_destroy/_execute, callback
 *   lib/flex_tx_impl.cc:51,56,72,188,198-201   flexframegenprops_init_default, flexframegen_*
 *   lib/frame_detector_cc_impl.cc:47-55,63,77  msequence_*, qdetector_cccf_*
 *
 * Every object here is a thin adapter over the batch C-ABI in lqb200.h with n_streams = 1 and
 * host buffers; the arithmetic runs on the GPU.  Differences from liquid-dsp that a caller can
 * observe (all about WHEN, never WHAT):
 *   - flexframesync_execute batches samples before they go to the GPU, so a callback comes some
 *     calls after the sample that completed its frame.  Delivery is paced at one callback per 256
 *     samples of the current call: the reference's loop (256 samples per call, single-slot
 *     packet_info, lib/flex_rx_impl.cc:212-251) never sees two callbacks in one call, a caller
 *     passing a large buffer gets every completed frame before execute returns.  Samples still
 *     waiting for a full batch at the end of a capture are processed by flexframesync_flush()
 *     (an extension; liquid-dsp works sample by sample and has nothing to flush).
 *   - qdetector_cccf_execute reports a detection on the call that completes a 256-sample hop
 *     rather than on the exact sample; the count and the estimates are the same.
 *   - errors never exit(): create functions return NULL (see lqb_last_error()).
 */
#ifndef LQB200_LIQUID_H
#define LQB200_LIQUID_H

#ifdef __cplusplus
#include <complex>
typedef std::complex<float> liquid_float_complex;
extern "C" {
#else
#include <complex.h>
typedef float complex liquid_float_complex;
#endif

/* enums (wire values) */
typedef enum {
    LIQUID_MODEM_UNKNOWN = 0,
    LIQUID_MODEM_PSK2, LIQUID_MODEM_PSK4, LIQUID_MODEM_PSK8, LIQUID_MODEM_PSK16,
    LIQUID_MODEM_PSK32, LIQUID_MODEM_PSK64, LIQUID_MODEM_PSK128, LIQUID_MODEM_PSK256,
    LIQUID_MODEM_DPSK2, LIQUID_MODEM_DPSK4, LIQUID_MODEM_DPSK8, LIQUID_MODEM_DPSK16,
    LIQUID_MODEM_DPSK32, LIQUID_MODEM_DPSK64, LIQUID_MODEM_DPSK128, LIQUID_MODEM_DPSK256,
    LIQUID_MODEM_ASK2, LIQUID_MODEM_ASK4, LIQUID_MODEM_ASK8, LIQUID_MODEM_ASK16,
    LIQUID_MODEM_ASK32, LIQUID_MODEM_ASK64, LIQUID_MODEM_ASK128, LIQUID_MODEM_ASK256,
    LIQUID_MODEM_QAM4, LIQUID_MODEM_QAM8, LIQUID_MODEM_QAM16, LIQUID_MODEM_QAM32,
    LIQUID_MODEM_QAM64, LIQUID_MODEM_QAM128, LIQUID_MODEM_QAM256,
    LIQUID_MODEM_APSK4, LIQUID_MODEM_APSK8, LIQUID_MODEM_APSK16, LIQUID_MODEM_APSK32,
    LIQUID_MODEM_APSK64, LIQUID_MODEM_APSK128, LIQUID_MODEM_APSK256,
    LIQUID_MODEM_BPSK, LIQUID_MODEM_QPSK, LIQUID_MODEM_OOK
} modulation_scheme;

typedef enum {
    LIQUID_FEC_UNKNOWN = 0, LIQUID_FEC_NONE, LIQUID_FEC_REP3, LIQUID_FEC_REP5, LIQUID_FEC_HAMMING74,
    LIQUID_FEC_HAMMING84, LIQUID_FEC_HAMMING128, LIQUID_FEC_GOLAY2412, LIQUID_FEC_SECDED2216,
    LIQUID_FEC_SECDED3932, LIQUID_FEC_SECDED7264, LIQUID_FEC_CONV_V27, LIQUID_FEC_CONV_V29,
    LIQUID_FEC_CONV_V39, LIQUID_FEC_CONV_V615, LIQUID_FEC_CONV_V27P23, LIQUID_FEC_CONV_V27P34,
    LIQUID_FEC_CONV_V27P45, LIQUID_FEC_CONV_V27P56, LIQUID_FEC_CONV_V27P67, LIQUID_FEC_CONV_V27P78,
    LIQUID_FEC_CONV_V29P23, LIQUID_FEC_CONV_V29P34, LIQUID_FEC_CONV_V29P45, LIQUID_FEC_CONV_V29P56,
    LIQUID_FEC_CONV_V29P67, LIQUID_FEC_CONV_V29P78, LIQUID_FEC_RS_M8
} fec_scheme;

typedef enum {
    LIQUID_CRC_UNKNOWN = 0, LIQUID_CRC_NONE, LIQUID_CRC_CHECKSUM, LIQUID_CRC_8, LIQUID_CRC_16,
    LIQUID_CRC_24, LIQUID_CRC_32
} crc_scheme;

#define LIQUID_FIRFILT_ARKAISER 9

/* ---- msequence ---- */
typedef struct msequence_s *msequence;
msequence    msequence_create(unsigned int m, unsigned int g, unsigned int a);
void         msequence_destroy(msequence ms);
unsigned int msequence_advance(msequence ms);

/* ---- frame synchroniser ---- */
typedef struct {
    float evm, rssi, cfo;
    liquid_float_complex *framesyms;
    unsigned int num_framesyms;
    unsigned int mod_scheme, mod_bps, check, fec0, fec1;
} framesyncstats_s;

typedef int (*framesync_callback)(unsigned char *header, int header_valid, unsigned char *payload,
                                  unsigned int payload_len, int payload_valid, framesyncstats_s stats, void *userdata);

typedef struct flexframesync_s *flexframesync;
flexframesync flexframesync_create(framesync_callback callback, void *userdata);
void flexframesync_destroy(flexframesync q);
void flexframesync_reset(flexframesync q);
void flexframesync_execute(flexframesync q, liquid_float_complex *x, unsigned int n);
void flexframesync_flush(flexframesync q);      /* extension: process pending samples, deliver all queued frames */

/* ---- frame generator ---- */
typedef struct { unsigned int check, fec0, fec1, mod_scheme; } flexframegenprops_s;
typedef struct flexframegen_s *flexframegen;
void flexframegenprops_init_default(flexframegenprops_s *props);
flexframegen flexframegen_create(flexframegenprops_s *props);
void flexframegen_destroy(flexframegen q);
int  flexframegen_setprops(flexframegen q, flexframegenprops_s *props);
void flexframegen_assemble(flexframegen q, const unsigned char *header, const unsigned char *payload, unsigned int payload_len);
unsigned int flexframegen_getframelen(flexframegen q);
int  flexframegen_write_samples(flexframegen q, liquid_float_complex *buffer, unsigned int buffer_len);

/* ---- detector ---- */
typedef struct qdetector_cccf_s *qdetector_cccf;
qdetector_cccf qdetector_cccf_create_linear(liquid_float_complex *sequence, unsigned int sequence_len,
                                            int ftype, unsigned int k, unsigned int m, float beta);
void  qdetector_cccf_destroy(qdetector_cccf q);
void  qdetector_cccf_reset(qdetector_cccf q);
void  qdetector_cccf_set_threshold(qdetector_cccf q, float threshold);
void *qdetector_cccf_execute(qdetector_cccf q, liquid_float_complex x);
float qdetector_cccf_get_tau(qdetector_cccf q);
float qdetector_cccf_get_gamma(qdetector_cccf q);
float qdetector_cccf_get_dphi(qdetector_cccf q);
float qdetector_cccf_get_phi(qdetector_cccf q);
unsigned int qdetector_cccf_get_buf_len(qdetector_cccf q);

#ifdef __cplusplus
}
#endif
#endif /* LQB200_LIQUID_H */
