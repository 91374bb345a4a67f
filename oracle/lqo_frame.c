/*
 * lqo_frame.c -- ORACLE (test infrastructure only; see lqo.h header).
 * qpilotgen/qpilotsync, qdetector_cccf, flexframegen, flexframesync and the
 * capture-level collectors used by tests and the CPU baseline.
 * SURVEY.md Appendix A.3, A.8, A.9.  Reference call sites replaced:
 *   /root/reference/lib/flex_rx_impl.cc:49,213        (flexframesync_create/execute)
 *   /root/reference/lib/flex_tx_impl.cc:56,188,198-201 (flexframegen_*)
 *   /root/reference/lib/frame_detector_cc_impl.cc:54-55,77 (qdetector_cccf_*)
 */
#include "lqo.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FF_K          2
#define FF_M          7
#define FF_TX_BETA    0.25f     /* flexframegen interpolator (A.4, [LQ-L]) */
#define FF_RX_BETA    0.30f     /* flexframesync detector template + matched filter bank */
#define FF_NPFB       32
#define FF_PREAMBLE   64
#define FF_H_USER     14
#define FF_H_DEC      (FF_H_USER + 6)
#define FF_PROTOCOL   102       /* 101 + PACKETIZER_VERSION(1) */
#define FF_H_CRC      LQ_CRC_32
#define FF_H_FEC0     LQ_FEC_SECDED7264
#define FF_H_FEC1     LQ_FEC_HAMMING84
#define FF_H_MOD      LQ_MODEM_QPSK
#define FF_PILOT_SPACING 16
#define FF_PLL_BW     1e-4f

static lqo_cf cmul_(lqo_cf a, lqo_cf b)
{
    lqo_cf y;
    y.re = fmaf(-a.im, b.im, a.re * b.re);
    y.im = fmaf(a.im, b.re, a.re * b.im);
    return y;
}
static lqo_cf conj_(lqo_cf a) { a.im = -a.im; return a; }
static float abs2_(lqo_cf a) { return fmaf(a.im, a.im, a.re * a.re); }
static float cabs_(lqo_cf a) { return sqrtf(abs2_(a)); }

/* energy of n (power of two) samples as a balanced pairwise tree in index order */
static float energy_tree_(const lqo_cf *x, unsigned n)
{
    float e[256];
    for (unsigned i = 0; i < n; i++) e[i] = abs2_(x[i]);
    for (unsigned w = n; w > 1; w >>= 1)
        for (unsigned i = 0; i < w / 2; i++) e[i] = e[2 * i] + e[2 * i + 1];
    return e[0];
}

static void preamble_pn_(lqo_cf *pn)
{
    lqo_mseq ms;
    lqo_mseq_init(&ms, 7, 0x0089, 1);
    for (unsigned i = 0; i < FF_PREAMBLE; i++) {
        pn[i].re = lqo_mseq_advance(&ms) ? (float)M_SQRT1_2 : -(float)M_SQRT1_2;
        pn[i].im = lqo_mseq_advance(&ms) ? (float)M_SQRT1_2 : -(float)M_SQRT1_2;
    }
}

/* ================================================================== qpilot */
unsigned lqo_qpilot_num_pilots(unsigned payload_len, unsigned spacing)
{
    return payload_len / (spacing - 1) + (payload_len % (spacing - 1) ? 1 : 0);
}
unsigned lqo_qpilot_frame_len(unsigned payload_len, unsigned spacing)
{
    return payload_len + lqo_qpilot_num_pilots(payload_len, spacing);
}
static unsigned nextpow2_(unsigned v) { unsigned n = 0; while ((1u << n) < v) n++; return n; }

static void pilots_(unsigned np, lqo_cf *p)
{
    lqo_mseq ms;
    lqo_mseq_init_default(&ms, nextpow2_(np));
    for (unsigned i = 0; i < np; i++) {
        unsigned s = lqo_mseq_symbol(&ms, 2);
        float theta = (2.0f * (float)M_PI * (float)s / 4.0f) + (float)M_PI / 4.0f;
        p[i].re = cosf(theta); p[i].im = sinf(theta);
    }
}

void lqo_qpilotgen(unsigned payload_len, unsigned spacing, const lqo_cf *payload, lqo_cf *frame)
{
    unsigned np = lqo_qpilot_num_pilots(payload_len, spacing), fl = payload_len + np, n = 0, p = 0;
    lqo_cf *pil = (lqo_cf *)malloc(np * sizeof(lqo_cf));
    pilots_(np, pil);
    for (unsigned i = 0; i < fl; i++) frame[i] = (i % spacing == 0) ? pil[p++] : payload[n++];
    free(pil);
}

void lqo_qpilotsync(unsigned payload_len, unsigned spacing, const lqo_cf *frame, lqo_cf *payload,
                    float *dphi_out, float *phi_out, float *g_out)
{
    unsigned np = lqo_qpilot_num_pilots(payload_len, spacing), fl = payload_len + np;
    unsigned nfft = 1u << nextpow2_(np + (np >> 1));
    lqo_cf *pil = (lqo_cf *)malloc(np * sizeof(lqo_cf));
    lqo_cf bt[512], bf[512];
    pilots_(np, pil);
    memset(bt, 0, sizeof bt);
    for (unsigned i = 0; i < np; i++) bt[i] = cmul_(frame[i * spacing], conj_(pil[i]));
    lqo_fft(bt, bf, nfft, LQO_FFT_FORWARD);
    unsigned i0 = 0; float y0 = 0.0f;
    for (unsigned i = 0; i < nfft; i++) { float a = cabs_(bf[i]); if (i == 0 || a > y0) { i0 = i; y0 = a; } }
    float ypos = cabs_(bf[(i0 + 1) % nfft]), yneg = cabs_(bf[(i0 + nfft - 1) % nfft]);
    float a = 0.5f * (ypos + yneg) - y0, b = 0.5f * (ypos - yneg);
    float idx = -b / (2.0f * a);
    float index = (float)i0 + idx;
    float dphi = (i0 > nfft / 2 ? index - (float)nfft : index) * 2.0f * (float)M_PI / (float)(nfft * spacing);
    lqo_cf metric = { 0.0f, 0.0f };
    for (unsigned i = 0; i < np; i++) {
        float ang = -dphi * (float)i * (float)spacing;
        lqo_cf rot; lqo_pm_sincosf(ang, &rot.im, &rot.re);      /* pinned exp(j ang), see lqo_modem.c */
        lqo_cf v = cmul_(bt[i], rot);
        metric.re += v.re; metric.im += v.im;
    }
    float phi = lqo_pm_atan2f(metric.im, metric.re);
    float g_hat = cabs_(metric) / (float)np;
    float g = 1.0f / g_hat;
    unsigned n = 0;
    for (unsigned i = 0; i < fl; i++) {
        if (i % spacing == 0) continue;
        float ang = -(dphi * (float)i + phi);
        lqo_cf rot; lqo_pm_sincosf(ang, &rot.im, &rot.re);      /* pinned exp(j ang), see lqo_modem.c */
        lqo_cf v = cmul_(frame[i], rot);
        payload[n].re = v.re * g; payload[n].im = v.im * g; n++;
    }
    *dphi_out = dphi; *phi_out = phi; *g_out = g_hat;
    free(pil);
}

/* ================================================================== qdetector */
struct lqo_qdetector_s {
    unsigned s_len, nfft, counter;
    lqo_cf *s, *S;
    float s2_sum, x2_sum_0, x2_sum_1, threshold;
    int range, offset, state, frame_detected, first_half_zero;
    lqo_cf buf_time_0[512], buf_time_1[512], buf_freq_0[512], buf_freq_1[512];
    float rxy, tau_hat, gamma_hat, dphi_hat, phi_hat;
};
enum { QD_SEEK = 0, QD_ALIGN = 1 };

lqo_qdetector lqo_qdetector_create_linear(const lqo_cf *seq, unsigned seq_len, int ftype,
                                          unsigned k, unsigned m, float beta)
{
    (void)ftype;
    lqo_qdetector q = (lqo_qdetector)calloc(1, sizeof *q);
    q->s_len = k * (seq_len + 2 * m);
    q->s = (lqo_cf *)calloc(q->s_len, sizeof(lqo_cf));
    float h[64];
    lqo_interp_taps(k, m, beta, h);
    unsigned sub = (2 * k * m + 1 + k - 1) / k;
    for (unsigned t = 0; t < seq_len + 2 * m; t++)           /* s[k t + i] = sum_n h[i + k n] sym[t - n] */
        for (unsigned i = 0; i < k; i++) {
            float ar = 0.0f, ai = 0.0f;
            for (unsigned n = sub; n-- > 0;) {               /* oldest symbol first */
                if (n > t || t - n >= seq_len) continue;
                ar = fmaf(h[i + k * n], seq[t - n].re, ar);
                ai = fmaf(h[i + k * n], seq[t - n].im, ai);
            }
            q->s[k * t + i].re = ar; q->s[k * t + i].im = ai;
        }
    q->s2_sum = 0.0f;
    for (unsigned i = 0; i < q->s_len; i++) q->s2_sum += abs2_(q->s[i]);
    q->nfft = 1u << nextpow2_(2 * q->s_len);
    q->S = (lqo_cf *)calloc(q->nfft, sizeof(lqo_cf));
    memcpy(q->buf_time_0, q->s, q->s_len * sizeof(lqo_cf));
    lqo_fft(q->buf_time_0, q->S, q->nfft, LQO_FFT_FORWARD);
    lqo_qdetector_set_threshold(q, 0.5f);
    lqo_qdetector_set_range(q, 0.3f);
    lqo_qdetector_reset(q);
    return q;
}

void lqo_qdetector_destroy(lqo_qdetector q) { if (q) { free(q->s); free(q->S); free(q); } }

void lqo_qdetector_reset(lqo_qdetector q)
{
    q->counter = q->nfft / 2;
    q->x2_sum_0 = q->x2_sum_1 = 0.0f;
    q->state = QD_SEEK;
    q->frame_detected = 0;
    q->first_half_zero = 1;
    memset(q->buf_time_0, 0, sizeof q->buf_time_0);
}
void lqo_qdetector_set_threshold(lqo_qdetector q, float t) { q->threshold = t; }
void lqo_qdetector_set_range(lqo_qdetector q, float d) { q->range = (int)(d * (float)q->nfft / (2.0f * (float)M_PI)); }
float lqo_qdetector_get_tau(lqo_qdetector q) { return q->tau_hat; }
float lqo_qdetector_get_gamma(lqo_qdetector q) { return q->gamma_hat; }
float lqo_qdetector_get_dphi(lqo_qdetector q) { return q->dphi_hat; }
float lqo_qdetector_get_phi(lqo_qdetector q) { return q->phi_hat; }
float lqo_qdetector_get_rxy(lqo_qdetector q) { return q->rxy; }
unsigned lqo_qdetector_get_buf_len(lqo_qdetector q) { return q->nfft; }
unsigned lqo_qdetector_get_seq_len(lqo_qdetector q) { return q->s_len; }
const lqo_cf *lqo_qdetector_get_template(lqo_qdetector q) { return q->s; }

static void qd_cross_(lqo_qdetector q, int offset)
{
    unsigned n = q->nfft;
    for (unsigned i = 0; i < n; i++) {
        unsigned j = (i + n - (unsigned)offset) % n;     /* offset may be negative: wraps mod 2^32 then mod n (n | 2^32) */
        q->buf_freq_1[i] = cmul_(q->buf_freq_0[i], conj_(q->S[j]));
    }
    lqo_fft(q->buf_freq_1, q->buf_time_1, n, LQO_FFT_BACKWARD);
}

static void qd_seek_(lqo_qdetector q, lqo_cf x)
{
    unsigned n = q->nfft, h = n / 2;
    q->buf_time_0[q->counter++] = x;
    if (q->counter < n) return;
    q->counter = h;
    q->x2_sum_1 = energy_tree_(q->buf_time_0 + h, h);
    lqo_fft(q->buf_time_0, q->buf_freq_0, n, LQO_FFT_FORWARD);
    float g0 = sqrtf(q->x2_sum_0 + q->x2_sum_1) * sqrtf((float)q->s_len / (float)n);
    if (g0 >= 1e-10f) {
        float g = 1.0f / ((float)n * g0 * sqrtf(q->s2_sum));
        float peak2 = 0.0f; unsigned pidx = 0; int poff = 0;
        for (int off = -q->range; off <= q->range; off++) {
            qd_cross_(q, off);
            for (unsigned i = 0; i < n; i++) {
                float a2 = abs2_(q->buf_time_1[i]);
                if (a2 > peak2) { peak2 = a2; pidx = i; poff = off; }
            }
        }
        float rxy_peak = sqrtf(peak2) * g;
        if (rxy_peak > q->threshold && pidx < n - q->s_len) {
            q->state = QD_ALIGN;
            q->offset = poff;
            q->rxy = rxy_peak;
            memmove(q->buf_time_0, q->buf_time_0 + pidx, (n - pidx) * sizeof(lqo_cf));
            q->counter = n - pidx;
            return;
        }
    }
    memmove(q->buf_time_0, q->buf_time_0 + h, h * sizeof(lqo_cf));
    q->x2_sum_0 = q->x2_sum_1;
    q->x2_sum_1 = 0.0f;
}

static void qd_align_finish_(lqo_qdetector q)
{
    unsigned n = q->nfft;
    lqo_fft(q->buf_time_0, q->buf_freq_0, n, LQO_FFT_FORWARD);
    qd_cross_(q, q->offset);
    float yneg = sqrtf(cabs_(q->buf_time_1[n - 1]));
    float y0   = sqrtf(cabs_(q->buf_time_1[0]));
    float ypos = sqrtf(cabs_(q->buf_time_1[1]));
    float a = 0.5f * (ypos + yneg) - y0, b = 0.5f * (ypos - yneg), c = y0;
    q->tau_hat = -b / (2.0f * a);
    float g_hat = a * q->tau_hat * q->tau_hat + b * q->tau_hat + c;
    q->gamma_hat = g_hat * g_hat / ((float)n * q->s2_sum);
    memcpy(q->buf_time_1, q->buf_time_0, n * sizeof(lqo_cf));
    for (unsigned i = 0; i < n; i++) {
        if (i < q->s_len) q->buf_time_0[i] = cmul_(q->buf_time_0[i], conj_(q->s[i]));
        else { q->buf_time_0[i].re = 0.0f; q->buf_time_0[i].im = 0.0f; }
    }
    lqo_fft(q->buf_time_0, q->buf_freq_0, n, LQO_FFT_FORWARD);
    float v0 = 0.0f; unsigned i0 = 0;
    for (unsigned i = 0; i < n; i++) { float v2 = abs2_(q->buf_freq_0[i]); if (v2 > v0) { v0 = v2; i0 = i; } }
    v0 = sqrtf(v0);
    float vneg = cabs_(q->buf_freq_0[(i0 + n - 1) % n]), vpos = cabs_(q->buf_freq_0[(i0 + 1) % n]);
    a = 0.5f * (vpos + vneg) - v0; b = 0.5f * (vpos - vneg);
    float idx = -b / (2.0f * a);
    float index = (float)i0 + idx;
    q->dphi_hat = (i0 > n / 2 ? index - (float)n : index) * 2.0f * (float)M_PI / (float)n;
    lqo_cf metric = { 0.0f, 0.0f };
    for (unsigned i = 0; i < q->s_len; i++) {
        float ang = -q->dphi_hat * (float)i;
        lqo_cf rot; lqo_pm_sincosf(ang, &rot.im, &rot.re);      /* pinned exp(j ang), see lqo_modem.c */
        lqo_cf v = cmul_(q->buf_time_0[i], rot);
        metric.re += v.re; metric.im += v.im;
    }
    q->phi_hat = lqo_pm_atan2f(metric.im, metric.re);
    q->frame_detected = 1;
    memmove(q->buf_time_0, q->buf_time_1 + n / 2, (n / 2) * sizeof(lqo_cf));
    q->state = QD_SEEK;
    q->x2_sum_0 = energy_tree_(q->buf_time_0, n / 2);
    q->x2_sum_1 = 0.0f;
    q->counter = n / 2;
}

const lqo_cf *lqo_qdetector_execute(lqo_qdetector q, lqo_cf x)
{
    if (q->state == QD_SEEK) {
        qd_seek_(q, x);
        /* peak at lag 0: the window already is the aligned buffer (liquid would overrun here) */
        if (q->state == QD_ALIGN && q->counter >= q->nfft) qd_align_finish_(q);
    } else {
        q->buf_time_0[q->counter++] = x;
        if (q->counter >= q->nfft) qd_align_finish_(q);
    }
    if (q->frame_detected) { q->frame_detected = 0; return q->buf_time_1; }
    return NULL;
}

/* ================================================================== flexframegen */
struct lqo_flexframegen_s {
    lqo_fgprops props;
    lqo_cf preamble[FF_PREAMBLE];
    float h[32];
    uint8_t header[FF_H_DEC];
    unsigned header_mod_len, header_sym_len, payload_sym_len, payload_dec_len;
    lqo_cf *header_sym, *payload_sym;
    /* write state */
    unsigned symbol_counter, sample_counter, frame_syms;
    lqo_cf hist[16];
    lqo_cf buf_interp[FF_K];
    int assembled;
};

void lqo_fgprops_init_default(lqo_fgprops *p)
{
    p->check = LQ_CRC_16; p->fec0 = LQ_FEC_NONE; p->fec1 = LQ_FEC_NONE; p->mod_scheme = LQ_MODEM_QPSK;
}

lqo_flexframegen lqo_flexframegen_create(const lqo_fgprops *p)
{
    lqo_flexframegen q = (lqo_flexframegen)calloc(1, sizeof *q);
    if (p) q->props = *p; else lqo_fgprops_init_default(&q->props);
    preamble_pn_(q->preamble);
    lqo_interp_taps(FF_K, FF_M, FF_TX_BETA, q->h);
    q->header_mod_len = lqo_qpm_frame_len(FF_H_DEC, FF_H_CRC, FF_H_FEC0, FF_H_FEC1, FF_H_MOD);
    q->header_sym_len = lqo_qpilot_frame_len(q->header_mod_len, FF_PILOT_SPACING);
    q->header_sym = (lqo_cf *)calloc(q->header_sym_len, sizeof(lqo_cf));
    return q;
}
void lqo_flexframegen_destroy(lqo_flexframegen q) { if (q) { free(q->header_sym); free(q->payload_sym); free(q); } }
void lqo_flexframegen_setprops(lqo_flexframegen q, const lqo_fgprops *p) { q->props = *p; }

void lqo_flexframegen_assemble(lqo_flexframegen q, const uint8_t *hdr, const uint8_t *payload, unsigned n)
{
    unsigned u = FF_H_USER;
    if (hdr) memcpy(q->header, hdr, u); else memset(q->header, 0, u);
    q->payload_dec_len = n;
    q->header[u + 0] = FF_PROTOCOL;
    q->header[u + 1] = (uint8_t)((n >> 8) & 0xff);
    q->header[u + 2] = (uint8_t)(n & 0xff);
    q->header[u + 3] = (uint8_t)q->props.mod_scheme;
    q->header[u + 4] = (uint8_t)(((q->props.check & 0x07) << 5) | (q->props.fec0 & 0x1f));
    q->header[u + 5] = (uint8_t)(q->props.fec1 & 0x1f);
    lqo_cf *hm = (lqo_cf *)calloc(q->header_mod_len, sizeof(lqo_cf));
    lqo_qpm_encode(FF_H_DEC, FF_H_CRC, FF_H_FEC0, FF_H_FEC1, FF_H_MOD, q->header, hm);
    lqo_qpilotgen(q->header_mod_len, FF_PILOT_SPACING, hm, q->header_sym);
    free(hm);
    q->payload_sym_len = lqo_qpm_frame_len(n, q->props.check, q->props.fec0, q->props.fec1, q->props.mod_scheme);
    free(q->payload_sym);
    q->payload_sym = (lqo_cf *)calloc(q->payload_sym_len + 1, sizeof(lqo_cf));
    lqo_qpm_encode(n, q->props.check, q->props.fec0, q->props.fec1, q->props.mod_scheme, payload, q->payload_sym);
    q->frame_syms = FF_PREAMBLE + q->header_sym_len + q->payload_sym_len + 2 * FF_M;
    q->symbol_counter = 0; q->sample_counter = 0;
    memset(q->hist, 0, sizeof q->hist);
    q->assembled = 1;
}

unsigned lqo_flexframegen_getframelen(lqo_flexframegen q) { return q->assembled ? FF_K * q->frame_syms : 0; }

static lqo_cf fg_symbol_(lqo_flexframegen q, unsigned t)
{
    lqo_cf z = { 0.0f, 0.0f };
    if (t < FF_PREAMBLE) return q->preamble[t];
    t -= FF_PREAMBLE;
    if (t < q->header_sym_len) return q->header_sym[t];
    t -= q->header_sym_len;
    if (t < q->payload_sym_len) return q->payload_sym[t];
    return z;
}

int lqo_flexframegen_write_samples(lqo_flexframegen q, lqo_cf *buf, unsigned n)
{
    for (unsigned i = 0; i < n; i++) {
        if (q->sample_counter == 0) {
            lqo_cf sym = { 0.0f, 0.0f };
            if (q->symbol_counter < q->frame_syms) sym = fg_symbol_(q, q->symbol_counter);
            q->symbol_counter++;
            memmove(q->hist, q->hist + 1, 14 * sizeof(lqo_cf));  /* hist[14] = newest, hist[0] = oldest */
            q->hist[14] = sym;
            for (unsigned ph = 0; ph < FF_K; ph++) {
                float ar = 0.0f, ai = 0.0f;
                for (unsigned j = 0; j < 15; j++) {             /* oldest first: tap index 14-j */
                    float c = q->h[ph + FF_K * (14 - j)];
                    ar = fmaf(c, q->hist[j].re, ar);
                    ai = fmaf(c, q->hist[j].im, ai);
                }
                q->buf_interp[ph].re = ar; q->buf_interp[ph].im = ai;
            }
        }
        buf[i] = q->buf_interp[q->sample_counter];
        q->sample_counter = (q->sample_counter + 1) % FF_K;
    }
    return q->symbol_counter >= q->frame_syms && q->sample_counter == 0;
}

/* ================================================================== flexframesync */
enum { FS_DETECT = 0, FS_RXPREAMBLE, FS_RXHEADER, FS_RXPAYLOAD };
struct lqo_flexframesync_s {
    lqo_framesync_callback cb; void *ud;
    lqo_qdetector det;
    float banks[FF_NPFB * 2 * FF_K * FF_M];
    lqo_cf win[2 * FF_K * FF_M];          /* oldest .. newest */
    float mf_scale;
    lqo_nco mixer, pll;
    float tau_hat, gamma_hat, dphi_hat, phi_hat, rxy;
    unsigned pfb_index; int mf_counter;
    int state;
    unsigned preamble_counter, symbol_counter;
    unsigned header_mod_len, header_sym_len;
    lqo_cf *header_sym, *header_mod;
    uint8_t header_dec[FF_H_DEC];
    int header_valid;
    lqo_modem payload_demod;
    unsigned payload_sym_len, payload_dec_len, payload_cap;
    int ms, check, fec0, fec1;
    lqo_cf *payload_sym; uint8_t *payload_dec;
    float evm_acc;
    uint64_t n_consumed, frame_start;
    int soft;                             /* opt-in: payloads through lqo_qpm_decode_soft */
};

void lqo_flexframesync_set_soft(lqo_flexframesync q, int soft) { q->soft = soft; }

lqo_flexframesync lqo_flexframesync_create(lqo_framesync_callback cb, void *ud)
{
    lqo_flexframesync q = (lqo_flexframesync)calloc(1, sizeof *q);
    q->cb = cb; q->ud = ud;
    lqo_cf pn[FF_PREAMBLE];
    preamble_pn_(pn);
    q->det = lqo_qdetector_create_linear(pn, FF_PREAMBLE, LQ_FIRFILT_ARKAISER, FF_K, FF_M, FF_RX_BETA);
    lqo_qdetector_set_threshold(q->det, 0.5f);
    lqo_pfb_rnyquist(FF_NPFB, FF_K, FF_M, FF_RX_BETA, q->banks);
    lqo_nco_pll_set_bandwidth(&q->pll, FF_PLL_BW);
    q->header_mod_len = lqo_qpm_frame_len(FF_H_DEC, FF_H_CRC, FF_H_FEC0, FF_H_FEC1, FF_H_MOD);
    q->header_sym_len = lqo_qpilot_frame_len(q->header_mod_len, FF_PILOT_SPACING);
    q->header_sym = (lqo_cf *)calloc(q->header_sym_len, sizeof(lqo_cf));
    q->header_mod = (lqo_cf *)calloc(q->header_mod_len, sizeof(lqo_cf));
    lqo_flexframesync_reset(q);
    return q;
}

void lqo_flexframesync_destroy(lqo_flexframesync q)
{
    if (!q) return;
    lqo_qdetector_destroy(q->det);
    free(q->header_sym); free(q->header_mod); free(q->payload_sym); free(q->payload_dec); free(q);
}

void lqo_flexframesync_reset(lqo_flexframesync q)
{
    lqo_qdetector_reset(q->det);
    lqo_nco_reset(&q->mixer);
    lqo_nco_reset(&q->pll);
    memset(q->win, 0, sizeof q->win);
    q->state = FS_DETECT;
    q->mf_counter = 0; q->pfb_index = 0;
    q->preamble_counter = 0; q->symbol_counter = 0;
    q->evm_acc = 0.0f;
}

/* mixer -> matched-filter bank -> decimate by k=2; returns 1 when a symbol is produced */
static int fs_step_(lqo_flexframesync q, lqo_cf x, lqo_cf *y)
{
    const unsigned L = 2 * FF_K * FF_M;
    lqo_cf v = lqo_nco_mix_down(&q->mixer, x);
    lqo_nco_step(&q->mixer);
    memmove(q->win, q->win + 1, (L - 1) * sizeof(lqo_cf));
    q->win[L - 1] = v;
    q->mf_counter++;
    if (q->mf_counter < 1) return 0;
    const float *h = q->banks + q->pfb_index * L;
    float ar = 0.0f, ai = 0.0f;
    for (unsigned n = 0; n < L; n++) { ar = fmaf(h[n], q->win[n].re, ar); ai = fmaf(h[n], q->win[n].im, ai); }
    y->re = ar * q->mf_scale; y->im = ai * q->mf_scale;
    q->mf_counter -= FF_K;
    return 1;
}

static void fs_stats_(lqo_flexframesync q, lqo_framesyncstats *st)
{
    memset(st, 0, sizeof *st);
    st->rssi = 20.0f * log10f(q->gamma_hat);
    st->cfo = lqo_nco_get_frequency(&q->mixer);
    st->tau_hat = q->tau_hat; st->gamma_hat = q->gamma_hat;
    st->dphi_hat = q->dphi_hat; st->phi_hat = q->phi_hat; st->rxy = q->rxy;
    st->sample_index = q->frame_start;
}

static void fs_decode_header_(lqo_flexframesync q)
{
    float dphi, phi, g;
    lqo_qpilotsync(q->header_mod_len, FF_PILOT_SPACING, q->header_sym, q->header_mod, &dphi, &phi, &g);
    q->header_valid = lqo_qpm_decode(FF_H_DEC, FF_H_CRC, FF_H_FEC0, FF_H_FEC1, FF_H_MOD, q->header_mod, q->header_dec);
    if (!q->header_valid) return;
    lqo_nco_set_frequency(&q->pll, dphi);
    lqo_nco_set_phase(&q->pll, phi + dphi * (float)q->header_sym_len);
    unsigned n = FF_H_USER;
    if (q->header_dec[n] != FF_PROTOCOL) { q->header_valid = 0; return; }
    unsigned plen = ((unsigned)q->header_dec[n + 1] << 8) | q->header_dec[n + 2];
    int ms = q->header_dec[n + 3];
    int check = (q->header_dec[n + 4] >> 5) & 0x07;
    int fec0 = q->header_dec[n + 4] & 0x1f, fec1 = q->header_dec[n + 5] & 0x1f;
    if (ms == LQ_MODEM_UNKNOWN || ms >= LQ_MODEM_NUM_SCHEMES || !lqo_modem_supported(ms)) { q->header_valid = 0; return; }
    if (check == LQ_CRC_UNKNOWN || check >= LQ_CRC_NUM_SCHEMES) { q->header_valid = 0; return; }
    if (fec0 == LQ_FEC_UNKNOWN || fec0 >= LQ_FEC_NUM_SCHEMES || (fec0 != LQ_FEC_NONE && lqo_fec_enc_len(fec0, 8) == 0)) { q->header_valid = 0; return; }
    if (fec1 == LQ_FEC_UNKNOWN || fec1 >= LQ_FEC_NUM_SCHEMES || (fec1 != LQ_FEC_NONE && lqo_fec_enc_len(fec1, 8) == 0)) { q->header_valid = 0; return; }
    q->payload_dec_len = plen; q->ms = ms; q->check = check; q->fec0 = fec0; q->fec1 = fec1;
    lqo_modem_init(&q->payload_demod, ms);
    q->payload_sym_len = lqo_qpm_frame_len(plen, check, fec0, fec1, ms);
    if (q->payload_sym_len + 1 > q->payload_cap) {
        q->payload_cap = q->payload_sym_len + 1;
        q->payload_sym = (lqo_cf *)realloc(q->payload_sym, q->payload_cap * sizeof(lqo_cf));
    }
    q->payload_dec = (uint8_t *)realloc(q->payload_dec, plen + 8);
}

static void fs_rx_(lqo_flexframesync q, lqo_cf x)
{
    lqo_cf y;
    if (!fs_step_(q, x, &y)) return;
    if (q->state == FS_RXPREAMBLE) {
        q->preamble_counter++;
        if (q->preamble_counter == FF_PREAMBLE + 2 * FF_M) q->state = FS_RXHEADER;
        return;
    }
    if (q->state == FS_RXHEADER) {
        q->header_sym[q->symbol_counter++] = y;
        if (q->symbol_counter < q->header_sym_len) return;
        fs_decode_header_(q);
        if (q->header_valid) { q->symbol_counter = 0; q->state = FS_RXPAYLOAD; return; }
        lqo_framesyncstats st;
        fs_stats_(q, &st);
        if (q->cb) q->cb(q->header_dec, 0, NULL, 0, 0, st, q->ud);
        lqo_flexframesync_reset(q);
        return;
    }
    /* FS_RXPAYLOAD */
    lqo_cf v = lqo_nco_mix_down(&q->pll, y);
    q->payload_sym[q->symbol_counter] = v;
    (void)lqo_modem_demodulate(&q->payload_demod, v);
    float phase_error = lqo_modem_phase_error(&q->payload_demod);
    float evm = lqo_modem_evm(&q->payload_demod);
    lqo_nco_pll_step(&q->pll, phase_error);
    lqo_nco_step(&q->pll);
    q->evm_acc += evm * evm;
    q->symbol_counter++;
    if (q->symbol_counter < q->payload_sym_len) return;
    int ok = q->soft ? lqo_qpm_decode_soft(q->payload_dec_len, q->check, q->fec0, q->fec1, q->ms, q->payload_sym, q->payload_dec)
                     : lqo_qpm_decode(q->payload_dec_len, q->check, q->fec0, q->fec1, q->ms, q->payload_sym, q->payload_dec);
    lqo_framesyncstats st;
    fs_stats_(q, &st);
    st.evm = 10.0f * log10f(q->evm_acc / (float)q->payload_sym_len);
    st.framesyms = q->payload_sym; st.num_framesyms = q->payload_sym_len;
    st.mod_scheme = (unsigned)q->ms; st.mod_bps = lqo_modem_bps(q->ms);
    st.check = (unsigned)q->check; st.fec0 = (unsigned)q->fec0; st.fec1 = (unsigned)q->fec1;
    if (q->cb) q->cb(q->header_dec, 1, q->payload_dec, q->payload_dec_len, ok, st, q->ud);
    lqo_flexframesync_reset(q);
}

void lqo_flexframesync_execute(lqo_flexframesync q, const lqo_cf *x, unsigned n)
{
    for (unsigned i = 0; i < n; i++) {
        uint64_t idx = q->n_consumed++;
        if (q->state != FS_DETECT) { fs_rx_(q, x[i]); continue; }
        const lqo_cf *v = lqo_qdetector_execute(q->det, x[i]);
        if (!v) continue;
        unsigned nbuf = lqo_qdetector_get_buf_len(q->det);
        q->tau_hat = lqo_qdetector_get_tau(q->det);
        q->gamma_hat = lqo_qdetector_get_gamma(q->det);
        q->dphi_hat = lqo_qdetector_get_dphi(q->det);
        q->phi_hat = lqo_qdetector_get_phi(q->det);
        q->rxy = lqo_qdetector_get_rxy(q->det);
        q->frame_start = idx + 1 - nbuf;        /* may wrap "negative" if the frame began before sample 0 */
        if (q->tau_hat > 0.0f) {
            q->pfb_index = (unsigned)(int)(q->tau_hat * (float)FF_NPFB) % FF_NPFB;
            q->mf_counter = 0;
        } else {
            q->pfb_index = (unsigned)(int)((1.0f + q->tau_hat) * (float)FF_NPFB) % FF_NPFB;
            q->mf_counter = 1;
        }
        q->mf_scale = 0.5f / q->gamma_hat;
        lqo_nco_set_frequency(&q->mixer, q->dphi_hat);
        lqo_nco_set_phase(&q->mixer, q->phi_hat);
        q->state = FS_RXPREAMBLE;
        lqo_cf replay[512];
        memcpy(replay, v, nbuf * sizeof(lqo_cf));
        for (unsigned j = 0; j < nbuf; j++) fs_rx_(q, replay[j]);
    }
}

/* ================================================================== collectors */
typedef struct {
    lqo_frame_record *recs; unsigned max_frames, n;
    uint8_t *ppool; uint64_t pcap, poff;
    lqo_cf *spool; uint64_t scap, soff;
    uint64_t n_valid;
} collect_t;

static int collect_cb_(const uint8_t *header, int hv, const uint8_t *payload, unsigned plen, int pv,
                       lqo_framesyncstats st, void *ud)
{
    collect_t *c = (collect_t *)ud;
    if (pv) c->n_valid++;
    if (c->recs && c->n < c->max_frames) {
        lqo_frame_record *r = &c->recs[c->n];
        memset(r, 0, sizeof *r);
        r->sample_index = st.sample_index;
        r->header_valid = hv; r->payload_valid = pv; r->payload_len = plen;
        r->num_framesyms = st.num_framesyms;
        r->mod_scheme = st.mod_scheme; r->mod_bps = st.mod_bps; r->check = st.check; r->fec0 = st.fec0; r->fec1 = st.fec1;
        r->evm = st.evm; r->rssi = st.rssi; r->cfo = st.cfo;
        r->tau_hat = st.tau_hat; r->gamma_hat = st.gamma_hat; r->dphi_hat = st.dphi_hat; r->phi_hat = st.phi_hat; r->rxy = st.rxy;
        memcpy(r->header, header, FF_H_DEC);
        r->payload_off = c->poff; r->syms_off = c->soff;
        if (payload && c->ppool && c->poff + plen <= c->pcap) { memcpy(c->ppool + c->poff, payload, plen); c->poff += plen; }
        if (st.framesyms && c->spool && c->soff + st.num_framesyms <= c->scap) {
            memcpy(c->spool + c->soff, st.framesyms, st.num_framesyms * sizeof(lqo_cf)); c->soff += st.num_framesyms;
        }
    }
    c->n++;
    return 0;
}

static unsigned rx_capture_(const lqo_cf *x, uint64_t n, unsigned chunk, lqo_frame_record *recs, unsigned max_frames,
                            uint8_t *ppool, uint64_t pcap, lqo_cf *spool, uint64_t scap, int soft);
unsigned lqo_rx_capture(const lqo_cf *x, uint64_t n, unsigned chunk, lqo_frame_record *recs, unsigned max_frames,
                        uint8_t *ppool, uint64_t pcap, lqo_cf *spool, uint64_t scap)
{
    return rx_capture_(x, n, chunk, recs, max_frames, ppool, pcap, spool, scap, 0);
}
unsigned lqo_rx_capture_soft(const lqo_cf *x, uint64_t n, unsigned chunk, lqo_frame_record *recs, unsigned max_frames,
                             uint8_t *ppool, uint64_t pcap, lqo_cf *spool, uint64_t scap)
{
    return rx_capture_(x, n, chunk, recs, max_frames, ppool, pcap, spool, scap, 1);
}
static unsigned rx_capture_(const lqo_cf *x, uint64_t n, unsigned chunk, lqo_frame_record *recs, unsigned max_frames,
                            uint8_t *ppool, uint64_t pcap, lqo_cf *spool, uint64_t scap, int soft)
{
    collect_t c; memset(&c, 0, sizeof c);
    c.recs = recs; c.max_frames = max_frames; c.ppool = ppool; c.pcap = pcap; c.spool = spool; c.scap = scap;
    lqo_flexframesync fs = lqo_flexframesync_create(collect_cb_, &c);
    lqo_flexframesync_set_soft(fs, soft);
    if (!chunk) chunk = 256;
    for (uint64_t i = 0; i < n; i += chunk) {
        unsigned m = (n - i < chunk) ? (unsigned)(n - i) : chunk;
        lqo_flexframesync_execute(fs, x + i, m);
    }
    lqo_flexframesync_destroy(fs);
    return c.n;
}

typedef struct { const lqo_cf *x; unsigned n_streams, tid, nth; uint64_t stride, n, frames, valid; } many_t;
static void *many_worker_(void *arg)
{
    many_t *w = (many_t *)arg;
    for (unsigned s = w->tid; s < w->n_streams; s += w->nth) {
        collect_t c; memset(&c, 0, sizeof c);
        lqo_flexframesync fs = lqo_flexframesync_create(collect_cb_, &c);
        const lqo_cf *p = w->x + (uint64_t)s * w->stride;
        for (uint64_t i = 0; i < w->n; i += 256) {
            unsigned m = (w->n - i < 256) ? (unsigned)(w->n - i) : 256;
            lqo_flexframesync_execute(fs, p + i, m);
        }
        lqo_flexframesync_destroy(fs);
        w->frames += c.n; w->valid += c.n_valid;
    }
    return NULL;
}

uint64_t lqo_rx_many(const lqo_cf *x, unsigned n_streams, uint64_t stride, uint64_t n, unsigned nth, uint64_t *n_valid)
{
    if (nth < 1) nth = 1;
    if (nth > 256) nth = 256;
    pthread_t th[256]; many_t w[256];
    (void)lqo_nco_sintab();
    for (unsigned t = 0; t < nth; t++) {
        memset(&w[t], 0, sizeof w[t]);
        w[t].x = x; w[t].n_streams = n_streams; w[t].tid = t; w[t].nth = nth; w[t].stride = stride; w[t].n = n;
        pthread_create(&th[t], NULL, many_worker_, &w[t]);
    }
    uint64_t frames = 0, valid = 0;
    for (unsigned t = 0; t < nth; t++) { pthread_join(th[t], NULL); frames += w[t].frames; valid += w[t].valid; }
    if (n_valid) *n_valid = valid;
    return frames;
}

unsigned lqo_detect_capture(const lqo_cf *x, uint64_t n, float beta, float threshold, lqo_detection *out, unsigned max_out)
{
    lqo_cf pn[FF_PREAMBLE];
    preamble_pn_(pn);
    lqo_qdetector d = lqo_qdetector_create_linear(pn, FF_PREAMBLE, LQ_FIRFILT_ARKAISER, FF_K, FF_M, beta);
    lqo_qdetector_set_threshold(d, threshold);
    unsigned cnt = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (!lqo_qdetector_execute(d, x[i])) continue;
        if (out && cnt < max_out) {
            out[cnt].sample_index = i + 1 - 512;
            out[cnt].tau_hat = d->tau_hat; out[cnt].gamma_hat = d->gamma_hat;
            out[cnt].dphi_hat = d->dphi_hat; out[cnt].phi_hat = d->phi_hat; out[cnt].rxy = d->rxy;
        }
        cnt++;
    }
    lqo_qdetector_destroy(d);
    return cnt;
}

unsigned lqo_tx_frame(int ms, int check, int fec0, int fec1, const uint8_t *header14,
                      const uint8_t *payload, unsigned payload_len, lqo_cf *out, unsigned out_cap)
{
    lqo_fgprops p = { check, fec0, fec1, ms };
    lqo_flexframegen fg = lqo_flexframegen_create(&p);
    lqo_flexframegen_assemble(fg, header14, payload, payload_len);
    unsigned n = lqo_flexframegen_getframelen(fg);
    if (out && n <= out_cap) lqo_flexframegen_write_samples(fg, out, n);
    lqo_flexframegen_destroy(fg);
    return n;
}
