/*
 * lqo_modem.c -- ORACLE (test infrastructure only; see lqo.h header).
 * Linear modems (PSK, DPSK, ASK, QAM, BPSK, QPSK), hard decision, and the
 * qpacketmodem that joins packetizer + modem.  SURVEY.md Appendix A.6.
 * Scheme numbers reachable through the reference's block API:
 * /root/reference/lib/flex_tx_impl.cc:77-115, lib/flex_rx_impl.cc:139-178.
 */
#include "lqo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static unsigned gray_encode_(unsigned s) { return s ^ (s >> 1); }
static unsigned gray_decode_(unsigned s)
{
    unsigned r = s;
    for (unsigned sh = 1; sh < 32; sh <<= 1) r ^= r >> sh;
    return r;
}

int lqo_modem_supported(int ms)
{
    return (ms >= LQ_MODEM_PSK2 && ms <= LQ_MODEM_QAM256) || ms == LQ_MODEM_BPSK || ms == LQ_MODEM_QPSK;
}

unsigned lqo_modem_bps(int ms)
{
    if (ms >= LQ_MODEM_PSK2 && ms <= LQ_MODEM_PSK256) return (unsigned)(ms - LQ_MODEM_PSK2 + 1);
    if (ms >= LQ_MODEM_DPSK2 && ms <= LQ_MODEM_DPSK256) return (unsigned)(ms - LQ_MODEM_DPSK2 + 1);
    if (ms >= LQ_MODEM_ASK2 && ms <= LQ_MODEM_ASK256) return (unsigned)(ms - LQ_MODEM_ASK2 + 1);
    if (ms >= LQ_MODEM_QAM4 && ms <= LQ_MODEM_QAM256) return (unsigned)(ms - LQ_MODEM_QAM4 + 2);
    if (ms == LQ_MODEM_BPSK) return 1;
    if (ms == LQ_MODEM_QPSK) return 2;
    return 0;
}

static int is_psk_(int ms)  { return ms >= LQ_MODEM_PSK2 && ms <= LQ_MODEM_PSK256; }
static int is_dpsk_(int ms) { return ms >= LQ_MODEM_DPSK2 && ms <= LQ_MODEM_DPSK256; }
static int is_ask_(int ms)  { return ms >= LQ_MODEM_ASK2 && ms <= LQ_MODEM_ASK256; }
static int is_qam_(int ms)  { return ms >= LQ_MODEM_QAM4 && ms <= LQ_MODEM_QAM256; }

static lqo_cf cexpj_(float t) { lqo_cf y = { cosf(t), sinf(t) }; return y; }

/* arg() and exp(j t) INSIDE the per-symbol receive / transmit loops (generic PSK and DPSK modems).  liquid-dsp calls the
 * platform's cargf / cexpf there, so its own results move in the last bit with the libm version (glibc 2.39's atan2f
 * differs from the correctly rounded value for 16 % of random arguments); that last bit reaches the PLL phase and moves
 * constellation points across a step of the NCO table.  The oracle therefore pins ONE implementation -- Cephes' single
 * precision atanf / sinf / cosf (S. Moshier, public domain algorithm: argument reduction + the polynomials below), every
 * operation a plain IEEE float operation in a fixed order (this file is compiled with -ffp-contract=off) -- and the CUDA
 * path (csrc/lqb_dev.cuh: pm_atan2f / pm_sincosf) executes the same sequence, so the two agree bit for bit
 * (tests/test_gpu_parity.py::test_pinned_arg_and_sincos_match_the_oracle_bit_for_bit).  Within 2 ulp of libm. */
static float pm_atanf_pos_(float x)                 /* x >= 0 (may be +inf) */
{
    float y;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    else y = 0.0f;
    const float z = x * x;
    const float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
    return y + p;
}
float lqo_pm_atan2f(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    float r = (ax == 0.0f && ay == 0.0f) ? 0.0f : pm_atanf_pos_(ay / ax);
    if (x < 0.0f) r = 3.14159274f - r;
    return y < 0.0f ? -r : r;
}
void lqo_pm_sincosf(float t, float *sn, float *cs)     /* |t| < 8192 */
{
    float x = fabsf(t);
    int j = (int)(1.27323954473516f * x);               /* 4 / pi */
    if (j & 1) j += 1;
    const float y = (float)j;
    j &= 7;
    int s_neg = t < 0.0f, c_neg = 0;
    if (j > 3) { s_neg = !s_neg; c_neg = !c_neg; j -= 4; }
    if (j > 1) c_neg = !c_neg;
    x = ((x - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    const float z = x * x;
    const float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    const float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    const int swap = (j == 1 || j == 2);
    const float s = swap ? pc : ps, c = swap ? ps : pc;
    *sn = s_neg ? -s : s;
    *cs = c_neg ? -c : c;
}
static lqo_cf cexpj_pm_(float t) { lqo_cf y; lqo_pm_sincosf(t, &y.im, &y.re); return y; }

static lqo_cf modulate_raw_(lqo_modem *q, unsigned s)
{
    lqo_cf y = { 0.0f, 0.0f };
    int ms = q->scheme;
    if (is_psk_(ms)) {
        y = cexpj_((float)gray_decode_(s) * 2.0f * q->alpha);
    } else if (is_ask_(ms)) {
        y.re = (float)(2 * (int)gray_decode_(s) - (int)q->M + 1) * q->alpha;
    } else if (is_qam_(ms)) {
        unsigned si = gray_decode_(s >> q->m_q), sq = gray_decode_(s & ((1u << q->m_q) - 1u));
        y.re = (float)(2 * (int)si - (int)(1u << q->m_i) + 1) * q->alpha;
        y.im = (float)(2 * (int)sq - (int)(1u << q->m_q) + 1) * q->alpha;
    } else if (ms == LQ_MODEM_BPSK) {
        y.re = s ? -1.0f : 1.0f;
    } else if (ms == LQ_MODEM_QPSK) {
        y.re = (s & 1u) ? -(float)M_SQRT1_2 : (float)M_SQRT1_2;
        y.im = (s & 2u) ? -(float)M_SQRT1_2 : (float)M_SQRT1_2;
    }
    return y;
}

int lqo_modem_init(lqo_modem *q, int ms)
{
    memset(q, 0, sizeof *q);
    if (!lqo_modem_supported(ms)) return -1;
    q->scheme = ms;
    q->bps = lqo_modem_bps(ms);
    q->M = 1u << q->bps;
    if (is_psk_(ms) || is_dpsk_(ms)) {
        q->alpha = (float)M_PI / (float)q->M;
        q->d_phi = (float)M_PI * (1.0f - 1.0f / (float)q->M);
    } else if (is_ask_(ms)) {
        static const float c[9] = { 0, 1.0f, 5.0f, 21.0f, 85.0f, 341.0f, 1365.0f, 5461.0f, 21845.0f };
        q->alpha = 1.0f / sqrtf(c[q->bps]);
    } else if (is_qam_(ms)) {
        static const float c[9] = { 0, 0, 2.0f, 6.0f, 10.0f, 26.0f, 42.0f, 106.0f, 170.0f };
        q->m_i = (q->bps + 1) >> 1;
        q->m_q = q->bps >> 1;
        q->alpha = 1.0f / sqrtf(c[q->bps]);
    }
    for (unsigned k = 0; k < q->bps && k < 8; k++) q->ref[k] = (float)(1u << k) * q->alpha;
    if (!is_dpsk_(ms))
        for (unsigned s = 0; s < q->M; s++) q->map[s] = modulate_raw_(q, s);
    q->x_hat.re = 1.0f;
    return 0;
}

void lqo_modem_reset(lqo_modem *q) { q->dpsk_phi = 0.0f; q->x_hat.re = 1.0f; q->x_hat.im = 0.0f; q->r = q->x_hat; }

lqo_cf lqo_modem_modulate(lqo_modem *q, unsigned s)
{
    if (is_dpsk_(q->scheme)) {
        q->dpsk_phi += (float)gray_decode_(s) * 2.0f * q->alpha;
        if (q->dpsk_phi > 2.0f * (float)M_PI) q->dpsk_phi -= 2.0f * (float)M_PI;
        return cexpj_pm_(q->dpsk_phi);
    }
    return q->map[s & (q->M - 1u)];
}

/* successive-approximation slicer over ref[k] = 2^k * alpha */
static void slice_(float v, unsigned m, const float *ref, unsigned *s_out, float *res)
{
    unsigned s = 0, k = m;
    for (unsigned i = 0; i < m; i++) {
        s <<= 1;
        s |= (v > 0.0f);
        float r = ref[--k];
        v += (v > 0.0f) ? -r : r;
    }
    *s_out = s; *res = v;
}

unsigned lqo_modem_demodulate(lqo_modem *q, lqo_cf x)
{
    int ms = q->scheme;
    unsigned s = 0, sym = 0;
    float res;
    if (is_psk_(ms)) {
        float theta = lqo_pm_atan2f(x.im, x.re) - q->d_phi;
        if (theta < -(float)M_PI) theta += 2.0f * (float)M_PI;
        slice_(theta, q->bps, q->ref, &s, &res);
        sym = gray_encode_(s);
        q->x_hat = q->map[sym];
    } else if (is_dpsk_(ms)) {
        float theta = lqo_pm_atan2f(x.im, x.re);
        float d = theta - q->dpsk_phi;
        q->dpsk_phi = theta;
        d -= q->d_phi;
        if (d > (float)M_PI) d -= 2.0f * (float)M_PI;
        else if (d < -(float)M_PI) d += 2.0f * (float)M_PI;
        slice_(d, q->bps, q->ref, &s, &res);
        sym = gray_encode_(s);
        q->x_hat = cexpj_pm_(theta - res);
    } else if (is_ask_(ms)) {
        slice_(x.re, q->bps, q->ref, &s, &res);
        sym = gray_encode_(s);
        q->x_hat = q->map[sym];
    } else if (is_qam_(ms)) {
        unsigned si, sq; float ri, rq;
        slice_(x.re, q->m_i, q->ref, &si, &ri);
        slice_(x.im, q->m_q, q->ref, &sq, &rq);
        sym = (gray_encode_(si) << q->m_q) + gray_encode_(sq);
        q->x_hat.re = x.re - ri;
        q->x_hat.im = x.im - rq;
    } else if (ms == LQ_MODEM_BPSK) {
        sym = (x.re > 0.0f) ? 0u : 1u;
        q->x_hat = q->map[sym];
    } else if (ms == LQ_MODEM_QPSK) {
        sym = ((x.re > 0.0f) ? 0u : 1u) + ((x.im > 0.0f) ? 0u : 2u);
        q->x_hat = q->map[sym];
    }
    q->r = x;
    return sym;
}

int lqo_modem_demodulate_soft(const lqo_modem *q, lqo_cf x, uint8_t *soft)
{
    if (is_dpsk_(q->scheme) || !q->bps) return 0;
    const unsigned M = 1u << q->bps;
    /* G = 64 / dmin^2 (every operation a single IEEE operation, in this order: the CUDA kernel repeats it) */
    float dmin2 = 3.0e38f;
    for (unsigned a = 0; a < M; a++)
        for (unsigned b = a + 1; b < M; b++) {
            float dr = q->map[a].re - q->map[b].re, di = q->map[a].im - q->map[b].im;
            float d = fmaf(di, di, dr * dr);
            if (d < dmin2) dmin2 = d;
        }
    const float G = 64.0f / dmin2;
    float d0[8], d1[8];
    for (unsigned k = 0; k < q->bps; k++) { d0[k] = 3.0e38f; d1[k] = 3.0e38f; }
    for (unsigned s = 0; s < M; s++) {
        float dr = x.re - q->map[s].re, di = x.im - q->map[s].im;
        float d = fmaf(di, di, dr * dr);
        for (unsigned k = 0; k < q->bps; k++) {
            if ((s >> (q->bps - 1 - k)) & 1u) { if (d < d1[k]) d1[k] = d; }
            else { if (d < d0[k]) d0[k] = d; }
        }
    }
    for (unsigned k = 0; k < q->bps; k++) {
        float t = fmaf(G, d0[k] - d1[k], 128.0f);
        soft[k] = (uint8_t)(t <= 0.0f ? 0 : t >= 255.0f ? 255 : (int)t);
    }
    return 1;
}

float lqo_modem_phase_error(const lqo_modem *q)
{   /* imag( r * conj(x_hat) ) */
    return fmaf(q->r.im, q->x_hat.re, -(q->r.re * q->x_hat.im));
}

float lqo_modem_evm(const lqo_modem *q)
{
    float dr = q->x_hat.re - q->r.re, di = q->x_hat.im - q->r.im;
    return sqrtf(fmaf(di, di, dr * dr));
}

/* ================================================================== qpacketmodem */
unsigned lqo_qpm_frame_len(unsigned n, int check, int fec0, int fec1, int ms)
{
    unsigned bps = lqo_modem_bps(ms);
    if (!bps) return 0;
    unsigned bits = 8 * lqo_packetizer_enc_len(n, check, fec0, fec1);
    return bits / bps + (bits % bps ? 1 : 0);
}

void lqo_qpm_encode(unsigned n, int check, int fec0, int fec1, int ms, const uint8_t *payload, lqo_cf *frame)
{
    unsigned enc_len = lqo_packetizer_enc_len(n, check, fec0, fec1);
    unsigned nsym = lqo_qpm_frame_len(n, check, fec0, fec1, ms);
    uint8_t *enc = (uint8_t *)calloc(enc_len + 8, 1);
    lqo_packetizer_encode(n, check, fec0, fec1, payload, enc);
    lqo_modem mod;
    lqo_modem_init(&mod, ms);
    unsigned bps = mod.bps, nbits = 8 * enc_len;
    for (unsigned i = 0; i < nsym; i++) {
        unsigned s = 0;
        for (unsigned b = 0; b < bps; b++) {
            unsigned pos = i * bps + b;
            unsigned bit = (pos < nbits) ? ((enc[pos >> 3] >> (7 - (pos & 7))) & 1u) : 0u; /* zero-padded tail */
            s = (s << 1) | bit;
        }
        frame[i] = lqo_modem_modulate(&mod, s);
    }
    free(enc);
}

int lqo_qpm_decode_soft(unsigned n, int check, int fec0, int fec1, int ms, const lqo_cf *frame, uint8_t *payload)
{
    /* the stage nearest the channel */
    const int stage1 = (fec1 != LQ_FEC_NONE);
    const int fs = stage1 ? fec1 : fec0;
    lqo_modem mod;
    lqo_modem_init(&mod, ms);
    uint8_t probe[8];
    lqo_cf zero = { 0.0f, 0.0f };
    if (!lqo_fec_is_conv(fs) || !lqo_modem_demodulate_soft(&mod, zero, probe))
        return lqo_qpm_decode(n, check, fec0, fec1, ms, frame, payload);
    const unsigned cl = lqo_crc_len(check), k0 = n + cl;
    const unsigned n0 = lqo_fec_enc_len(fec0, k0), n1 = lqo_fec_enc_len(fec1, n0);
    const unsigned nsym = lqo_qpm_frame_len(n, check, fec0, fec1, ms);
    const unsigned bps = mod.bps, nbits = 8 * n1;             /* (n1 == n0 when fec1 is "none") */
    uint8_t *raw = (uint8_t *)calloc((size_t)nsym * bps + 8, 1), *soft = (uint8_t *)calloc(nbits + 8, 1);
    for (unsigned i = 0; i < nsym; i++) lqo_modem_demodulate_soft(&mod, frame[i], raw + (size_t)i * bps);
    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * nbits);
    lqo_deinterleave_bit_perm(n1, perm);
    for (unsigned i = 0; i < nbits; i++) soft[i] = raw[perm[i]];
    const unsigned dec_len = stage1 ? n0 : k0;
    uint8_t *b0 = (uint8_t *)calloc(dec_len + k0 + 8, 1), *b1 = (uint8_t *)calloc(dec_len + k0 + 8, 1);
    lqo_fec_decode_soft(fs, dec_len, soft, b0);
    if (stage1) {                                             /* then fec0, hard */
        lqo_deinterleave(b0, n0, fec0 == LQ_FEC_NONE ? 0 : 4);
        lqo_fec_decode(fec0, k0, b0, b1);
        memcpy(b0, b1, k0);
    }
    lqo_scramble(b0, k0);
    unsigned key = 0;
    for (unsigned i = 0; i < cl; i++) key = (key << 8) | b0[n + i];
    memcpy(payload, b0, n);
    int ok = (lqo_crc_key(check, b0, n) == key);
    free(raw); free(soft); free(perm); free(b0); free(b1);
    return ok;
}

int lqo_qpm_decode(unsigned n, int check, int fec0, int fec1, int ms, const lqo_cf *frame, uint8_t *payload)
{
    unsigned enc_len = lqo_packetizer_enc_len(n, check, fec0, fec1);
    unsigned nsym = lqo_qpm_frame_len(n, check, fec0, fec1, ms);
    uint8_t *enc = (uint8_t *)calloc(enc_len + 8, 1);
    lqo_modem mod;
    lqo_modem_init(&mod, ms);
    unsigned bps = mod.bps, nbits = 8 * enc_len;
    for (unsigned i = 0; i < nsym; i++) {
        unsigned s = lqo_modem_demodulate(&mod, frame[i]);
        for (unsigned b = 0; b < bps; b++) {
            unsigned pos = i * bps + b;
            if (pos < nbits && ((s >> (bps - 1 - b)) & 1u)) enc[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
        }
    }
    int ok = lqo_packetizer_decode(n, check, fec0, fec1, enc, payload);
    free(enc);
    return ok;
}
