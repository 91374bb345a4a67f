/*
 * lqo_math.c -- ORACLE (test infrastructure only; see lqo.h header).
 * m-sequence, ARKAISER filter design, polyphase banks, radix-2 FFT, NCO/PLL.
 * Follows SURVEY.md Appendix A.2, A.4, A.5 (liquid-dsp call sites:
 * /root/reference/lib/frame_detector_cc_impl.cc:47-54).
 */
#include "lqo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ m-sequence */
void lqo_mseq_init(lqo_mseq *q, unsigned m, unsigned g, unsigned a)
{
    q->m = m;
    q->g = g >> 1;
    q->a = 0;
    for (unsigned i = 0; i < m; i++) { q->a = (q->a << 1) | (a & 1u); a >>= 1; }
    q->n = (1u << m) - 1u;
    q->v = q->a;
    q->b = 0;
}

void lqo_mseq_init_default(lqo_mseq *q, unsigned m)
{
    static const unsigned genpoly[16] = { 0, 0, 0x7, 0xB, 0x13, 0x25, 0x43, 0x89, 0x11D,
                                          0x211, 0x409, 0x805, 0x1053, 0x201B, 0x402B, 0x8003 };
    lqo_mseq_init(q, m, genpoly[m & 15], 1);
}

unsigned lqo_mseq_advance(lqo_mseq *q)
{
    q->b = (unsigned)__builtin_parity(q->v & q->g);
    q->v = ((q->v << 1) | q->b) & q->n;
    return q->b;
}

unsigned lqo_mseq_symbol(lqo_mseq *q, unsigned bps)
{
    unsigned s = 0;
    for (unsigned i = 0; i < bps; i++) s = (s << 1) | lqo_mseq_advance(q);
    return s;
}

/* ------------------------------------------------------------------ filter design */
static float besseli0f_(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    for (unsigned k = 0; k < 32; k++) {
        float t = (float)k * logf(0.5f * z) - lgammaf((float)k + 1.0f);
        y += expf(2.0f * t);
    }
    return y;
}

static float kaiser_window_(unsigned i, unsigned n, float beta, float mu)
{
    float t = (float)i - (float)(n - 1) / 2.0f + mu;
    float r = 2.0f * t / (float)n;
    float a = besseli0f_(beta * sqrtf(1.0f - r * r));
    float b = besseli0f_(beta);
    return a / b;
}

static float sincf_(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf((float)M_PI * x / 2.0f) * cosf((float)M_PI * x / 4.0f) * cosf((float)M_PI * x / 8.0f);
    return sinf((float)M_PI * x) / ((float)M_PI * x);
}

static float kaiser_beta_As_(float As)
{
    As = fabsf(As);
    if (As > 50.0f) return 0.1102f * (As - 8.7f);
    if (As > 21.0f) return 0.5842f * powf(As - 21.0f, 0.4f) + 0.07886f * (As - 21.0f);
    return 0.0f;
}

static float filter_len_kaiser_(float df, float As) { return (As - 7.95f) / (14.26f * df); }

static float filter_len_herrmann_(float df, float As)
{
    if (As > 105.0f) return filter_len_kaiser_(df, As);
    As += 7.4f;
    float d = powf(10.0f, -As / 20.0f);
    float t1 = log10f(d), t2 = log10f(d);
    float Dinf = (0.005309f * t1 * t1 + 0.07114f * t1 - 0.4761f) * t2
               - (0.002660f * t1 * t1 + 0.5941f * t1 + 0.4278f);
    float f = 11.012f + 0.51244f * (t1 - t2);
    return (Dinf - f * df * df) / df + 1.0f;
}

static float req_filter_As_(float df, unsigned N)
{
    float As0 = 0.01f, As1 = 200.0f, As_hat = 0.0f;
    for (unsigned i = 0; i < 20; i++) {
        As_hat = 0.5f * (As1 + As0);
        float N_hat = filter_len_herrmann_(df, As_hat);
        if (N_hat < (float)N) As0 = As_hat; else As1 = As_hat;
    }
    return As_hat;
}

static float rkaiser_approximate_rho_(unsigned m, float beta)
{
    float c0 = 0.762886f + 0.067663f * logf((float)m);
    float c1 = 0.065515f;
    float c2 = logf(1.0f - 0.088f * powf((float)m, -1.6f));
    float lb = logf(beta);
    float rho = c0 + c1 * lb + c2 * lb * lb;
    if (rho < 0.0f) rho = 0.0f;
    if (rho > 1.0f) rho = 1.0f;
    return rho;
}

void lqo_firdes_arkaiser(unsigned k, unsigned m, float beta, float dt, float *h)
{
    float c0 = 0.762886f + 0.067663f * logf((float)m);
    float c1 = 0.065515f;
    float c2 = logf(1.0f - 0.088f * powf((float)m, -1.6f));
    float lb = logf(beta);
    float rho = c0 + c1 * lb + c2 * lb * lb;
    if (rho <= 0.0f || rho >= 1.0f) rho = rkaiser_approximate_rho_(m, beta);

    unsigned n = 2 * k * m + 1;
    float kf = (float)k;
    float del = beta * rho / kf;
    float As = req_filter_As_(del, n);
    float fc = 0.5f * (1.0f + beta * (1.0f - rho)) / kf;

    float kb = kaiser_beta_As_(As);
    for (unsigned i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2.0f + dt;
        h[i] = sincf_(2.0f * fc * t) * kaiser_window_(i, n, kb, dt);
    }
    float e2 = 0.0f;
    for (unsigned i = 0; i < n; i++) e2 += h[i] * h[i];
    float s = sqrtf(kf / e2);
    for (unsigned i = 0; i < n; i++) h[i] *= s;
}

void lqo_interp_taps(unsigned k, unsigned m, float beta, float *h)
{
    unsigned n = 2 * k * m + 1;
    lqo_firdes_arkaiser(k, m, beta, 0.0f, h);
    unsigned sub = 0;
    while (k * sub < n) sub++;
    for (unsigned i = n; i < k * sub; i++) h[i] = 0.0f;
}

void lqo_pfb_rnyquist(unsigned npfb, unsigned k, unsigned m, float beta, float *banks)
{
    unsigned H_len = 2 * npfb * k * m + 1;
    float *H = (float *)malloc(H_len * sizeof(float));
    lqo_firdes_arkaiser(npfb * k, m, beta, 0.0f, H);
    unsigned sub = H_len / npfb; /* = 2km, last prototype tap dropped */
    for (unsigned i = 0; i < npfb; i++)
        for (unsigned n = 0; n < sub; n++)
            banks[i * sub + (sub - 1 - n)] = H[i + n * npfb]; /* stored oldest->newest */
    free(H);
}

/* ------------------------------------------------------------------ FFT */
/* twiddle cache: W[k] = exp(-j 2 pi k / n), k < n/2, rounded from double */
static lqo_cf *twiddles_(unsigned n)
{
    static lqo_cf *cache[16];
    unsigned lg = 0;
    while ((1u << lg) < n) lg++;
    lqo_cf *w = __atomic_load_n(&cache[lg], __ATOMIC_ACQUIRE);
    if (w) return w;
    w = (lqo_cf *)malloc((n / 2 + 1) * sizeof(lqo_cf));
    for (unsigned k = 0; k < n / 2; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        w[k].re = (float)cos(a);
        w[k].im = (float)(-sin(a));
    }
    if (n >= 4) { w[n / 4].re = 0.0f; w[n / 4].im = -1.0f; } /* exact quarter turn (cos(pi/2) is not 0 in double) */
    lqo_cf *expected = NULL;
    if (!__atomic_compare_exchange_n(&cache[lg], &expected, w, 0, __ATOMIC_RELEASE, __ATOMIC_ACQUIRE)) {
        free(w);
        w = expected;
    }
    return w;
}

static const unsigned short *bitrev_(unsigned n, unsigned lg)
{
    static unsigned short *cache[16];
    unsigned short *t = __atomic_load_n(&cache[lg], __ATOMIC_ACQUIRE);
    if (t) return t;
    t = (unsigned short *)malloc(n * sizeof(unsigned short));
    for (unsigned i = 0; i < n; i++) {
        unsigned r = 0;
        for (unsigned b = 0; b < lg; b++) r |= ((i >> b) & 1u) << (lg - 1 - b);
        t[i] = (unsigned short)r;
    }
    unsigned short *expected = NULL;
    if (!__atomic_compare_exchange_n(&cache[lg], &expected, t, 0, __ATOMIC_RELEASE, __ATOMIC_ACQUIRE)) { free(t); t = expected; }
    return t;
}

void lqo_fft(const lqo_cf *in, lqo_cf *out, unsigned n, int dir)
{
    unsigned lg = 0;
    while ((1u << lg) < n) lg++;
    const lqo_cf *W = twiddles_(n);
    lqo_cf tmp[512];
    lqo_cf *a = (in == out) ? tmp : out;
    const unsigned short *rev = bitrev_(n, lg);
    for (unsigned i = 0; i < n; i++) a[i] = in[rev[i]];
    for (unsigned s = 1; s <= lg; s++) {
        unsigned m = 1u << s, half = m >> 1, step = n / m;
        for (unsigned k = 0; k < n; k += m) {
            for (unsigned j = 0; j < half; j++) {
                float wr = W[j * step].re;
                float wi = (dir > 0) ? W[j * step].im : -W[j * step].im;
                lqo_cf b = a[k + j + half], u = a[k + j], t;
                t.re = fmaf(-wi, b.im, wr * b.re);
                t.im = fmaf(wi, b.re, wr * b.im);
                a[k + j].re = u.re + t.re;
                a[k + j].im = u.im + t.im;
                a[k + j + half].re = u.re - t.re;
                a[k + j + half].im = u.im - t.im;
            }
        }
    }
    if (a == tmp) memcpy(out, tmp, n * sizeof(lqo_cf));
}

/* ------------------------------------------------------------------ NCO / PLL */
static float g_sintab[1024];
static int g_sintab_ready;

const float *lqo_nco_sintab(void)
{
    if (!__atomic_load_n(&g_sintab_ready, __ATOMIC_ACQUIRE)) {
        for (unsigned i = 0; i < 1024; i++)
            g_sintab[i] = sinf(2.0f * (float)M_PI * (float)i / 1024.0f);
        __atomic_store_n(&g_sintab_ready, 1, __ATOMIC_RELEASE);
    }
    return g_sintab;
}

void lqo_nco_reset(lqo_nco *q) { q->theta = 0; q->d_theta = 0; }

uint32_t lqo_nco_constrain(float theta)
{
    float p = theta * 0.15915494309189535f;     /* 1/(2 pi) */
    float fpart = p - (float)((long long)p);    /* in (-1,1) */
    if (fpart < 0.0f) fpart += 1.0f;
    return (uint32_t)(long long)(fpart * 4294967296.0f);
}

void lqo_nco_set_frequency(lqo_nco *q, float d) { q->d_theta = lqo_nco_constrain(d); }
void lqo_nco_set_phase(lqo_nco *q, float t) { q->theta = lqo_nco_constrain(t); }

float lqo_nco_get_frequency(const lqo_nco *q)
{
    float d = (float)q->d_theta * 1.4629180792671596e-09f; /* 2 pi / 2^32 */
    return d > (float)M_PI ? d - 2.0f * (float)M_PI : d;
}

void lqo_nco_pll_set_bandwidth(lqo_nco *q, float bw) { q->alpha = bw; q->beta = sqrtf(bw); }

void lqo_nco_pll_step(lqo_nco *q, float dphi)
{
    q->d_theta += lqo_nco_constrain(dphi * q->alpha);
    q->theta   += lqo_nco_constrain(dphi * q->beta);
}

void lqo_nco_step(lqo_nco *q) { q->theta += q->d_theta; }

lqo_cf lqo_nco_mix_down(const lqo_nco *q, lqo_cf x)
{
    const float *tab = lqo_nco_sintab();
    unsigned idx = ((q->theta + (1u << 21)) >> 22) & 0x3ffu;
    float s = tab[idx], c = tab[(idx + 256u) & 0x3ffu];
    lqo_cf y; /* x * (c - j s) */
    y.re = fmaf(x.im, s, x.re * c);
    y.im = fmaf(-x.re, s, x.im * c);
    return y;
}
