// lqb_tables.cpp -- host-side table generation (see lqb_tables.h).
// Compiled with -ffp-contract=off (nvcc: -Xcompiler -ffp-contract=off) so the values are
// reproducible; fused operations are spelled std::fmaf.
#include "lqb_tables.h"
#include "lqb_lens.h"
#include <cmath>
#include <cstring>
#include <mutex>

namespace lqb {

static const float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------ approximate r-Kaiser design
namespace {
float bessel_i0(float z)
{
    if (z == 0.0f) return 1.0f;
    float acc = 0.0f;
    for (unsigned k = 0; k < 32; ++k)
        acc += std::exp(2.0f * ((float)k * std::log(0.5f * z) - std::lgamma((float)k + 1.0f)));
    return acc;
}
float kaiser_win(unsigned i, unsigned n, float beta, float mu)
{
    float t = (float)i - (float)(n - 1) / 2.0f + mu;
    float r = 2.0f * t / (float)n;
    return bessel_i0(beta * std::sqrt(1.0f - r * r)) / bessel_i0(beta);
}
float sinc(float x)
{
    if (std::fabs(x) < 0.01f)
        return std::cos(kPi * x / 2.0f) * std::cos(kPi * x / 4.0f) * std::cos(kPi * x / 8.0f);
    return std::sin(kPi * x) / (kPi * x);
}
float len_kaiser(float df, float As) { return (As - 7.95f) / (14.26f * df); }
float len_herrmann(float df, float As)
{
    if (As > 105.0f) return len_kaiser(df, As);
    As += 7.4f;
    float d = std::pow(10.0f, -As / 20.0f);
    float t1 = std::log10(d), t2 = std::log10(d);
    float Dinf = (0.005309f * t1 * t1 + 0.07114f * t1 - 0.4761f) * t2
               - (0.002660f * t1 * t1 + 0.5941f * t1 + 0.4278f);
    float f = 11.012f + 0.51244f * (t1 - t2);
    return (Dinf - f * df * df) / df + 1.0f;
}
float stopband_for(float df, unsigned N)
{
    float lo = 0.01f, hi = 200.0f, As = 0.0f;
    for (int it = 0; it < 20; ++it) {
        As = 0.5f * (hi + lo);
        if (len_herrmann(df, As) < (float)N) lo = As; else hi = As;
    }
    return As;
}
float kaiser_beta(float As)
{
    As = std::fabs(As);
    if (As > 50.0f) return 0.1102f * (As - 8.7f);
    if (As > 21.0f) return 0.5842f * std::pow(As - 21.0f, 0.4f) + 0.07886f * (As - 21.0f);
    return 0.0f;
}
}  // namespace

std::vector<float> firdes_arkaiser(unsigned k, unsigned m, float beta, float dt)
{
    const float c0 = 0.762886f + 0.067663f * std::log((float)m);
    const float c1 = 0.065515f;
    const float c2 = std::log(1.0f - 0.088f * std::pow((float)m, -1.6f));
    const float lb = std::log(beta);
    float rho = c0 + c1 * lb + c2 * lb * lb;
    if (rho <= 0.0f || rho >= 1.0f) rho = rho <= 0.0f ? 0.0f : 1.0f;

    const unsigned n = 2 * k * m + 1;
    const float kf = (float)k;
    const float del = beta * rho / kf;
    const float As = stopband_for(del, n);
    const float fc = 0.5f * (1.0f + beta * (1.0f - rho)) / kf;
    const float kb = kaiser_beta(As);

    std::vector<float> h(n);
    for (unsigned i = 0; i < n; ++i) {
        float t = (float)i - (float)(n - 1) / 2.0f + dt;
        h[i] = sinc(2.0f * fc * t) * kaiser_win(i, n, kb, dt);
    }
    float e2 = 0.0f;
    for (float v : h) e2 += v * v;
    const float s = std::sqrt(kf / e2);
    for (float &v : h) v *= s;
    return h;
}

std::vector<float> interp_taps(float beta)
{
    std::vector<float> h = firdes_arkaiser(kK, kM, beta, 0.0f);
    h.resize(30, 0.0f);
    return h;
}

std::vector<float> pfb_banks(float beta)
{
    std::vector<float> H = firdes_arkaiser(kNpfb * kK, kM, beta, 0.0f);   // 897 taps
    std::vector<float> banks(kNpfb * kTaps);
    for (unsigned b = 0; b < kNpfb; ++b)
        for (unsigned n = 0; n < kTaps; ++n) banks[b * kTaps + (kTaps - 1 - n)] = H[b + n * kNpfb];
    return banks;
}

static unsigned mseq_step(unsigned &v, unsigned g, unsigned mask)
{
    unsigned b = (unsigned)__builtin_parity(v & g);
    v = ((v << 1) | b) & mask;
    return b;
}

void preamble_pn(cf pn[64])
{
    // msequence_create(7, 0x0089, 1): taps 0x89 >> 1, start state = bit-reversed 1 over 7 bits
    unsigned v = 0x40, g = 0x44, mask = 0x7f;
    const float a = (float)M_SQRT1_2;
    for (unsigned i = 0; i < 64; ++i) {
        pn[i].re = mseq_step(v, g, mask) ? a : -a;
        pn[i].im = mseq_step(v, g, mask) ? a : -a;
    }
}

void header_pilots(cf p[15])
{
    // default m-sequence of order ceil(log2(15)) = 4: genpoly 0x13
    unsigned v = 0x8, g = 0x13 >> 1, mask = 0xf;
    for (unsigned i = 0; i < 15; ++i) {
        unsigned s = mseq_step(v, g, mask);
        s = (s << 1) | mseq_step(v, g, mask);
        float theta = (2.0f * kPi * (float)s / 4.0f) + kPi / 4.0f;
        p[i].re = std::cos(theta);
        p[i].im = std::sin(theta);
    }
}

std::vector<cf> detector_template(float beta)
{
    cf pn[64];
    preamble_pn(pn);
    std::vector<float> h = interp_taps(beta);
    const unsigned nsym = kPreamble + 2 * kM;
    std::vector<cf> s(kK * nsym);
    for (unsigned t = 0; t < nsym; ++t)
        for (unsigned ph = 0; ph < kK; ++ph) {
            float ar = 0.0f, ai = 0.0f;
            for (int n = 14; n >= 0; --n) {          // oldest contributing symbol first
                int u = (int)t - n;
                if (u < 0 || u >= (int)kPreamble) continue;
                ar = std::fmaf(h[ph + kK * n], pn[u].re, ar);
                ai = std::fmaf(h[ph + kK * n], pn[u].im, ai);
            }
            s[kK * t + ph] = { ar, ai };
        }
    return s;
}

std::vector<cf> twiddles(unsigned n)
{
    std::vector<cf> w(n / 2);
    for (unsigned k = 0; k < n / 2; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        w[k] = { (float)std::cos(a), (float)(-std::sin(a)) };
    }
    if (n >= 4) w[n / 4] = { 0.0f, -1.0f };
    return w;
}

void host_fft(const cf *in, cf *out, unsigned n, int dir)
{
    unsigned lg = 0;
    while ((1u << lg) < n) ++lg;
    std::vector<cf> W = twiddles(n), a(n);
    for (unsigned i = 0; i < n; ++i) {
        unsigned r = 0;
        for (unsigned b = 0; b < lg; ++b) r |= ((i >> b) & 1u) << (lg - 1 - b);
        a[i] = in[r];
    }
    for (unsigned half = 1; half < n; half <<= 1) {
        unsigned step = n / (2 * half);
        for (unsigned k = 0; k < n; k += 2 * half)
            for (unsigned j = 0; j < half; ++j) {
                float wr = W[j * step].re, wi = dir > 0 ? W[j * step].im : -W[j * step].im;
                cf lo = a[k + j], hi = a[k + j + half];
                float tr = std::fmaf(-wi, hi.im, wr * hi.re), ti = std::fmaf(wi, hi.re, wr * hi.im);
                a[k + j] = { lo.re + tr, lo.im + ti };
                a[k + j + half] = { lo.re - tr, lo.im - ti };
            }
    }
    std::memcpy(out, a.data(), n * sizeof(cf));
}

const float *nco_sintab()
{
    static float tab[1024];
    static std::once_flag once;
    std::call_once(once, [] {
        for (unsigned i = 0; i < 1024; ++i) tab[i] = std::sin(2.0f * kPi * (float)i / 1024.0f);
    });
    return tab;
}

uint32_t nco_constrain(float theta)
{
    float p = theta * 0.15915494309189535f;
    float f = p - (float)((long long)p);
    if (f < 0.0f) f += 1.0f;
    return (uint32_t)(long long)(f * 4294967296.0f);
}

// ------------------------------------------------------------------ modem
bool modem_supported(unsigned ms) { return modem_supported_hd(ms); }
unsigned modem_bps(unsigned ms) { return modem_bps_hd(ms); }

std::vector<cf> psk_maps()
{
    std::vector<cf> maps(8 * 256, cf{ 0.0f, 0.0f });
    for (unsigned b = 1; b <= 8; ++b) {
        unsigned M = 1u << b;
        float alpha = kPi / (float)M;
        for (unsigned s = 0; s < M; ++s) {
            unsigned g = s;
            for (unsigned sh = 1; sh < 32; sh <<= 1) g ^= g >> sh;   // gray decode
            float th = (float)g * 2.0f * alpha;
            maps[(b - 1) * 256 + s] = { std::cos(th), std::sin(th) };
        }
    }
    return maps;
}

// ------------------------------------------------------------------ lengths
unsigned crc_len(unsigned c) { return crc_len_hd(c); }
bool fec_supported(unsigned fs) { return fec_supported_hd(fs); }
unsigned fec_enc_len(unsigned fs, unsigned n) { return fec_enc_len_hd(fs, n); }
unsigned packetizer_enc_len(unsigned n, unsigned check, unsigned fec0, unsigned fec1) { return packetizer_enc_len_hd(n, check, fec0, fec1); }
unsigned qpm_frame_len(unsigned n, unsigned check, unsigned fec0, unsigned fec1, unsigned ms) { return qpm_frame_len_hd(n, check, fec0, fec1, ms); }

// ------------------------------------------------------------------ interleaver
void ilv_dims(unsigned n, unsigned &M, unsigned &N)
{
    unsigned r = 0;
    while ((r + 1) * (r + 1) <= n) ++r;
    if (n == 0) r = 0;
    M = 1 + r;
    N = n / M;
    while (n >= M * N) ++N;
}

std::vector<uint32_t> ilv_maps(unsigned n)
{
    unsigned M, N, n2 = n / 2;
    std::vector<uint32_t> maps(4 * (size_t)n2);
    if (n2 == 0) return maps;
    ilv_dims(n, M, N);
    const unsigned extra[4] = { 0, 2, 4, 8 };
    for (unsigned p = 0; p < 4; ++p) {
        unsigned cols = N + extra[p], row = 0, col = cols / 3, i = 0;
        while (i < n2) {                                  // walk the M x cols grid column by column
            unsigned j = row * cols + col;
            if (++row == M) { row = 0; col = (col + 1) % cols; }
            if (j < n2) maps[p * (size_t)n2 + i++] = j;
        }
    }
    return maps;
}

// The deinterleaver of an n-byte block as a permutation of its 8 n coded bits (MSB-first positions):
// deinterleaved[i] = interleaved[perm[i]].  The masked byte swaps only move bits, so the passes (3, 2, 1, 0) are run
// once on position labels.  Used by the soft-decision path, where every coded bit is a byte of its own.
std::vector<uint32_t> ilv_bit_perm(unsigned n)
{
    const unsigned n2 = n / 2, masks[4] = { 0xffu, 0x0fu, 0x55u, 0x33u };
    std::vector<uint32_t> maps = ilv_maps(n), lab(8 * (size_t)n), perm(8 * (size_t)n);
    for (size_t i = 0; i < lab.size(); ++i) lab[i] = (uint32_t)i;            // index 8 * byte + b, b counted from the LSB
    if (n >= 2)
        for (int pass = 3; pass >= 0; --pass)
            for (unsigned i = 0; i < n2; ++i) {
                const unsigned j = maps[(size_t)pass * n2 + i];
                for (unsigned b = 0; b < 8; ++b)
                    if ((masks[pass] >> b) & 1u) std::swap(lab[8 * (size_t)(2 * j + 1) + b], lab[8 * (size_t)(2 * i) + b]);
            }
    auto flip = [](uint32_t k) { return (k & ~7u) | (7u - (k & 7u)); };      // LSB-based index <-> MSB-first position
    for (size_t pos = 0; pos < perm.size(); ++pos) perm[pos] = flip(lab[flip((uint32_t)pos)]);
    return perm;
}

// ------------------------------------------------------------------ small code tables
static const uint8_t kH84[16] = { 0x00, 0xd2, 0x55, 0x87, 0x99, 0x4b, 0xcc, 0x1e,
                                  0xe1, 0x33, 0xb4, 0x66, 0x78, 0xaa, 0x2d, 0xff };

void hamming_dec_tables(uint8_t h84[256], uint8_t h74[128])
{
    for (unsigned r = 0; r < 256; ++r) {
        unsigned best = 0, bd = 99;
        for (unsigned s = 0; s < 16; ++s) {
            unsigned d = (unsigned)__builtin_popcount(kH84[s] ^ r);
            if (d < bd) { bd = d; best = s; }
        }
        h84[r] = (uint8_t)best;
    }
    for (unsigned r = 0; r < 128; ++r) {
        unsigned best = 0, bd = 99;
        for (unsigned s = 0; s < 16; ++s) {
            unsigned d = (unsigned)__builtin_popcount((kH84[s] >> 1) ^ r);
            if (d < bd) { bd = d; best = s; }
        }
        h74[r] = (uint8_t)best;
    }
}

// liquid-dsp's Hsiao SEC-DED codes (fec_secded2216 / 3932 / 7264): parity matrices [R x C], row 0 = the most significant
// parity bit; col[c][i] = contribution of data bit i (0 = MSB of the first data byte) to the parity byte of code c
// (0: (22,16), 1: (39,32), 2: (72,64)).  Every column is distinct and of odd weight.
static const uint8_t kSecded2216P[12] = { 0x99, 0x3c, 0x3e, 0x8a, 0xee, 0x60, 0xe1, 0xd1, 0x13, 0xc7, 0x44, 0x3f };
static const uint8_t kSecded3932P[28] = { 0x8a, 0x82, 0x0f, 0x1b, 0x10, 0x1f, 0x71, 0x61, 0x16, 0xf0, 0x92, 0xa6, 0xff, 0x01, 0xa4, 0x44,
                                          0x6c, 0xff, 0x08, 0x08, 0x21, 0x24, 0xff, 0x90, 0xc1, 0x48, 0x40, 0xff };
static const uint8_t kSecded7264P[64] = { 0xff, 0x0f, 0x0f, 0x0c, 0x68, 0x88, 0x88, 0x80, 0xf0, 0xff, 0x00, 0xf3, 0x64, 0x44, 0x44, 0x40,
                                          0x30, 0xf0, 0xff, 0x0f, 0x02, 0x22, 0x22, 0x26, 0xcf, 0x00, 0xf0, 0xff, 0x01, 0x11, 0x11, 0x16,
                                          0x68, 0x88, 0x88, 0x80, 0xff, 0x0f, 0x00, 0xf3, 0x64, 0x44, 0x44, 0x40, 0xf0, 0xff, 0x0f, 0x0c,
                                          0x02, 0x22, 0x22, 0x26, 0xcf, 0x00, 0xff, 0x0f, 0x01, 0x11, 0x11, 0x16, 0x30, 0xf0, 0xf0, 0xff };
void secded_cols(uint8_t col[3][64])
{
    const uint8_t *P[3] = { kSecded2216P, kSecded3932P, kSecded7264P };
    const unsigned nb[3] = { 2, 4, 8 }, R[3] = { 6, 7, 8 };
    for (unsigned c = 0; c < 3; ++c)
        for (unsigned i = 0; i < 64; ++i) {
            unsigned v = 0;
            if (i < 8 * nb[c])
                for (unsigned r = 0; r < R[c]; ++r) v |= ((P[c][r * nb[c] + (i >> 3)] >> (7 - (i & 7))) & 1u) << (R[c] - 1 - r);
            col[c][i] = (uint8_t)v;
        }
}

void gf256_tables(uint8_t gf_exp[512], uint8_t gf_log[256], uint8_t rs_gen[33])
{
    unsigned x = 1;
    for (unsigned i = 0; i < 255; ++i) {
        gf_exp[i] = (uint8_t)x;
        gf_log[x] = (uint8_t)i;
        x <<= 1;
        if (x & 0x100) x ^= 0x11d;
    }
    for (unsigned i = 255; i < 512; ++i) gf_exp[i] = gf_exp[i - 255];
    gf_log[0] = 255;
    std::memset(rs_gen, 0, 33);
    rs_gen[0] = 1;
    for (unsigned r = 1; r <= 32; ++r) {                  // times (x + alpha^r)
        for (unsigned j = r; j > 0; --j)
            rs_gen[j] = rs_gen[j - 1] ^ (rs_gen[j] ? gf_exp[gf_log[rs_gen[j]] + r] : 0);
        rs_gen[0] = gf_exp[gf_log[rs_gen[0]] + r];
    }
}

void crc_table(unsigned check, uint32_t tab[256])
{
    unsigned poly = 0, bits = 0;
    switch (check) {
    case CRC_8: poly = 0x07; bits = 8; break;
    case CRC_16: poly = 0x8005; bits = 16; break;
    case CRC_24: poly = 0x5D6DCB; bits = 24; break;
    case CRC_32: poly = 0x04C11DB7; bits = 32; break;
    default: std::memset(tab, 0, 256 * sizeof(uint32_t)); return;
    }
    uint32_t rp = 0;
    for (unsigned i = 0; i < bits; ++i) if (poly & (1u << i)) rp |= 1u << (bits - 1 - i);
    for (unsigned v = 0; v < 256; ++v) {
        uint32_t k = v;
        for (int j = 0; j < 8; ++j) k = (k >> 1) ^ (rp & (0u - (k & 1u)));
        tab[v] = k;
    }
}

}  // namespace lqb
