/*
 * lqo_fec.c -- ORACLE (test infrastructure only; see lqo.h header).
 * CRC, scrambler, interleaver, block codes, convolutional codes + Viterbi,
 * Reed-Solomon (255,223) and the packetizer.  Follows SURVEY.md Appendix A.7
 * (reference call sites: /root/reference/lib/flex_tx_impl.cc:52,119-181 select
 * the schemes; /root/reference/lib/flex_rx_impl.cc:75-136 name them on RX).
 */
#include "lqo.h"
#include <stdlib.h>
#include <string.h>

/* ================================================================== CRC */
unsigned lqo_crc_len(int scheme)
{
    switch (scheme) {
    case LQ_CRC_CHECKSUM: case LQ_CRC_8: return 1;
    case LQ_CRC_16: return 2;
    case LQ_CRC_24: return 3;
    case LQ_CRC_32: return 4;
    default: return 0;
    }
}

static unsigned reflect_(unsigned v, unsigned bits)
{
    unsigned r = 0;
    for (unsigned i = 0; i < bits; i++) if (v & (1u << i)) r |= 1u << (bits - 1 - i);
    return r;
}

/* reflected shift register; the register starts all-ones over 32 bits for every
 * width (the "key = ~0" quirk) and the result is complemented and masked */
static unsigned crc_generic_(const uint8_t *msg, unsigned n, unsigned poly, unsigned bits)
{
    unsigned key = ~0u, rpoly = reflect_(poly, bits);
    for (unsigned i = 0; i < n; i++) {
        key ^= msg[i];
        for (unsigned j = 0; j < 8; j++) {
            unsigned mask = -(key & 1u);
            key = (key >> 1) ^ (rpoly & mask);
        }
    }
    return (~key) & (bits == 32 ? 0xffffffffu : ((1u << bits) - 1u));
}

unsigned lqo_crc_key(int scheme, const uint8_t *msg, unsigned n)
{
    switch (scheme) {
    case LQ_CRC_CHECKSUM: {
        unsigned sum = 0;
        for (unsigned i = 0; i < n; i++) sum += msg[i];
        return (~sum + 1u) & 0xffu;
    }
    case LQ_CRC_8:  return crc_generic_(msg, n, 0x07, 8);
    case LQ_CRC_16: return crc_generic_(msg, n, 0x8005, 16);
    case LQ_CRC_24: return crc_generic_(msg, n, 0x5D6DCB, 24);
    case LQ_CRC_32: return crc_generic_(msg, n, 0x04C11DB7, 32);
    default: return 0;
    }
}

/* ================================================================== scrambler */
void lqo_scramble(uint8_t *x, unsigned n)
{
    static const uint8_t mask[4] = { 0xb4, 0x6a, 0x8b, 0xc5 };
    for (unsigned i = 0; i < n; i++) x[i] ^= mask[i & 3];
}

/* ================================================================== interleaver */
static void ilv_dims_(unsigned n, unsigned *M, unsigned *N)
{
    unsigned m = 1;
    while ((m + 1) * (m + 1) <= n) m++;          /* floor(sqrt(n)) */
    if (n == 0) m = 0;
    *M = 1 + m;
    *N = n / *M;
    while (n >= (*M) * (*N)) (*N)++;
}

/* one pass: bytes 2i (even) are exchanged -- fully or under a bit mask -- with
 * bytes 2j+1 (odd), j walking an M x N grid column by column from column N/3 */
static void ilv_pass_(uint8_t *x, unsigned n, unsigned M, unsigned N, uint8_t mask)
{
    unsigned n2 = n / 2, m = 0, col = N / 3, j;
    for (unsigned i = 0; i < n2; i++) {
        do {
            j = m * N + col;
            if (++m == M) { col = (col + 1) % N; m = 0; }
        } while (j >= n2);
        uint8_t a = x[2 * j + 1], b = x[2 * i];
        x[2 * j + 1] = (uint8_t)((a & ~mask) | (b & mask));
        x[2 * i]     = (uint8_t)((a & mask) | (b & ~mask));
    }
}

void lqo_interleave(uint8_t *x, unsigned n, int depth)
{
    unsigned M, N;
    if (n < 2 || depth <= 0) return;
    ilv_dims_(n, &M, &N);
    if (depth > 0) ilv_pass_(x, n, M, N, 0xff);
    if (depth > 1) ilv_pass_(x, n, M, N + 2, 0x0f);
    if (depth > 2) ilv_pass_(x, n, M, N + 4, 0x55);
    if (depth > 3) ilv_pass_(x, n, M, N + 8, 0x33);
}

void lqo_deinterleave(uint8_t *x, unsigned n, int depth)
{
    unsigned M, N;
    if (n < 2 || depth <= 0) return;
    ilv_dims_(n, &M, &N);
    if (depth > 3) ilv_pass_(x, n, M, N + 8, 0x33);
    if (depth > 2) ilv_pass_(x, n, M, N + 4, 0x55);
    if (depth > 1) ilv_pass_(x, n, M, N + 2, 0x0f);
    if (depth > 0) ilv_pass_(x, n, M, N, 0xff);
}

/* ================================================================== bit packing helpers */
static void pack_bits_(uint8_t *dst, unsigned k, unsigned b, unsigned sym)
{   /* write b bits of sym, MSB first, at bit index k */
    for (unsigned i = 0; i < b; i++) {
        unsigned bit = (sym >> (b - 1 - i)) & 1u, pos = k + i;
        if (bit) dst[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
        else     dst[pos >> 3] &= (uint8_t)~(0x80u >> (pos & 7));
    }
}
static unsigned unpack_bits_(const uint8_t *src, unsigned k, unsigned b)
{
    unsigned s = 0;
    for (unsigned i = 0; i < b; i++) { unsigned pos = k + i; s = (s << 1) | ((src[pos >> 3] >> (7 - (pos & 7))) & 1u); }
    return s;
}
static unsigned block_enc_len_(unsigned n, unsigned m, unsigned k)
{
    unsigned bits = n * 8, blocks = bits / m + (bits % m ? 1 : 0), out = blocks * k;
    return out / 8 + (out % 8 ? 1 : 0);
}

/* ================================================================== Hamming codes */
static const uint8_t h84_enc[16] = { 0x00, 0xd2, 0x55, 0x87, 0x99, 0x4b, 0xcc, 0x1e,
                                     0xe1, 0x33, 0xb4, 0x66, 0x78, 0xaa, 0x2d, 0xff };
static unsigned nearest_(unsigned r, unsigned shift)
{   /* nearest codeword of the (8,4) table (shift=0) or the (7,4) table (shift=1); ties -> lowest symbol */
    unsigned best = 0, bd = 99;
    for (unsigned s = 0; s < 16; s++) {
        unsigned d = (unsigned)__builtin_popcount((h84_enc[s] >> shift) ^ r);
        if (d < bd) { bd = d; best = s; }
    }
    return best;
}
static unsigned h128_encode_(unsigned d)
{
    unsigned p = ((unsigned)__builtin_parity(d & 0xda) << 3) | ((unsigned)__builtin_parity(d & 0xb6) << 2)
               | ((unsigned)__builtin_parity(d & 0x71) << 1) | (unsigned)__builtin_parity(d & 0x0f);
    return (p << 8) | d;
}
static unsigned h128_decode_(unsigned r)
{
    /* classic positions 1..12: parity at 1,2,4,8 = p bits 3,2,1,0; data d7..d0 at 3,5,6,7,9,10,11,12 */
    static const int pos2bit[13] = { -1, 11, 10, 7, 9, 6, 5, 4, 8, 3, 2, 1, 0 };
    unsigned d = r & 0xff, p = (r >> 8) & 0xf;
    unsigned z = (((unsigned)__builtin_parity(d & 0xda) ^ ((p >> 3) & 1)) << 0)
               | (((unsigned)__builtin_parity(d & 0xb6) ^ ((p >> 2) & 1)) << 1)
               | (((unsigned)__builtin_parity(d & 0x71) ^ ((p >> 1) & 1)) << 2)
               | (((unsigned)__builtin_parity(d & 0x0f) ^ ((p >> 0) & 1)) << 3);
    if (z >= 1 && z <= 12) r ^= 1u << pos2bit[z];
    return r & 0xff;
}

/* ================================================================== Golay (24,12) */
static const unsigned golay_P[12] = { 0x8ed, 0x1db, 0x3b5, 0x769, 0xed1, 0xda3,
                                      0xb47, 0x68f, 0xd1d, 0xa3b, 0x477, 0xffe };
static unsigned golay_mulP_(unsigned v)
{   /* v (12-bit row vector) * P; P symmetric */
    unsigned r = 0;
    for (unsigned i = 0; i < 12; i++) r |= (unsigned)__builtin_parity(v & golay_P[i]) << (11 - i);
    return r;
}
static unsigned golay_encode_(unsigned m) { return (golay_mulP_(m) << 12) | (m & 0xfff); } /* [parity | message] */
static unsigned golay_decode_(unsigned r)
{
    unsigned rp = (r >> 12) & 0xfff, rm = r & 0xfff;
    unsigned s = rp ^ golay_mulP_(rm);           /* syndrome w.r.t. H = [I | P] */
    unsigned ep = 0, em = 0; int found = 0;
    if (__builtin_popcount(s) <= 3) { ep = s; em = 0; found = 1; }
    for (unsigned i = 0; i < 12 && !found; i++)
        if (__builtin_popcount(s ^ golay_P[i]) <= 2) { ep = s ^ golay_P[i]; em = 1u << (11 - i); found = 1; }
    if (!found) {
        unsigned q = golay_mulP_(s);
        if (__builtin_popcount(q) <= 3) { ep = 0; em = q; found = 1; }
        for (unsigned i = 0; i < 12 && !found; i++)
            if (__builtin_popcount(q ^ golay_P[i]) <= 2) { ep = 1u << (11 - i); em = q ^ golay_P[i]; found = 1; }
    }
    (void)ep;
    return (rm ^ em) & 0xfff;                    /* uncorrectable -> message part as received */
}

/* ================================================================== SECDED (22,16) / (39,32) / (72,64) */
/* liquid-dsp's Hsiao codes (src/fec/src/fec_secded{2216,3932,7264}.c): parity matrices P [R x C], one row per parity
 * bit, row 0 the most significant bit of the parity byte, blocks stored [parity byte][data bytes].  Every column is
 * distinct and of odd weight (3 or 5), so a nonzero syndrome that equals a column is a single data-bit error (corrected),
 * a syndrome of weight one a parity-bit error, anything else at least two errors (left as received).  The matrices are
 * restated from liquid-dsp; tests/test_oracle_kats.py checks the structural properties that make them SEC-DED codes. */
static const uint8_t secded2216_P[12] = { 0x99, 0x3c, 0x3e, 0x8a, 0xee, 0x60, 0xe1, 0xd1, 0x13, 0xc7, 0x44, 0x3f };
static const uint8_t secded3932_P[28] = { 0x8a, 0x82, 0x0f, 0x1b, 0x10, 0x1f, 0x71, 0x61, 0x16, 0xf0, 0x92, 0xa6, 0xff, 0x01, 0xa4, 0x44,
                                          0x6c, 0xff, 0x08, 0x08, 0x21, 0x24, 0xff, 0x90, 0xc1, 0x48, 0x40, 0xff };
static const uint8_t secded7264_P[64] = { 0xff, 0x0f, 0x0f, 0x0c, 0x68, 0x88, 0x88, 0x80, 0xf0, 0xff, 0x00, 0xf3, 0x64, 0x44, 0x44, 0x40,
                                          0x30, 0xf0, 0xff, 0x0f, 0x02, 0x22, 0x22, 0x26, 0xcf, 0x00, 0xf0, 0xff, 0x01, 0x11, 0x11, 0x16,
                                          0x68, 0x88, 0x88, 0x80, 0xff, 0x0f, 0x00, 0xf3, 0x64, 0x44, 0x44, 0x40, 0xf0, 0xff, 0x0f, 0x0c,
                                          0x02, 0x22, 0x22, 0x26, 0xcf, 0x00, 0xff, 0x0f, 0x01, 0x11, 0x11, 0x16, 0x30, 0xf0, 0xf0, 0xff };
/* column of data bit i (0 = MSB of the first data byte) of the code with nb data bytes: its contribution to the parity byte */
static unsigned secded_col_(unsigned nb, unsigned i)
{
    const uint8_t *P = nb == 2 ? secded2216_P : nb == 4 ? secded3932_P : secded7264_P;
    const unsigned R = nb == 2 ? 6 : nb == 4 ? 7 : 8;
    unsigned c = 0;
    for (unsigned r = 0; r < R; r++) c |= ((P[r * nb + (i >> 3)] >> (7 - (i & 7))) & 1u) << (R - 1 - r);
    return c;
}
void lqo_secded_columns(unsigned nb, uint8_t *col /* 8 nb */) { for (unsigned i = 0; i < 8 * nb; i++) col[i] = (uint8_t)secded_col_(nb, i); }
static unsigned secded_parity_(const uint8_t *d, unsigned nb)
{
    unsigned p = 0;
    for (unsigned i = 0; i < nb * 8; i++)
        if ((d[i >> 3] >> (7 - (i & 7))) & 1u) p ^= secded_col_(nb, i);
    return p;
}
static void secded_encode_(unsigned nb, unsigned n, const uint8_t *dec, uint8_t *enc)
{
    unsigned i = 0, j = 0;
    uint8_t blk[8];
    while (i < n) {
        unsigned r = (n - i >= nb) ? nb : (n - i);
        memset(blk, 0, sizeof blk); memcpy(blk, dec + i, r);
        enc[j++] = (uint8_t)secded_parity_(blk, nb);
        memcpy(enc + j, dec + i, r); j += r; i += r;
    }
}
static void secded_decode_(unsigned nb, unsigned n, const uint8_t *enc, uint8_t *dec)
{
    const unsigned R = nb == 2 ? 6 : nb == 4 ? 7 : 8;
    unsigned i = 0, j = 0;
    uint8_t blk[8];
    while (i < n) {
        unsigned r = (n - i >= nb) ? nb : (n - i);
        unsigned rp = enc[j++] & ((1u << R) - 1u);          /* (the unused high bits of the parity byte do not enter) */
        memset(blk, 0, sizeof blk); memcpy(blk, enc + j, r); j += r;
        unsigned syn = secded_parity_(blk, nb) ^ rp;
        if (syn)                                             /* a column: that data bit; else a parity bit or >= 2 errors */
            for (unsigned b = 0; b < nb * 8; b++)
                if (secded_col_(nb, b) == syn) { blk[b >> 3] ^= (uint8_t)(0x80u >> (b & 7)); break; }
        memcpy(dec + i, blk, r); i += r;
    }
}

/* ================================================================== convolutional codes */
typedef struct { unsigned K, P; unsigned poly[2]; const uint8_t *pm; } conv_t;
static const uint8_t pm_none[2]   = { 1, 1 };
static const uint8_t pm27_23[4]   = { 1,1, 1,0 };
static const uint8_t pm27_34[6]   = { 1,1,0, 1,0,1 };
static const uint8_t pm27_45[8]   = { 1,1,1,1, 1,0,0,0 };
static const uint8_t pm27_56[10]  = { 1,1,0,1,0, 1,0,1,0,1 };
static const uint8_t pm27_67[12]  = { 1,1,1,0,1,0, 1,0,0,1,0,1 };
static const uint8_t pm27_78[14]  = { 1,1,1,1,0,1,0, 1,0,0,0,1,0,1 };
static const uint8_t pm29_23[4]   = { 1,1, 1,0 };
static const uint8_t pm29_34[6]   = { 1,1,1, 1,0,0 };
static const uint8_t pm29_45[8]   = { 1,0,1,1, 1,1,0,0 };
static const uint8_t pm29_56[10]  = { 1,1,0,1,0, 1,0,1,0,1 };
static const uint8_t pm29_67[12]  = { 1,1,0,1,1,0, 1,0,1,0,0,1 };
static const uint8_t pm29_78[14]  = { 1,1,0,1,0,1,1, 1,0,1,0,1,0,0 };

static int conv_lookup_(int fs, conv_t *c)
{
    c->poly[0] = 0x6d; c->poly[1] = 0x4f; c->K = 7; c->P = 1; c->pm = pm_none;
    switch (fs) {
    case LQ_FEC_CONV_V27: return 1;
    case LQ_FEC_CONV_V27P23: c->P = 2; c->pm = pm27_23; return 1;
    case LQ_FEC_CONV_V27P34: c->P = 3; c->pm = pm27_34; return 1;
    case LQ_FEC_CONV_V27P45: c->P = 4; c->pm = pm27_45; return 1;
    case LQ_FEC_CONV_V27P56: c->P = 5; c->pm = pm27_56; return 1;
    case LQ_FEC_CONV_V27P67: c->P = 6; c->pm = pm27_67; return 1;
    case LQ_FEC_CONV_V27P78: c->P = 7; c->pm = pm27_78; return 1;
    default: break;
    }
    c->poly[0] = 0x1af; c->poly[1] = 0x11d; c->K = 9;
    switch (fs) {
    case LQ_FEC_CONV_V29: return 1;
    case LQ_FEC_CONV_V29P23: c->P = 2; c->pm = pm29_23; return 1;
    case LQ_FEC_CONV_V29P34: c->P = 3; c->pm = pm29_34; return 1;
    case LQ_FEC_CONV_V29P45: c->P = 4; c->pm = pm29_45; return 1;
    case LQ_FEC_CONV_V29P56: c->P = 5; c->pm = pm29_56; return 1;
    case LQ_FEC_CONV_V29P67: c->P = 6; c->pm = pm29_67; return 1;
    case LQ_FEC_CONV_V29P78: c->P = 7; c->pm = pm29_78; return 1;
    default: return 0;
    }
}

static unsigned conv_enc_len_(const conv_t *c, unsigned n)
{
    if (c->P == 1) return 2 * n + 2;
    unsigned nb = n * 8 + c->K - 1;
    unsigned out = nb + (nb + c->P - 1) / c->P;
    return out / 8 + (out % 8 ? 1 : 0);
}

static void conv_encode_(const conv_t *c, unsigned n, const uint8_t *dec, uint8_t *enc)
{
    unsigned sr = 0, nout = 0, p = 0, total = n * 8 + c->K - 1;
    uint8_t byte_out = 0;
    for (unsigned t = 0; t < total; t++) {
        unsigned bit = (t < n * 8) ? ((dec[t >> 3] >> (7 - (t & 7))) & 1u) : 0u;
        sr = (sr << 1) | bit;
        for (unsigned r = 0; r < 2; r++) {
            if (c->pm[r * c->P + p]) {
                byte_out = (uint8_t)((byte_out << 1) | (unsigned)__builtin_parity(sr & c->poly[r]));
                enc[nout >> 3] = byte_out;
                nout++;
            }
        }
        p = (p + 1) % c->P;
    }
    while (nout & 7) { byte_out <<= 1; enc[nout >> 3] = byte_out; nout++; }
}

/* hard-input Viterbi: bits expanded to 0/255, punctured positions erased to 127; 32-bit
 * metrics, start state 0 biased (others 63), traceback from state 0 */
static void conv_decode_x_(const conv_t *c, unsigned n, const uint8_t *enc, const uint8_t *soft, uint8_t *dec);
static void conv_decode_(const conv_t *c, unsigned n, const uint8_t *enc, uint8_t *dec) { conv_decode_x_(c, n, enc, NULL, dec); }

/* soft == NULL: hard input (bits of enc expanded to 0 / 255); else one soft byte per kept coded bit */
static void conv_decode_x_(const conv_t *c, unsigned n, const uint8_t *enc, const uint8_t *soft, uint8_t *dec)
{
    unsigned ns = 1u << (c->K - 1), half = ns >> 1, T = n * 8 + c->K - 1;
    unsigned words = ns / 32;
    uint32_t *dcs = (uint32_t *)calloc((size_t)T * words, sizeof(uint32_t));
    uint32_t *m_old = (uint32_t *)malloc(ns * sizeof(uint32_t));
    uint32_t *m_new = (uint32_t *)malloc(ns * sizeof(uint32_t));
    uint8_t *bt0 = (uint8_t *)malloc(half), *bt1 = (uint8_t *)malloc(half);
    for (unsigned i = 0; i < half; i++) {
        bt0[i] = __builtin_parity((2 * i) & c->poly[0]) ? 255 : 0;
        bt1[i] = __builtin_parity((2 * i) & c->poly[1]) ? 255 : 0;
    }
    for (unsigned i = 0; i < ns; i++) m_old[i] = 63;
    m_old[0] = 0;
    unsigned ib = 0, p = 0;
    for (unsigned t = 0; t < T; t++) {
        unsigned sym[2];
        for (unsigned r = 0; r < 2; r++) {
            if (c->pm[r * c->P + p]) { sym[r] = soft ? soft[ib] : (((enc[ib >> 3] >> (7 - (ib & 7))) & 1u) ? 255u : 0u); ib++; }
            else sym[r] = 127u;
        }
        p = (p + 1) % c->P;
        uint32_t *d = dcs + (size_t)t * words;
        for (unsigned i = 0; i < half; i++) {
            uint32_t metric = (bt0[i] ^ sym[0]) + (bt1[i] ^ sym[1]);
            uint32_t m0 = m_old[i] + metric, m1 = m_old[i + half] + (510u - metric);
            unsigned dcn = (int32_t)(m0 - m1) > 0;
            m_new[2 * i] = dcn ? m1 : m0;
            d[(2 * i) >> 5] |= (uint32_t)dcn << ((2 * i) & 31);
            m0 = m_old[i] + (510u - metric); m1 = m_old[i + half] + metric;
            dcn = (int32_t)(m0 - m1) > 0;
            m_new[2 * i + 1] = dcn ? m1 : m0;
            d[(2 * i + 1) >> 5] |= (uint32_t)dcn << ((2 * i + 1) & 31);
        }
        uint32_t *tmp = m_old; m_old = m_new; m_new = tmp;
    }
    memset(dec, 0, n);
    unsigned state = 0;
    for (unsigned t = T; t-- > 0;) {
        const uint32_t *d = dcs + (size_t)t * words;
        unsigned k = (d[state >> 5] >> (state & 31)) & 1u;
        if (t >= c->K - 1) {                     /* the bit shifted out at step t entered at t-(K-1) */
            unsigned bi = t - (c->K - 1);
            if (k) dec[bi >> 3] |= (uint8_t)(0x80u >> (bi & 7));
        }
        state = (state >> 1) | (k << (c->K - 2));
    }
    free(dcs); free(m_old); free(m_new); free(bt0); free(bt1);
}

/* ================================================================== Reed-Solomon (255,223), GF(256)/0x11d, roots a^1..a^32 */
static uint8_t gf_exp[512], gf_log[256], rs_gen[33];
static int rs_ready;
static void rs_init_(void)
{
    if (__atomic_load_n(&rs_ready, __ATOMIC_ACQUIRE)) return;
    unsigned x = 1;
    for (unsigned i = 0; i < 255; i++) {
        gf_exp[i] = (uint8_t)x; gf_log[x] = (uint8_t)i;
        x <<= 1; if (x & 0x100) x ^= 0x11d;
    }
    for (unsigned i = 255; i < 512; i++) gf_exp[i] = gf_exp[i - 255];
    gf_log[0] = 255;
    uint8_t g[33]; memset(g, 0, sizeof g); g[0] = 1;      /* g[i] = coeff of x^i */
    for (unsigned r = 1; r <= 32; r++) {                   /* multiply by (x + a^r) */
        for (unsigned j = r; j > 0; j--)
            g[j] = g[j - 1] ^ (g[j] ? gf_exp[gf_log[g[j]] + r] : 0);
        g[0] = gf_exp[gf_log[g[0]] + r];
    }
    memcpy(rs_gen, g, 33);
    __atomic_store_n(&rs_ready, 1, __ATOMIC_RELEASE);
}
static inline uint8_t gf_mul_(uint8_t a, uint8_t b) { return (a && b) ? gf_exp[gf_log[a] + gf_log[b]] : 0; }

void lqo_rs_encode_block(const uint8_t *data, unsigned pad, uint8_t *par)
{
    rs_init_();
    memset(par, 0, 32);
    for (unsigned i = 0; i < 223 - pad; i++) {
        uint8_t fb = data[i] ^ par[0];
        for (unsigned j = 0; j < 31; j++) par[j] = par[j + 1] ^ gf_mul_(fb, rs_gen[31 - j]);
        par[31] = gf_mul_(fb, rs_gen[0]);
    }
}

int lqo_rs_decode_block(uint8_t *data, unsigned pad)
{
    rs_init_();
    unsigned n = 255 - pad;
    uint8_t s[32];
    unsigned any = 0;
    for (unsigned i = 0; i < 32; i++) s[i] = data[0];
    for (unsigned j = 1; j < n; j++)
        for (unsigned i = 0; i < 32; i++)
            s[i] = data[j] ^ (s[i] ? gf_exp[gf_log[s[i]] + i + 1] : 0);
    for (unsigned i = 0; i < 32; i++) any |= s[i];
    if (!any) return 0;

    uint8_t lambda[33], b[33], t[33];
    memset(lambda, 0, 33); memset(b, 0, 33); lambda[0] = 1; b[0] = 1;
    unsigned el = 0;
    for (unsigned r = 1; r <= 32; r++) {
        uint8_t d = 0;
        for (unsigned i = 0; i < r; i++) d ^= gf_mul_(lambda[i], s[r - i - 1]);
        if (d == 0) {
            memmove(b + 1, b, 32); b[0] = 0;
        } else {
            t[0] = lambda[0];
            for (unsigned i = 0; i < 32; i++) t[i + 1] = lambda[i + 1] ^ gf_mul_(d, b[i]);
            if (2 * el <= r - 1) {
                el = r - el;
                uint8_t dinv = gf_exp[255 - gf_log[d]];
                for (unsigned i = 0; i <= 32; i++) b[i] = gf_mul_(lambda[i], dinv);
            } else {
                memmove(b + 1, b, 32); b[0] = 0;
            }
            memcpy(lambda, t, 33);
        }
    }
    unsigned deg = 0;
    for (unsigned i = 0; i <= 32; i++) if (lambda[i]) deg = i;

    unsigned root[32], loc[32], count = 0;
    for (unsigned i = 1; i <= 255; i++) {                  /* Chien: lambda(a^i) == 0 ? */
        uint8_t q = 1;
        for (unsigned j = 1; j <= deg; j++)
            if (lambda[j]) q ^= gf_exp[(gf_log[lambda[j]] + i * j) % 255];
        if (q) continue;
        root[count] = i; loc[count] = i - 1;
        if (++count == deg) break;
    }
    if (count != deg) return -1;

    uint8_t omega[32];
    for (unsigned i = 0; i < deg; i++) {
        uint8_t acc = 0;
        for (unsigned j = 0; j <= i; j++) acc ^= gf_mul_(s[i - j], lambda[j]);
        omega[i] = acc;
    }
    for (unsigned jj = count; jj-- > 0;) {
        uint8_t num = 0, den = 0;
        for (unsigned i = 0; i < deg; i++)
            if (omega[i]) num ^= gf_exp[(gf_log[omega[i]] + i * root[jj]) % 255];
        unsigned top = (deg < 31 ? deg : 31) & ~1u;
        for (unsigned i = 0; i <= top; i += 2)
            if (lambda[i + 1]) den ^= gf_exp[(gf_log[lambda[i + 1]] + i * root[jj]) % 255];
        if (num != 0 && loc[jj] >= pad)
            data[loc[jj] - pad] ^= gf_exp[(gf_log[num] + 255 - gf_log[den]) % 255];
    }
    return (int)count;
}

typedef struct { unsigned blocks, dec_block, enc_block, pad; } rs_plan_t;
static rs_plan_t rs_plan_(unsigned n)
{
    rs_plan_t p;
    p.blocks = n / 223 + (n % 223 ? 1 : 0);
    if (p.blocks == 0) p.blocks = 1;
    p.dec_block = n / p.blocks + (n % p.blocks ? 1 : 0);
    p.enc_block = p.dec_block + 32;
    p.pad = 223 - p.dec_block;
    return p;
}
static void rs_encode_(unsigned n, const uint8_t *dec, uint8_t *enc)
{
    rs_plan_t p = rs_plan_(n);
    uint8_t blk[255];
    unsigned n0 = 0;
    for (unsigned i = 0; i < p.blocks; i++) {
        unsigned take = (n - n0 >= p.dec_block) ? p.dec_block : (n - n0);
        memset(blk, 0, p.dec_block); memcpy(blk, dec + n0, take);
        lqo_rs_encode_block(blk, p.pad, blk + p.dec_block);
        memcpy(enc + i * p.enc_block, blk, p.enc_block);
        n0 += take;
    }
}
static void rs_decode_(unsigned n, const uint8_t *enc, uint8_t *dec)
{
    rs_plan_t p = rs_plan_(n);
    uint8_t blk[255];
    unsigned n0 = 0;
    for (unsigned i = 0; i < p.blocks; i++) {
        unsigned take = (n - n0 >= p.dec_block) ? p.dec_block : (n - n0);
        memcpy(blk, enc + i * p.enc_block, p.enc_block);
        lqo_rs_decode_block(blk, p.pad);
        memcpy(dec + n0, blk, take);
        n0 += take;
    }
}

/* ================================================================== FEC dispatch */
unsigned lqo_fec_enc_len(int fs, unsigned n)
{
    conv_t c;
    switch (fs) {
    case LQ_FEC_NONE: return n;
    case LQ_FEC_REP3: return 3 * n;
    case LQ_FEC_REP5: return 5 * n;
    case LQ_FEC_HAMMING74: return block_enc_len_(n, 4, 7);
    case LQ_FEC_HAMMING84: return 2 * n;
    case LQ_FEC_HAMMING128: return block_enc_len_(n, 8, 12);
    case LQ_FEC_GOLAY2412: return block_enc_len_(n, 12, 24);
    case LQ_FEC_SECDED2216: return n + n / 2 + (n % 2 ? 1 : 0);
    case LQ_FEC_SECDED3932: return n + n / 4 + (n % 4 ? 1 : 0);
    case LQ_FEC_SECDED7264: return n + n / 8 + (n % 8 ? 1 : 0);
    case LQ_FEC_RS_M8: { rs_plan_t p = rs_plan_(n); return n ? p.blocks * p.enc_block : 0; }
    default:
        if (conv_lookup_(fs, &c)) return conv_enc_len_(&c, n);
        return 0;
    }
}

void lqo_fec_encode(int fs, unsigned n, const uint8_t *dec, uint8_t *enc)
{
    conv_t c;
    unsigned k = 0;
    switch (fs) {
    case LQ_FEC_NONE: memcpy(enc, dec, n); return;
    case LQ_FEC_REP3: for (unsigned r = 0; r < 3; r++) memcpy(enc + r * n, dec, n); return;
    case LQ_FEC_REP5: for (unsigned r = 0; r < 5; r++) memcpy(enc + r * n, dec, n); return;
    case LQ_FEC_HAMMING74:
        memset(enc, 0, lqo_fec_enc_len(fs, n));
        for (unsigned i = 0; i < n; i++) {
            pack_bits_(enc, k, 7, h84_enc[dec[i] >> 4] >> 1); k += 7;
            pack_bits_(enc, k, 7, h84_enc[dec[i] & 15] >> 1); k += 7;
        }
        return;
    case LQ_FEC_HAMMING84:
        for (unsigned i = 0; i < n; i++) { enc[2 * i] = h84_enc[dec[i] >> 4]; enc[2 * i + 1] = h84_enc[dec[i] & 15]; }
        return;
    case LQ_FEC_HAMMING128:
        memset(enc, 0, lqo_fec_enc_len(fs, n));
        for (unsigned i = 0; i < n; i++) { pack_bits_(enc, k, 12, h128_encode_(dec[i])); k += 12; }
        return;
    case LQ_FEC_GOLAY2412: {
        unsigned i = 0, j = 0, r = n % 3;
        for (; i + 3 <= n; i += 3, j += 6) {
            unsigned s0 = ((unsigned)dec[i] << 4) | (dec[i + 1] >> 4);
            unsigned s1 = (((unsigned)dec[i + 1] & 15) << 8) | dec[i + 2];
            unsigned v0 = golay_encode_(s0), v1 = golay_encode_(s1);
            enc[j] = (uint8_t)(v0 >> 16); enc[j + 1] = (uint8_t)(v0 >> 8); enc[j + 2] = (uint8_t)v0;
            enc[j + 3] = (uint8_t)(v1 >> 16); enc[j + 4] = (uint8_t)(v1 >> 8); enc[j + 5] = (uint8_t)v1;
        }
        for (unsigned q = 0; q < r; q++, i++, j += 3) {      /* leftover bytes: one codeword each */
            unsigned v = golay_encode_(dec[i]);
            enc[j] = (uint8_t)(v >> 16); enc[j + 1] = (uint8_t)(v >> 8); enc[j + 2] = (uint8_t)v;
        }
        return;
    }
    case LQ_FEC_SECDED2216: secded_encode_(2, n, dec, enc); return;
    case LQ_FEC_SECDED3932: secded_encode_(4, n, dec, enc); return;
    case LQ_FEC_SECDED7264: secded_encode_(8, n, dec, enc); return;
    case LQ_FEC_RS_M8: rs_encode_(n, dec, enc); return;
    default:
        if (conv_lookup_(fs, &c)) conv_encode_(&c, n, dec, enc);
        return;
    }
}

void lqo_fec_decode(int fs, unsigned n, const uint8_t *enc, uint8_t *dec)
{
    conv_t c;
    unsigned k = 0;
    switch (fs) {
    case LQ_FEC_NONE: memcpy(dec, enc, n); return;
    case LQ_FEC_REP3:
        for (unsigned i = 0; i < n; i++) {
            uint8_t a = enc[i], b = enc[i + n], d = enc[i + 2 * n];
            dec[i] = (uint8_t)((a & b) | (a & d) | (b & d));
        }
        return;
    case LQ_FEC_REP5:
        for (unsigned i = 0; i < n; i++) {
            uint8_t out = 0;
            for (unsigned bit = 0; bit < 8; bit++) {
                unsigned cnt = 0;
                for (unsigned r = 0; r < 5; r++) cnt += (enc[i + r * n] >> bit) & 1u;
                if (cnt >= 3) out |= (uint8_t)(1u << bit);
            }
            dec[i] = out;
        }
        return;
    case LQ_FEC_HAMMING74:
        for (unsigned i = 0; i < n; i++) {
            unsigned r0 = unpack_bits_(enc, k, 7); k += 7;
            unsigned r1 = unpack_bits_(enc, k, 7); k += 7;
            dec[i] = (uint8_t)((nearest_(r0, 1) << 4) | nearest_(r1, 1));
        }
        return;
    case LQ_FEC_HAMMING84:
        for (unsigned i = 0; i < n; i++) dec[i] = (uint8_t)((nearest_(enc[2 * i], 0) << 4) | nearest_(enc[2 * i + 1], 0));
        return;
    case LQ_FEC_HAMMING128:
        for (unsigned i = 0; i < n; i++) { dec[i] = (uint8_t)h128_decode_(unpack_bits_(enc, k, 12)); k += 12; }
        return;
    case LQ_FEC_GOLAY2412: {
        unsigned i = 0, j = 0, r = n % 3;
        for (; i + 3 <= n; i += 3, j += 6) {
            unsigned v0 = ((unsigned)enc[j] << 16) | ((unsigned)enc[j + 1] << 8) | enc[j + 2];
            unsigned v1 = ((unsigned)enc[j + 3] << 16) | ((unsigned)enc[j + 4] << 8) | enc[j + 5];
            unsigned s0 = golay_decode_(v0), s1 = golay_decode_(v1);
            dec[i] = (uint8_t)(s0 >> 4); dec[i + 1] = (uint8_t)(((s0 & 15) << 4) | (s1 >> 8)); dec[i + 2] = (uint8_t)s1;
        }
        for (unsigned q = 0; q < r; q++, i++, j += 3) {
            unsigned v = ((unsigned)enc[j] << 16) | ((unsigned)enc[j + 1] << 8) | enc[j + 2];
            dec[i] = (uint8_t)golay_decode_(v);
        }
        return;
    }
    case LQ_FEC_SECDED2216: secded_decode_(2, n, enc, dec); return;
    case LQ_FEC_SECDED3932: secded_decode_(4, n, enc, dec); return;
    case LQ_FEC_SECDED7264: secded_decode_(8, n, enc, dec); return;
    case LQ_FEC_RS_M8: rs_decode_(n, enc, dec); return;
    default:
        if (conv_lookup_(fs, &c)) conv_decode_(&c, n, enc, dec);
        return;
    }
}

int lqo_fec_is_conv(int fs)
{
    conv_t c;
    return conv_lookup_(fs, &c) ? 1 : 0;
}

int lqo_fec_decode_soft(int fs, unsigned n, const uint8_t *soft, uint8_t *dec)
{
    conv_t c;
    if (!conv_lookup_(fs, &c)) return 0;
    conv_decode_x_(&c, n, NULL, soft, dec);
    return 1;
}

void lqo_deinterleave_bit_perm(unsigned n, uint32_t *perm)
{
    /* run the byte deinterleaver on the bit planes of the position labels */
    uint8_t *buf = (uint8_t *)malloc(n + 8);
    for (unsigned i = 0; i < 8 * n; i++) perm[i] = 0;
    unsigned planes = 1;
    while ((1u << planes) < 8 * n) planes++;
    for (unsigned p = 0; p < planes; p++) {
        for (unsigned i = 0; i < n; i++) {
            unsigned v = 0;
            for (unsigned q = 0; q < 8; q++) if (((8 * i + q) >> p) & 1u) v |= 0x80u >> q;
            buf[i] = (uint8_t)v;
        }
        lqo_deinterleave(buf, n, 4);
        for (unsigned i = 0; i < 8 * n; i++) if ((buf[i >> 3] >> (7 - (i & 7))) & 1u) perm[i] |= 1u << p;
    }
    free(buf);
}

/* ================================================================== packetizer */
unsigned lqo_packetizer_enc_len(unsigned n, int check, int fec0, int fec1)
{
    unsigned k = n + lqo_crc_len(check);
    return lqo_fec_enc_len(fec1, lqo_fec_enc_len(fec0, k));
}

void lqo_packetizer_encode(unsigned n, int check, int fec0, int fec1, const uint8_t *msg, uint8_t *pkt)
{
    unsigned cl = lqo_crc_len(check), k0 = n + cl;
    unsigned n0 = lqo_fec_enc_len(fec0, k0), n1 = lqo_fec_enc_len(fec1, n0);
    unsigned cap = (n1 > n0 ? n1 : n0) + k0 + 8;
    uint8_t *b0 = (uint8_t *)calloc(cap, 1), *b1 = (uint8_t *)calloc(cap, 1);
    memcpy(b0, msg, n);
    unsigned key = lqo_crc_key(check, b0, n);
    for (unsigned i = 0; i < cl; i++) { b0[n + cl - i - 1] = (uint8_t)(key & 0xff); key >>= 8; }
    lqo_scramble(b0, k0);
    lqo_fec_encode(fec0, k0, b0, b1);
    lqo_interleave(b1, n0, fec0 == LQ_FEC_NONE ? 0 : 4);
    lqo_fec_encode(fec1, n0, b1, b0);
    lqo_interleave(b0, n1, fec1 == LQ_FEC_NONE ? 0 : 4);
    memcpy(pkt, b0, n1);
    free(b0); free(b1);
}

int lqo_packetizer_decode(unsigned n, int check, int fec0, int fec1, const uint8_t *pkt, uint8_t *msg)
{
    unsigned cl = lqo_crc_len(check), k0 = n + cl;
    unsigned n0 = lqo_fec_enc_len(fec0, k0), n1 = lqo_fec_enc_len(fec1, n0);
    unsigned cap = (n1 > n0 ? n1 : n0) + k0 + 8;
    uint8_t *b0 = (uint8_t *)calloc(cap, 1), *b1 = (uint8_t *)calloc(cap, 1);
    memcpy(b0, pkt, n1);
    lqo_deinterleave(b0, n1, fec1 == LQ_FEC_NONE ? 0 : 4);
    lqo_fec_decode(fec1, n0, b0, b1);
    lqo_deinterleave(b1, n0, fec0 == LQ_FEC_NONE ? 0 : 4);
    lqo_fec_decode(fec0, k0, b1, b0);
    lqo_scramble(b0, k0);
    unsigned key = 0;
    for (unsigned i = 0; i < cl; i++) key = (key << 8) | b0[n + i];
    memcpy(msg, b0, n);
    int ok = (lqo_crc_key(check, b0, n) == key);
    free(b0); free(b1);
    return ok;
}
