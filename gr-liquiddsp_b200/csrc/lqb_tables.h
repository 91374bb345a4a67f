// lqb_tables.h -- host-side constant tables and length arithmetic for the B200 packet PHY.
//
// Everything here runs once at handle creation (or once per new payload size) on the host
// and is uploaded to the GPU; the per-sample work lives in the .cu files.  The formulas
// restate liquid-dsp >= 1.3.1; the reference selects the scheme
// numbers at lib/flex_tx_impl.cc:77-181 and decodes them at lib/flex_rx_impl.cc:75-179.
#pragma once
#include <cstdint>
#include <vector>

namespace lqb {

struct cf { float re, im; };

// frame constants (flexframegen / flexframesync)
constexpr unsigned kK = 2, kM = 7, kNpfb = 32, kPreamble = 64, kTaps = 28;
constexpr float kTxBeta = 0.25f, kRxBeta = 0.30f;
constexpr unsigned kHdrUser = 14, kHdrDec = 20, kHdrEnc = 54, kHdrMod = 216, kHdrSym = 231;
constexpr unsigned kPilotSpacing = 16, kNumPilots = 15, kProtocol = 102;
constexpr unsigned kNfft = 512, kSLen = 156;
constexpr float kPllBw = 1e-4f;

enum : unsigned {
    MODEM_PSK2 = 1, MODEM_PSK256 = 8, MODEM_DPSK2 = 9, MODEM_DPSK256 = 16, MODEM_ASK2 = 17,
    MODEM_ASK256 = 24, MODEM_QAM4 = 25, MODEM_QAM256 = 31, MODEM_BPSK = 39, MODEM_QPSK = 40,
    MODEM_NUM = 52
};
enum : unsigned {
    FEC_UNKNOWN = 0, FEC_NONE = 1, FEC_REP3, FEC_REP5, FEC_HAMMING74, FEC_HAMMING84, FEC_HAMMING128,
    FEC_GOLAY2412, FEC_SECDED2216, FEC_SECDED3932, FEC_SECDED7264, FEC_CONV_V27, FEC_CONV_V29,
    FEC_CONV_V39, FEC_CONV_V615, FEC_CONV_V27P23, FEC_CONV_V27P34, FEC_CONV_V27P45, FEC_CONV_V27P56,
    FEC_CONV_V27P67, FEC_CONV_V27P78, FEC_CONV_V29P23, FEC_CONV_V29P34, FEC_CONV_V29P45,
    FEC_CONV_V29P56, FEC_CONV_V29P67, FEC_CONV_V29P78, FEC_RS_M8, FEC_NUM
};
enum : unsigned { CRC_UNKNOWN = 0, CRC_NONE, CRC_CHECKSUM, CRC_8, CRC_16, CRC_24, CRC_32, CRC_NUM };

// ---- filter design
std::vector<float> firdes_arkaiser(unsigned k, unsigned m, float beta, float dt);
std::vector<float> interp_taps(float beta);              // 30 taps (29 + zero pad)
std::vector<float> pfb_banks(float beta);                // [32][28], each bank oldest -> newest
void preamble_pn(cf pn[64]);
std::vector<cf> detector_template(float beta);           // 156 samples
std::vector<cf> twiddles(unsigned n);                    // n/2 entries exp(-j 2 pi k / n)
void host_fft(const cf *in, cf *out, unsigned n, int dir);
const float *nco_sintab();                               // 1024
uint32_t nco_constrain(float theta);
void header_pilots(cf p[15]);

// ---- modem
bool modem_supported(unsigned ms);
unsigned modem_bps(unsigned ms);
std::vector<cf> psk_maps();                              // [8][256] PSK-2^b symbol maps, b = 1..8 at row b-1

// ---- length arithmetic
unsigned crc_len(unsigned check);
bool fec_supported(unsigned fs);
unsigned fec_enc_len(unsigned fs, unsigned n);
unsigned packetizer_enc_len(unsigned n, unsigned check, unsigned fec0, unsigned fec1);
unsigned qpm_frame_len(unsigned n, unsigned check, unsigned fec0, unsigned fec1, unsigned ms);

// ---- interleaver: for pass p (0..3) entry i is the j that byte 2i exchanges bits with byte 2j+1
void ilv_dims(unsigned n, unsigned &M, unsigned &N);
std::vector<uint32_t> ilv_maps(unsigned n);              // 4 * (n/2) entries
std::vector<uint32_t> ilv_bit_perm(unsigned n);          // 8 n entries: deinterleaved bit i = interleaved bit perm[i]

// ---- small code tables
void hamming_dec_tables(uint8_t h84[256], uint8_t h74[128]);
void secded_cols(uint8_t col[3][64]);   // [code 0: (22,16), 1: (39,32), 2: (72,64)][data bit]
void gf256_tables(uint8_t gf_exp[512], uint8_t gf_log[256], uint8_t rs_gen[33]);
void crc_table(unsigned check, uint32_t tab[256]);

}  // namespace lqb
