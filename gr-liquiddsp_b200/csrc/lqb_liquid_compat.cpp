// lqb_liquid_compat.cpp -- liquid-dsp-signature veneer (include/lqb200_liquid.h) over the batch
// C-ABI, one stream per object, host buffers.  See the header for the (timing-only) differences.
#include "../../include/lqb200.h"
#include "../../include/lqb200_liquid.h"
#include "lqb_tables.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

extern "C" void lqb_internal_set_error(const char *msg);

// ------------------------------------------------------------------ msequence
struct msequence_s { unsigned m, g, a, n, v, b; };

extern "C" msequence msequence_create(unsigned int m, unsigned int g, unsigned int a)
{
    msequence ms = new msequence_s;
    ms->m = m; ms->g = g >> 1; ms->a = 0;
    for (unsigned i = 0; i < m; ++i) { ms->a = (ms->a << 1) | (a & 1u); a >>= 1; }
    ms->n = (1u << m) - 1u; ms->v = ms->a; ms->b = 0;
    return ms;
}
extern "C" void msequence_destroy(msequence ms) { delete ms; }
extern "C" unsigned int msequence_advance(msequence ms)
{
    ms->b = (unsigned)__builtin_parity(ms->v & ms->g);
    ms->v = ((ms->v << 1) | ms->b) & ms->n;
    return ms->b;
}

// ------------------------------------------------------------------ flexframesync
namespace {
struct QueuedFrame {
    std::vector<unsigned char> header, payload;
    std::vector<liquid_float_complex> syms;
    int header_valid, payload_valid;
    framesyncstats_s stats;
};
unsigned batch_samples()
{
    const char *e = getenv("LQB_COMPAT_BATCH");
    long v = e ? atol(e) : 4096;
    return (unsigned)(v < 1 ? 1 : v);
}
}  // namespace

struct flexframesync_s {
    lqb_rx rx;
    framesync_callback cb;
    void *ud;
    std::vector<float> pend;
    std::deque<QueuedFrame> q;
    QueuedFrame cur;                 // buffers handed to the callback stay alive until the next delivery
    unsigned batch;
};

extern "C" flexframesync flexframesync_create(framesync_callback callback, void *userdata)
{
    lqb_rx_opts o = { 0, 1, 1u << 22, 0, NULL };      // one stream; carry sized for any 65535-byte frame at >= 1 bit/symbol, rate >= 1/4
    lqb_rx rx = lqb_rx_create(&o);
    if (!rx) return NULL;
    flexframesync q = new flexframesync_s;
    q->rx = rx; q->cb = callback; q->ud = userdata; q->batch = batch_samples();
    return q;
}
extern "C" void flexframesync_destroy(flexframesync q) { if (q) { lqb_rx_destroy(q->rx); delete q; } }
extern "C" void flexframesync_reset(flexframesync q) { if (q) { lqb_rx_reset(q->rx, -1); q->pend.clear(); q->q.clear(); } }

namespace {
// run the receiver over everything pending and queue the completed frames
void run_pending(flexframesync q)
{
    if (q->pend.empty()) return;
    const float *p = q->pend.data();
    uint64_t len = q->pend.size() / 2;
    if (lqb_rx_execute(q->rx, 1, NULL, &p, &len, LQB_MEM_HOST) == 0) {
        uint64_t frames = 0;
        lqb_rx_counts(q->rx, &frames, NULL);
        std::vector<lqb_frame_result> res((size_t)frames + 1);
        uint32_t got = 0;
        lqb_rx_poll(q->rx, res.data(), (uint32_t)frames, &got);
        for (uint32_t i = 0; i < got && i < frames; ++i) {
            const lqb_frame_result &r = res[i];
            if (r.flags & 1u) continue;               // longer than the carry (cannot happen with the 4 Mi-sample carry used here)
            QueuedFrame f;
            f.header.assign(r.header, r.header + 20);
            if (r.payload) f.payload.assign(r.payload, r.payload + r.payload_len);
            if (r.framesyms) {
                const liquid_float_complex *s = reinterpret_cast<const liquid_float_complex *>(r.framesyms);
                f.syms.assign(s, s + r.num_framesyms);
            }
            f.header_valid = r.header_valid; f.payload_valid = r.payload_valid;
            f.stats.evm = r.evm; f.stats.rssi = r.rssi; f.stats.cfo = r.cfo;
            f.stats.framesyms = NULL; f.stats.num_framesyms = r.header_valid ? r.num_framesyms : 0;
            f.stats.mod_scheme = r.mod_scheme; f.stats.mod_bps = r.mod_bps; f.stats.check = r.check;
            f.stats.fec0 = r.fec0; f.stats.fec1 = r.fec1;
            q->q.push_back(std::move(f));
        }
    }
    q->pend.clear();
}
// deliver up to `budget` queued frames; the buffers of a delivered frame live until the next delivery (liquid's rule)
void deliver(flexframesync q, size_t budget)
{
    while (budget-- && !q->q.empty() && q->cb) {
        q->cur = std::move(q->q.front());
        q->q.pop_front();
        QueuedFrame &f = q->cur;
        f.stats.framesyms = f.syms.empty() ? NULL : f.syms.data();
        q->cb(f.header.data(), f.header_valid, f.header_valid ? f.payload.data() : NULL,
              f.header_valid ? (unsigned)f.payload.size() : 0, f.payload_valid, f.stats, q->ud);
    }
}
}  // namespace

// Samples are batched (LQB_COMPAT_BATCH, default 4096) before they go to the GPU, so callbacks come later than in
// liquid-dsp but in the same order.  Delivery is paced at one callback per 256 samples of the current call: the
// reference's loop (256 samples per call, one result slot, lib/flex_rx_impl.cc:212-251) therefore never sees two
// callbacks in one call, while a caller that passes a large buffer gets every completed frame before execute returns
// (a frame is at least 618 samples long, so the queue cannot grow).
extern "C" void flexframesync_execute(flexframesync q, liquid_float_complex *x, unsigned int n)
{
    if (!q) return;
    const float *xf = reinterpret_cast<const float *>(x);
    q->pend.insert(q->pend.end(), xf, xf + 2 * (size_t)n);
    if (q->pend.size() / 2 >= q->batch) run_pending(q);
    deliver(q, ((size_t)n + 255) / 256);
}

// Extension (not in liquid-dsp): process the samples still waiting for a full batch and deliver every queued frame.
// Call it at the end of a capture; liquid-dsp has nothing to flush because it works sample by sample.
extern "C" void flexframesync_flush(flexframesync q)
{
    if (!q) return;
    run_pending(q);
    deliver(q, (size_t)-1);
}

// ------------------------------------------------------------------ flexframegen
struct flexframegen_s {
    lqb_tx tx;
    lqb_tx_props props;
    std::vector<liquid_float_complex> frame;
    size_t pos;
};

extern "C" void flexframegenprops_init_default(flexframegenprops_s *p)
{
    lqb_tx_props d;
    lqb_tx_props_init_default(&d);
    p->check = d.check; p->fec0 = d.fec0; p->fec1 = d.fec1; p->mod_scheme = d.mod_scheme;
}
extern "C" flexframegen flexframegen_create(flexframegenprops_s *props)
{
    lqb_tx_opts o = { 0, 0, NULL };
    lqb_tx tx = lqb_tx_create(&o);
    if (!tx) return NULL;
    flexframegen q = new flexframegen_s;
    q->tx = tx; q->pos = 0;
    lqb_tx_props_init_default(&q->props);
    if (props) flexframegen_setprops(q, props);
    return q;
}
extern "C" void flexframegen_destroy(flexframegen q) { if (q) { lqb_tx_destroy(q->tx); delete q; } }
extern "C" int flexframegen_setprops(flexframegen q, flexframegenprops_s *p)
{
    if (!q || !p) return -1;
    q->props.check = p->check; q->props.fec0 = p->fec0; q->props.fec1 = p->fec1; q->props.mod_scheme = p->mod_scheme;
    return 0;
}
extern "C" void flexframegen_assemble(flexframegen q, const unsigned char *header, const unsigned char *payload, unsigned int payload_len)
{
    if (!q) return;
    uint32_t n = 0, len = payload_len;
    q->frame.clear(); q->pos = 0;
    if (lqb_tx_frame_len(&q->props, len, &n) != 0) { lqb_internal_set_error("flexframegen_assemble: unsupported properties"); return; }
    q->frame.resize(n);
    float *out = reinterpret_cast<float *>(q->frame.data());
    if (lqb_tx_assemble(q->tx, 1, &q->props, &header, &payload, &len, &out, LQB_MEM_HOST) != 0) q->frame.clear();
}
extern "C" unsigned int flexframegen_getframelen(flexframegen q) { return q ? (unsigned)q->frame.size() : 0; }
extern "C" int flexframegen_write_samples(flexframegen q, liquid_float_complex *buffer, unsigned int buffer_len)
{
    if (!q) return 1;
    for (unsigned i = 0; i < buffer_len; ++i)
        buffer[i] = q->pos < q->frame.size() ? q->frame[q->pos++] : liquid_float_complex(0.0f, 0.0f);
    return q->pos >= q->frame.size();
}

// ------------------------------------------------------------------ qdetector_cccf
struct qdetector_cccf_s {
    lqb_det det;
    float beta, threshold;
    std::vector<float> pend;                      // samples not yet sent to the GPU (one 256-sample hop)
    std::vector<liquid_float_complex> hist;       // recent samples, hist[i] has absolute index hist_base + i
    int64_t hist_base, n_in;
    std::deque<lqb_detection> q;
    std::vector<liquid_float_complex> aligned;    // the 512 samples returned to the caller
    float tau, gamma, dphi, phi;
};

static bool is_flexframe_preamble(const liquid_float_complex *s, unsigned n)
{
    if (n != 64) return false;
    lqb::cf pn[64];
    lqb::preamble_pn(pn);
    for (unsigned i = 0; i < 64; ++i)
        if (std::fabs(s[i].real() - pn[i].re) > 1e-4f || std::fabs(s[i].imag() - pn[i].im) > 1e-4f) return false;
    return true;
}

extern "C" qdetector_cccf qdetector_cccf_create_linear(liquid_float_complex *sequence, unsigned int sequence_len,
                                                        int ftype, unsigned int k, unsigned int m, float beta)
{
    if (ftype != LIQUID_FIRFILT_ARKAISER || k != 2 || m != 7 || !is_flexframe_preamble(sequence, sequence_len)) {
        lqb_internal_set_error("qdetector_cccf_create_linear: only the 64-symbol flexframe preamble with ARKAISER k=2 m=7 is built in");
        return NULL;
    }
    qdetector_cccf q = new qdetector_cccf_s;
    q->det = NULL; q->beta = beta; q->threshold = 0.5f;        // liquid's default until set_threshold
    q->hist_base = 0; q->n_in = 0; q->tau = q->gamma = q->dphi = q->phi = 0.0f;
    q->aligned.assign(512, liquid_float_complex(0.0f, 0.0f));
    return q;
}
extern "C" void qdetector_cccf_destroy(qdetector_cccf q) { if (q) { if (q->det) lqb_det_destroy(q->det); delete q; } }
extern "C" void qdetector_cccf_reset(qdetector_cccf q)
{
    if (!q) return;
    if (q->det) lqb_det_reset(q->det, -1);
    q->pend.clear(); q->q.clear(); q->hist.clear(); q->hist_base = q->n_in;
}
extern "C" void qdetector_cccf_set_threshold(qdetector_cccf q, float threshold)
{
    if (!q) return;
    q->threshold = threshold;
    if (q->det) { lqb_det_destroy(q->det); q->det = NULL; }   // the threshold lives in the device tables: rebuild lazily
}
extern "C" void *qdetector_cccf_execute(qdetector_cccf q, liquid_float_complex x)
{
    if (!q) return NULL;
    if (!q->det) {
        lqb_det_opts o = { 0, 1, q->beta, q->threshold, 0.0f, NULL };
        q->det = lqb_det_create(&o);
        if (!q->det) return NULL;
    }
    q->pend.push_back(x.real()); q->pend.push_back(x.imag());
    q->hist.push_back(x);
    q->n_in++;
    if (q->pend.size() >= 512) {                   // one 256-sample hop
        const float *p = q->pend.data();
        uint64_t len = q->pend.size() / 2;
        if (lqb_det_execute(q->det, 1, NULL, &p, &len, LQB_MEM_HOST) == 0) {
            uint32_t n = 0;
            lqb_det_poll(q->det, NULL, 0, &n);
            std::vector<lqb_detection> d(n + 1);
            lqb_det_poll(q->det, d.data(), n, &n);
            for (uint32_t i = 0; i < n; ++i) q->q.push_back(d[i]);
        }
        q->pend.clear();
        if (q->hist.size() > 8192) {               // keep the last 4096 samples for the aligned buffer
            size_t drop = q->hist.size() - 4096;
            q->hist.erase(q->hist.begin(), q->hist.begin() + drop);
            q->hist_base += (int64_t)drop;
        }
    }
    if (q->q.empty()) return NULL;
    lqb_detection d = q->q.front();
    q->q.pop_front();
    q->tau = d.tau_hat; q->gamma = d.gamma_hat; q->dphi = d.dphi_hat; q->phi = d.phi_hat;
    for (int i = 0; i < 512; ++i) {
        int64_t a = d.sample_index + i - q->hist_base;
        q->aligned[i] = (a >= 0 && a < (int64_t)q->hist.size()) ? q->hist[(size_t)a] : liquid_float_complex(0.0f, 0.0f);
    }
    return q->aligned.data();
}
extern "C" float qdetector_cccf_get_tau(qdetector_cccf q) { return q ? q->tau : 0.0f; }
extern "C" float qdetector_cccf_get_gamma(qdetector_cccf q) { return q ? q->gamma : 0.0f; }
extern "C" float qdetector_cccf_get_dphi(qdetector_cccf q) { return q ? q->dphi : 0.0f; }
extern "C" float qdetector_cccf_get_phi(qdetector_cccf q) { return q ? q->phi : 0.0f; }
extern "C" unsigned int qdetector_cccf_get_buf_len(qdetector_cccf) { return 512; }
