// test_harness.cc -- C entry points that let the Python tests drive the three block classes through
// their GNU Radio-facing interface (make(), work(), message ports) on the test shim.
#include <liquiddsp/flex_rx.h>
#include <liquiddsp/flex_tx.h>
#include <liquiddsp/frame_detector_cc.h>
#include "lib/flex_rx_impl.h"
#include "lib/flex_tx_impl.h"
#include <cstring>
#include <string>

using namespace gr::liquiddsp;

struct blk { boost::shared_ptr<gr::sync_block> b; std::string err; };

extern "C" {

void *blk_make_flex_tx(unsigned m, unsigned i, unsigned o) { try { blk *h = new blk; h->b = flex_tx::make(m, i, o); return h; } catch (...) { return NULL; } }
void *blk_make_flex_rx(void) { try { blk *h = new blk; h->b = flex_rx::make(); return h; } catch (...) { return NULL; } }
void *blk_make_flex_rx_multi(unsigned n) { try { blk *h = new blk; h->b = flex_rx::make_multi(n); return h; } catch (...) { return NULL; } }
void *blk_make_frame_detector(void) { try { blk *h = new blk; h->b = frame_detector_cc::make(); return h; } catch (...) { return NULL; } }
long blk_rx_decode_capture(void *p, const float *iq, unsigned long n, unsigned workers, unsigned seg_len, unsigned preroll)
{
    flex_rx *rx = dynamic_cast<flex_rx *>(static_cast<blk *>(p)->b.get());
    if (!rx) return -2;
    try { return rx->decode_capture(reinterpret_cast<const gr_complex *>(iq), n, workers, seg_len, preroll); } catch (...) { return -3; }
}
long blk_det_detect_capture(void *p, const float *iq, unsigned long n, long long *idx, unsigned long max_out, unsigned workers, unsigned seg_len, unsigned preroll)
{
    frame_detector_cc *d = dynamic_cast<frame_detector_cc *>(static_cast<blk *>(p)->b.get());
    if (!d) return -2;
    try { return d->detect_capture(reinterpret_cast<const gr_complex *>(iq), n, idx, max_out, workers, seg_len, preroll); } catch (...) { return -3; }
}
unsigned long blk_det_frames(void *p)
{
    frame_detector_cc *d = dynamic_cast<frame_detector_cc *>(static_cast<blk *>(p)->b.get());
    return d ? d->frames_detected() : 0;
}
void blk_destroy(void *p) { delete static_cast<blk *>(p); }
const char *blk_name(void *p) { return static_cast<blk *>(p)->b->name().c_str(); }
int blk_output_multiple(void *p) { return static_cast<blk *>(p)->b->output_multiple(); }
int blk_sig(void *p, int out, int which)
{
    gr::io_signature::sptr s = out ? static_cast<blk *>(p)->b->output_signature() : static_cast<blk *>(p)->b->input_signature();
    return which == 0 ? s->min_streams() : which == 1 ? s->max_streams() : s->sizeof_stream_item(0);
}
int blk_ports(void *p, int out, char *buf, int cap)
{
    const std::vector<std::string> &v = out ? static_cast<blk *>(p)->b->out_ports() : static_cast<blk *>(p)->b->in_ports();
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + v[i];
    std::strncpy(buf, s.c_str(), cap - 1); buf[cap - 1] = 0;
    return (int)v.size();
}
const char *blk_error(void *p) { return static_cast<blk *>(p)->err.c_str(); }

// work() on n_in input streams of n items each (inputs contiguous: stream c at in + c*n); out may be NULL
int blk_work(void *p, const float *in, int n_in, int n, float *out)
{
    blk *h = static_cast<blk *>(p);
    gr_vector_const_void_star iv; gr_vector_void_star ov;
    for (int c = 0; c < n_in; ++c) iv.push_back(in + 2 * (size_t)c * n);
    if (out) ov.push_back(out);
    try { return h->b->work(n, iv, ov); } catch (std::exception &e) { h->err = e.what(); return -1; }
}
// post a PDU (NIL . u8vector) to an input message port
int blk_post_pdu(void *p, const char *port, const unsigned char *bytes, int n)
{
    blk *h = static_cast<blk *>(p);
    try { h->b->post(port, pmt::cons(pmt::PMT_NIL, pmt::init_u8vector(n, bytes))); return 0; } catch (std::exception &e) { h->err = e.what(); return -1; }
}
// post a dict of longs; keys comma separated; negative count of keys allowed = empty dict
int blk_post_dict(void *p, const char *port, const char *keys, const long *vals, int n)
{
    blk *h = static_cast<blk *>(p);
    pmt::pmt_t d = pmt::make_dict();
    std::string ks(keys); size_t pos = 0;
    for (int i = 0; i < n; ++i) {
        size_t c = ks.find(',', pos);
        d = pmt::dict_add(d, pmt::mp(ks.substr(pos, c == std::string::npos ? c : c - pos)), pmt::from_long(vals[i]));
        pos = c == std::string::npos ? ks.size() : c + 1;
    }
    try { h->b->post(port, d); return 0; } catch (std::exception &e) { h->err = e.what(); return -1; }
}
int blk_pending(void *p) { return (int)static_cast<blk *>(p)->b->published().size(); }
// pops the oldest published message: returns kind (1 = PDU c32vector, 2 = PDU u8vector, 3 = dict), fills port name;
// data: c32 -> floats (2 per item), u8 -> bytes, dict -> "key=value;" text.  *n = item count
int blk_pop(void *p, char *port, int port_cap, void *data, int data_cap_bytes, int *n)
{
    blk *h = static_cast<blk *>(p);
    if (h->b->published().empty()) return 0;
    std::pair<std::string, pmt::pmt_t> m = h->b->published().front();
    h->b->published().pop_front();
    std::strncpy(port, m.first.c_str(), port_cap - 1); port[port_cap - 1] = 0;
    pmt::pmt_t v = m.second;
    if (v->kind == pmt::pmt_base::PAIR) v = pmt::cdr(v);
    if (v->kind == pmt::pmt_base::C32VEC) {
        *n = (int)v->c32.size();
        size_t bytes = v->c32.size() * 8; if ((int)bytes > data_cap_bytes) bytes = data_cap_bytes;
        std::memcpy(data, v->c32.data(), bytes);
        return 1;
    }
    if (v->kind == pmt::pmt_base::U8VEC) {
        *n = (int)v->u8.size();
        size_t bytes = v->u8.size(); if ((int)bytes > data_cap_bytes) bytes = data_cap_bytes;
        std::memcpy(data, v->u8.data(), bytes);
        return 2;
    }
    if (v->kind == pmt::pmt_base::DICT) {
        std::string s;
        for (auto &kv : v->dict) s += kv.first + "=" + std::to_string(pmt::to_long(kv.second)) + ";";
        *n = (int)v->dict.size();
        std::strncpy(static_cast<char *>(data), s.c_str(), data_cap_bytes - 1);
        static_cast<char *>(data)[data_cap_bytes - 1] = 0;
        return 3;
    }
    return -1;
}
int blk_rx_index(int which, unsigned scheme)
{
    return which == 0 ? flex_rx_impl::mod_index(scheme) : which == 1 ? flex_rx_impl::inner_index(scheme) : flex_rx_impl::outer_index(scheme);
}
int blk_tx_props(void *p, unsigned *out4)
{
    flex_tx_impl *t = dynamic_cast<flex_tx_impl *>(static_cast<blk *>(p)->b.get());
    if (!t) return -1;
    out4[0] = t->props().mod_scheme; out4[1] = t->props().check; out4[2] = t->props().fec0; out4[3] = t->props().fec1;
    return 0;
}

}
