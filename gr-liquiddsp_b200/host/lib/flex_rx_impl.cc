// flex_rx block over the B200 receiver (C-ABI in include/lqb200.h).
#include "flex_rx_impl.h"
#include "scheme_tables.h"
#include <algorithm>
#include <iostream>
#include <stdexcept>

namespace gr { namespace liquiddsp {

flex_rx::sptr flex_rx::make() { return gnuradio::get_initial_sptr(new flex_rx_impl(1, 0)); }
flex_rx::sptr flex_rx::make_multi(unsigned n_channels, int device) { return gnuradio::get_initial_sptr(new flex_rx_impl(n_channels, device)); }

flex_rx_impl::flex_rx_impl(unsigned n_channels, int device)
    : gr::sync_block("flex_rx",
                     gr::io_signature::make(n_channels == 1 ? 0 : (int)n_channels, (int)n_channels, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      d_rx(NULL), d_bulk(NULL), d_bulk_workers(0), d_device(device), d_channels(n_channels), d_num_frames(0)
{
    lqb_rx_opts o = { device, n_channels, 0, 0, NULL };
    d_rx = lqb_rx_create(&o);
    if (!d_rx) throw std::runtime_error(std::string("flex_rx: ") + lqb_last_error());
    set_output_multiple(d_inbuf_len);
    message_port_register_out(pmt::mp("constellation"));
    message_port_register_out(pmt::mp("payload_data"));
    message_port_register_out(pmt::mp("packet_info"));
}

flex_rx_impl::~flex_rx_impl() { lqb_rx_destroy(d_rx); if (d_bulk) lqb_rx_destroy(d_bulk); }

int flex_rx_impl::mod_index(unsigned ms)
{
    int i = tables::index_of(tables::kModulation, ms);
    if (i < 0) std::cout << "Unsupported Received Modulation Defaulting to BPSK." << std::endl;
    return i;
}
int flex_rx_impl::inner_index(unsigned fec0)
{
    int i = tables::index_of(tables::kInner, fec0);
    if (i < 0) std::cout << "Unsupported Received FEC Defaulting to none." << std::endl;
    return i;
}
int flex_rx_impl::outer_index(unsigned fec1)
{
    int i = tables::index_of(tables::kOuter, fec1);
    if (i < 0) std::cout << "Unsupported FEC received defaulting to none." << std::endl;
    return i;
}

void flex_rx_impl::publish(const lqb_frame_result &r)
{
    if (r.flags & 1u) {
        // a frame longer than the receiver's per-stream buffer (lqb_rx_opts.max_frame_samples) cannot be received:
        // it is announced here instead of vanishing silently; no message goes out because there are no payload bytes
        std::cout << "flex_rx: dropped a frame of " << r.payload_len << " payload bytes (longer than the receive buffer)" << std::endl;
        return;
    }
    // the constellation goes out for every frame, also when the header check failed (then it is empty)
    const gr_complex *syms = reinterpret_cast<const gr_complex *>(r.framesyms);
    message_port_pub(pmt::mp("constellation"), pmt::cons(pmt::PMT_NIL, pmt::init_c32vector(syms ? r.num_framesyms : 0, syms)));
    if (!r.header_valid) return;
    // payload bytes are published whether or not their CRC passed, as the reference does
    message_port_pub(pmt::mp("payload_data"), pmt::cons(pmt::PMT_NIL, pmt::init_u8vector(r.payload_len, r.payload)));
    pmt::pmt_t info = pmt::make_dict();
    info = pmt::dict_add(info, pmt::mp("header_valid"), pmt::from_long(1));
    info = pmt::dict_add(info, pmt::mp("payload_valid"), pmt::from_long((long)r.payload_valid));
    info = pmt::dict_add(info, pmt::mp("modulation"), pmt::from_long((long)mod_index(r.mod_scheme)));
    info = pmt::dict_add(info, pmt::mp("inner_code"), pmt::from_long((long)inner_index(r.fec0)));
    info = pmt::dict_add(info, pmt::mp("outer_code"), pmt::from_long((long)outer_index(r.fec1)));
    message_port_pub(pmt::mp("packet_info"), info);
    d_num_frames++;
}

long flex_rx_impl::decode_capture(const gr_complex *samples, size_t n_samples, unsigned workers, unsigned seg_len, unsigned preroll)
{
    if (d_channels != 1 || (!samples && n_samples)) return -1;
    if (!seg_len) seg_len = 1u << 20;
    const size_t n_seg = (n_samples + seg_len - 1) / seg_len;
    workers = (unsigned)std::max<size_t>(1, std::min<size_t>(workers ? workers : 512, n_seg));
    if (!d_bulk || d_bulk_workers < workers) {
        if (d_bulk) lqb_rx_destroy(d_bulk);
        lqb_rx_opts o = { d_device, workers, 0, 0, NULL };
        d_bulk = lqb_rx_create(&o);
        d_bulk_workers = d_bulk ? workers : 0;
        if (!d_bulk) return -1;
    }
    if (lqb_rx_execute_sharded(d_bulk, reinterpret_cast<const float *>(samples), n_samples, LQB_MEM_HOST, seg_len, preroll) != 0) return -1;
    uint64_t frames = 0;
    lqb_rx_counts(d_bulk, &frames, NULL);
    d_results.resize((size_t)frames + 1);
    uint32_t got = 0;
    lqb_rx_poll(d_bulk, d_results.data(), (uint32_t)frames, &got);
    for (uint32_t i = 0; i < got && i < frames; ++i) publish(d_results[i]);   // in time order
    lqb_rx_reset(d_bulk, -1);
    return (long)frames;
}

int flex_rx_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
    if (noutput_items % d_inbuf_len != 0) throw std::runtime_error("flex_rx: work() needs a multiple of 256 items");
    const unsigned n = (unsigned)input_items.size();
    if (n == 0 || noutput_items == 0) return noutput_items;
    std::vector<const float *> iq(n);
    std::vector<uint64_t> len(n, (uint64_t)noutput_items);
    for (unsigned c = 0; c < n; ++c) iq[c] = static_cast<const float *>(input_items[c]);
    if (lqb_rx_execute(d_rx, n, NULL, iq.data(), len.data(), LQB_MEM_HOST) != 0)
        throw std::runtime_error(std::string("flex_rx: ") + lqb_last_error());
    uint64_t frames = 0;
    lqb_rx_counts(d_rx, &frames, NULL);
    d_results.resize((size_t)frames + 1);
    uint32_t got = 0;
    lqb_rx_poll(d_rx, d_results.data(), (uint32_t)frames, &got);
    for (uint32_t i = 0; i < got && i < frames; ++i) publish(d_results[i]);   // ordered by (channel, frame)
    return noutput_items;
}

}}
