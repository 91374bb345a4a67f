// flex_tx block over the B200 frame generator (C-ABI in include/lqb200.h).
#include "flex_tx_impl.h"
#include "scheme_tables.h"
#include <cstring>
#include <iostream>
#include <stdexcept>

namespace gr { namespace liquiddsp {

flex_tx::sptr flex_tx::make(unsigned int modulation, unsigned int inner_code, unsigned int outer_code)
{
    return gnuradio::get_initial_sptr(new flex_tx_impl(modulation, inner_code, outer_code));
}

flex_tx_impl::flex_tx_impl(unsigned int modulation, unsigned int inner_code, unsigned int outer_code)
    : gr::sync_block("flex_tx", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)), d_tx(NULL), d_num_frames(0)
{
    lqb_tx_props_init_default(&d_props);
    d_props.check = LQB_CRC_24;                 // payload check is fixed, as in the reference (:52)
    set_inner_code(inner_code);
    set_outer_code(outer_code);
    set_modulation(modulation);
    std::memset(d_header, 0, sizeof d_header);  // 14 user header bytes, all zero (:58-59)
    lqb_tx_opts o = { 0, 0, NULL };
    d_tx = lqb_tx_create(&o);
    if (!d_tx) throw std::runtime_error(std::string("flex_tx: ") + lqb_last_error());
    message_port_register_out(pmt::mp("pdus"));
    message_port_register_in(pmt::mp("pdus"));
    set_msg_handler(pmt::mp("pdus"), [this](pmt::pmt_t m) { this->send_pkt(m); });
    message_port_register_in(pmt::mp("configuration"));
    set_msg_handler(pmt::mp("configuration"), [this](pmt::pmt_t m) { this->configure(m); });
}

flex_tx_impl::~flex_tx_impl() { lqb_tx_destroy(d_tx); }

void flex_tx_impl::set_modulation(unsigned int modulation)
{
    if (modulation < sizeof tables::kModulation / sizeof tables::kModulation[0]) { d_props.mod_scheme = tables::kModulation[modulation]; return; }
    std::cout << "Unsupported Modulation Defaulting to BPSK." << std::endl;
    d_props.mod_scheme = LQB_MODEM_PSK2;
}
void flex_tx_impl::set_inner_code(unsigned int inner_code)
{
    if (inner_code < sizeof tables::kInner / sizeof tables::kInner[0]) { d_props.fec0 = tables::kInner[inner_code]; return; }
    std::cout << "Unsupported FEC Defaulting to none." << std::endl;
    d_props.fec0 = LQB_FEC_NONE;
}
void flex_tx_impl::set_outer_code(unsigned int outer_code)
{
    if (outer_code < sizeof tables::kOuter / sizeof tables::kOuter[0]) { d_props.fec1 = tables::kOuter[outer_code]; return; }
    std::cout << "Unsupported FEC Defaulting to none." << std::endl;
    d_props.fec1 = LQB_FEC_NONE;
}

void flex_tx_impl::configure(pmt::pmt_t cfg)
{
    // each key is optional; the new properties apply from the next frame
    if (pmt::dict_has_key(cfg, pmt::mp("modulation"))) set_modulation((unsigned)pmt::to_long(pmt::dict_ref(cfg, pmt::mp("modulation"), pmt::PMT_NIL)));
    if (pmt::dict_has_key(cfg, pmt::mp("inner_code"))) set_inner_code((unsigned)pmt::to_long(pmt::dict_ref(cfg, pmt::mp("inner_code"), pmt::PMT_NIL)));
    if (pmt::dict_has_key(cfg, pmt::mp("outer_code"))) set_outer_code((unsigned)pmt::to_long(pmt::dict_ref(cfg, pmt::mp("outer_code"), pmt::PMT_NIL)));
}

void flex_tx_impl::send_pkt(pmt::pmt_t pdu)
{
    pmt::pmt_t bytes = pmt::cdr(pdu);           // car(pdu) is metadata and is ignored
    std::vector<uint8_t> payload = pmt::u8vector_elements(bytes);
    uint32_t frame_len = 0, plen = (uint32_t)payload.size();
    if (lqb_tx_frame_len(&d_props, plen, &frame_len) != 0) throw std::runtime_error("flex_tx: payload too long for one frame");
    std::vector<gr_complex> vec(frame_len);
    const uint8_t *hp = d_header, *pp = payload.data();
    float *op = reinterpret_cast<float *>(vec.data());
    if (lqb_tx_assemble(d_tx, 1, &d_props, &hp, &pp, &plen, &op, LQB_MEM_HOST) != 0)
        throw std::runtime_error(std::string("flex_tx: ") + lqb_last_error());
    message_port_pub(pmt::mp("pdus"), pmt::cons(pmt::PMT_NIL, pmt::init_c32vector(frame_len, vec)));
    d_num_frames++;
}

int flex_tx_impl::work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &)
{
    throw std::runtime_error("This is not a stream block.");
    return noutput_items;
}

}}
