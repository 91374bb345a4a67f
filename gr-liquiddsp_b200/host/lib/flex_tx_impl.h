#ifndef INCLUDED_LIQUIDDSP_FLEX_TX_IMPL_H
#define INCLUDED_LIQUIDDSP_FLEX_TX_IMPL_H
#include <liquiddsp/flex_tx.h>
#include "../../../include/lqb200.h"

namespace gr { namespace liquiddsp {

// Host side of flex_tx: message-only block, PDU bytes in -> one whole frame of samples out
// (reference: lib/flex_tx_impl.cc:191-209), scheme selection by index (:75-189).
class flex_tx_impl : public flex_tx {
public:
    flex_tx_impl(unsigned int modulation, unsigned int inner_code, unsigned int outer_code);
    ~flex_tx_impl();
    void send_pkt(pmt::pmt_t pdu);
    void configure(pmt::pmt_t configuration);
    void set_modulation(unsigned int modulation);
    void set_inner_code(unsigned int inner_code);
    void set_outer_code(unsigned int outer_code);
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
    const lqb_tx_props &props() const { return d_props; }
private:
    lqb_tx d_tx;
    lqb_tx_props d_props;
    unsigned char d_header[14];
    unsigned long d_num_frames;
};

}}
#endif
