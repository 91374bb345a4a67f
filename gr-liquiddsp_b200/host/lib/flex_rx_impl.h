#ifndef INCLUDED_LIQUIDDSP_FLEX_RX_IMPL_H
#define INCLUDED_LIQUIDDSP_FLEX_RX_IMPL_H
#include <liquiddsp/flex_rx.h>
#include "../../../include/lqb200.h"
#include <vector>

namespace gr { namespace liquiddsp {

// Host side of flex_rx.  Where the reference feeds 256-sample chunks to one flexframesync and
// publishes from its callback slot (lib/flex_rx_impl.cc:204-254), this block hands the whole
// work() buffer of every channel to the batched GPU receiver in one call and then publishes the
// completed frames in order: constellation, payload_data, packet_info per frame.
class flex_rx_impl : public flex_rx {
public:
    flex_rx_impl(unsigned n_channels, int device);
    ~flex_rx_impl();
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
    long decode_capture(const gr_complex *samples, size_t n_samples, unsigned workers, unsigned seg_len, unsigned preroll);
    // liquid enum -> block-API index, -1 (and a note on stdout) when the scheme is outside the tables
    static int mod_index(unsigned mod_scheme);
    static int inner_index(unsigned fec0);
    static int outer_index(unsigned fec1);
    static const int d_inbuf_len = 256;
private:
    void publish(const lqb_frame_result &r);
    lqb_rx d_rx;
    lqb_rx d_bulk;            // batch handle of decode_capture (made on first use)
    unsigned d_bulk_workers;
    int d_device;
    unsigned d_channels;
    unsigned long d_num_frames;
    std::vector<lqb_frame_result> d_results;
};

}}
#endif
