// gr::liquiddsp::flex_rx -- public block interface, drop-in for the reference's
// include/liquiddsp/flex_rx.h:40-50: sptr typedef and `static sptr make()` (no arguments);
// one optional complex input, no stream output; message ports constellation / payload_data /
// packet_info (lib/flex_rx_impl.cc:44-46, :61-63).  make_multi() is an additive extension that
// batches N channels on one GPU handle (one complex input per channel).
#ifndef INCLUDED_LIQUIDDSP_FLEX_RX_H
#define INCLUDED_LIQUIDDSP_FLEX_RX_H
#include <liquiddsp/api.h>
#include <gnuradio/sync_block.h>
#include <pmt/pmt.h>

namespace gr { namespace liquiddsp {
class LIQUIDDSP_API flex_rx : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<flex_rx> sptr;
    static sptr make();
    static sptr make_multi(unsigned n_channels, int device = 0);
    // Additive (not in the reference): decode ONE recorded capture of this block's single channel at batch speed and
    // publish its frames exactly as work() would have, in order.  The capture is cut in time over `workers` GPU
    // streams with an exact seam rule (lqb_rx_execute_sharded): the messages are those of feeding the capture through
    // work() chunk by chunk.  Returns the number of frames, -1 on error.  Only for blocks made by make() / n_channels 1.
    virtual long decode_capture(const gr_complex *samples, size_t n_samples, unsigned workers = 512,
                                unsigned seg_len = 1u << 20, unsigned preroll = 1u << 16) = 0;
};
}}
#endif
