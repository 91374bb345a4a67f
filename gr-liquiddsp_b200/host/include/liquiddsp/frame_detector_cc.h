// gr::liquiddsp::frame_detector_cc -- drop-in for the reference's
// include/liquiddsp/frame_detector_cc.h:40-49: `static sptr make()`, one complex stream in and the
// same stream out (lib/frame_detector_cc_impl.cc:43-44, :82).
#ifndef INCLUDED_LIQUIDDSP_FRAME_DETECTOR_CC_H
#define INCLUDED_LIQUIDDSP_FRAME_DETECTOR_CC_H
#include <liquiddsp/api.h>
#include <gnuradio/sync_block.h>

namespace gr { namespace liquiddsp {
class LIQUIDDSP_API frame_detector_cc : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<frame_detector_cc> sptr;
    static sptr make();
    virtual unsigned long frames_detected() const = 0;   // extension: the counter the reference only prints
};
}}
#endif
