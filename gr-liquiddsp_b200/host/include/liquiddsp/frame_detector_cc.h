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
    // Additive: search ONE recorded capture at batch speed (cut in time over `workers` GPU streams with an exact seam
    // rule, lqb_det_execute_sharded): prints and counts what work() would have over the same samples; returns the number
    // of detections in the capture (-1 on error), their absolute sample indices in *indices when given (up to max_out).
    virtual long detect_capture(const gr_complex *samples, size_t n_samples, long long *indices = 0, size_t max_out = 0,
                                unsigned workers = 1024, unsigned seg_len = 1u << 18, unsigned preroll = 1u << 14) = 0;
};
}}
#endif
