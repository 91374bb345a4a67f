#ifndef INCLUDED_LIQUIDDSP_FRAME_DETECTOR_CC_IMPL_H
#define INCLUDED_LIQUIDDSP_FRAME_DETECTOR_CC_IMPL_H
#include <liquiddsp/frame_detector_cc.h>
#include "../../../include/lqb200.h"

namespace gr { namespace liquiddsp {

// Host side of frame_detector_cc: pass-through stream block that counts preamble detections
// (reference: lib/frame_detector_cc_impl.cc:67-97; k=2, m=7, beta=0.3, threshold 0.45 from
// lib/frame_detector_cc_impl.h:34-36 and .cc:55 are the C-ABI defaults).
class frame_detector_cc_impl : public frame_detector_cc {
public:
    frame_detector_cc_impl();
    ~frame_detector_cc_impl();
    int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
    unsigned long frames_detected() const { return d_num_frames; }
    long detect_capture(const gr_complex *samples, size_t n_samples, long long *indices, size_t max_out,
                        unsigned workers, unsigned seg_len, unsigned preroll);
private:
    lqb_det d_det;
    lqb_det d_bulk;           // batch handle of detect_capture (made on first use)
    unsigned d_bulk_workers;
    unsigned long d_num_frames;
};

}}
#endif
