// frame_detector_cc block over the B200 detector (C-ABI in include/lqb200.h).
#include "frame_detector_cc_impl.h"
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

namespace gr { namespace liquiddsp {

frame_detector_cc::sptr frame_detector_cc::make() { return gnuradio::get_initial_sptr(new frame_detector_cc_impl()); }

frame_detector_cc_impl::frame_detector_cc_impl()
    : gr::sync_block("frame_detector_cc", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_det(NULL), d_num_frames(0)
{
    lqb_det_opts o = { 0, 1, 0.0f, 0.0f, 0.0f, NULL };   // zeros select the frame_detector_cc constants
    d_det = lqb_det_create(&o);
    if (!d_det) throw std::runtime_error(std::string("frame_detector_cc: ") + lqb_last_error());
}

frame_detector_cc_impl::~frame_detector_cc_impl() { lqb_det_destroy(d_det); }

int frame_detector_cc_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items)
{
    const float *in = static_cast<const float *>(input_items[0]);
    uint64_t n = (uint64_t)noutput_items;
    if (lqb_det_execute(d_det, 1, NULL, &in, &n, LQB_MEM_HOST) != 0)
        throw std::runtime_error(std::string("frame_detector_cc: ") + lqb_last_error());
    uint32_t found = 0;
    lqb_det_poll(d_det, NULL, 0, &found);
    for (uint32_t i = 0; i < found; ++i) {
        std::cout << "Detected " << d_num_frames << " frames!" << std::endl;
        d_num_frames++;
    }
    std::memcpy(output_items[0], input_items[0], (size_t)noutput_items * sizeof(gr_complex));   // stream passes through unchanged
    return noutput_items;
}

}}
