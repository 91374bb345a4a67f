// frame_detector_cc block over the B200 detector (C-ABI in include/lqb200.h).
#include "frame_detector_cc_impl.h"
#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

namespace gr { namespace liquiddsp {

frame_detector_cc::sptr frame_detector_cc::make() { return gnuradio::get_initial_sptr(new frame_detector_cc_impl()); }

frame_detector_cc_impl::frame_detector_cc_impl()
    : gr::sync_block("frame_detector_cc", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_det(NULL), d_bulk(NULL), d_bulk_workers(0), d_num_frames(0)
{
    lqb_det_opts o = { 0, 1, 0.0f, 0.0f, 0.0f, NULL };   // zeros select the frame_detector_cc constants
    d_det = lqb_det_create(&o);
    if (!d_det) throw std::runtime_error(std::string("frame_detector_cc: ") + lqb_last_error());
}

frame_detector_cc_impl::~frame_detector_cc_impl() { lqb_det_destroy(d_det); if (d_bulk) lqb_det_destroy(d_bulk); }

long frame_detector_cc_impl::detect_capture(const gr_complex *samples, size_t n_samples, long long *indices, size_t max_out,
                                            unsigned workers, unsigned seg_len, unsigned preroll)
{
    if (!samples && n_samples) return -1;
    if (!seg_len) seg_len = 1u << 18;
    const size_t n_seg = (n_samples + seg_len - 1) / seg_len;
    workers = (unsigned)std::max<size_t>(1, std::min<size_t>(workers ? workers : 1024, n_seg));
    if (!d_bulk || d_bulk_workers < workers) {
        if (d_bulk) lqb_det_destroy(d_bulk);
        lqb_det_opts o = { 0, workers, 0.0f, 0.0f, 0.0f, NULL };
        d_bulk = lqb_det_create(&o);
        d_bulk_workers = d_bulk ? workers : 0;
        if (!d_bulk) return -1;
    }
    if (lqb_det_execute_sharded(d_bulk, reinterpret_cast<const float *>(samples), n_samples, LQB_MEM_HOST, seg_len, preroll) != 0) return -1;
    uint32_t found = 0;
    lqb_det_poll(d_bulk, NULL, 0, &found);
    std::vector<lqb_detection> det(found ? found : 1);
    lqb_det_poll(d_bulk, det.data(), found, &found);
    for (uint32_t i = 0; i < found; ++i) {
        std::cout << "Detected " << d_num_frames << " frames!" << std::endl;
        d_num_frames++;
        if (indices && i < max_out) indices[i] = (long long)det[i].sample_index;
    }
    lqb_det_reset(d_bulk, -1);
    return (long)found;
}

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
