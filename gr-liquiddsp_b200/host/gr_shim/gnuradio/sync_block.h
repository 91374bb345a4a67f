// sync_block.h (TEST SHIM) -- minimal gr::sync_block / gr::io_signature for exercising the block
// sources without GNU Radio: ports, message handlers, set_output_multiple, and a message log that
// the test harness reads.  Real deployments compile the same sources against GNU Radio 3.7.
#pragma once
#include <pmt/pmt.h>
#include <complex>
#include <deque>
#include <functional>
#include <map>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {

class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int item_size) { return sptr(new io_signature(min_streams, max_streams, item_size)); }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int) const { return d_size; }
private:
    io_signature(int a, int b, int c) : d_min(a), d_max(b), d_size(c) {}
    int d_min, d_max, d_size;
};

class sync_block {
public:
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out), d_multiple(1) {}
    virtual ~sync_block() {}
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
    const std::string &name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }
    void set_output_multiple(int m) { d_multiple = m; }
    int output_multiple() const { return d_multiple; }
    void message_port_register_out(pmt::pmt_t port) { d_out_ports.push_back(pmt::symbol_to_string(port)); }
    void message_port_register_in(pmt::pmt_t port) { d_in_ports.push_back(pmt::symbol_to_string(port)); }
    template <class F> void set_msg_handler(pmt::pmt_t port, F f) { d_handlers[pmt::symbol_to_string(port)] = f; }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg) { d_log.push_back(std::make_pair(pmt::symbol_to_string(port), msg)); }
    // ---- test-side access
    void post(const std::string &port, pmt::pmt_t msg) { d_handlers.at(port)(msg); }
    std::deque<std::pair<std::string, pmt::pmt_t>> &published() { return d_log; }
    const std::vector<std::string> &out_ports() const { return d_out_ports; }
    const std::vector<std::string> &in_ports() const { return d_in_ports; }
private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    int d_multiple;
    std::vector<std::string> d_out_ports, d_in_ports;
    std::map<std::string, std::function<void(pmt::pmt_t)>> d_handlers;
    std::deque<std::pair<std::string, pmt::pmt_t>> d_log;
};

namespace block { template <class T> boost::shared_ptr<T> make_sptr(T *p) { return boost::shared_ptr<T>(p); } }

}  // namespace gr

namespace gnuradio { template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); } }
