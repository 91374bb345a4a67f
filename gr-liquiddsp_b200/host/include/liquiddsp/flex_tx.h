// gr::liquiddsp::flex_tx -- drop-in for the reference's include/liquiddsp/flex_tx.h:40-52:
// make(modulation, inner_code, outer_code) with the reference's index tables, public
// set_modulation / set_inner_code / set_outer_code, message ports pdus (in/out) and configuration.
#ifndef INCLUDED_LIQUIDDSP_FLEX_TX_H
#define INCLUDED_LIQUIDDSP_FLEX_TX_H
#include <liquiddsp/api.h>
#include <gnuradio/sync_block.h>
#include <pmt/pmt.h>

namespace gr { namespace liquiddsp {
class LIQUIDDSP_API flex_tx : virtual public gr::sync_block {
public:
    typedef boost::shared_ptr<flex_tx> sptr;
    static sptr make(unsigned int modulation, unsigned int inner_code, unsigned int outer_code);
    virtual void set_modulation(unsigned int modulation) = 0;
    virtual void set_inner_code(unsigned int inner_code) = 0;
    virtual void set_outer_code(unsigned int outer_code) = 0;
};
}}
#endif
