/* -*- c++ -*- */
/* SWIG module for builds against GNU Radio 3.7: exposes the three blocks as
 * liquiddsp_swig.flex_tx / flex_rx / frame_detector_cc, the names `import liquiddsp` re-exports
 * (reference: swig/liquiddsp_swig.i:19-26). */
#define LIQUIDDSP_API

%include "gnuradio.i"

%{
#include "liquiddsp/flex_rx.h"
#include "liquiddsp/flex_tx.h"
#include "liquiddsp/frame_detector_cc.h"
%}

%include "liquiddsp/flex_rx.h"
GR_SWIG_BLOCK_MAGIC2(liquiddsp, flex_rx);
%include "liquiddsp/flex_tx.h"
GR_SWIG_BLOCK_MAGIC2(liquiddsp, flex_tx);
%include "liquiddsp/frame_detector_cc.h"
GR_SWIG_BLOCK_MAGIC2(liquiddsp, frame_detector_cc);
