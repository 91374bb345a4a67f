#!/bin/bash
# Build the block classes against the GNU Radio TEST SHIM (host/gr_shim) into
# gr-liquiddsp_b200/lib/liblqb_blocks_test.so for the unit tests.  With a real GNU Radio 3.7 tree the
# same sources build through host/CMakeLists.txt instead.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
LIB="$HERE/../lib"
g++ -std=c++17 -O2 -fPIC -shared -Wall -I"$HERE/gr_shim" -I"$HERE/include" -I"$HERE" \
    "$HERE/lib/flex_rx_impl.cc" "$HERE/lib/flex_tx_impl.cc" "$HERE/lib/frame_detector_cc_impl.cc" "$HERE/test_harness.cc" \
    -L"$LIB" -llqb200 -Wl,-rpath,'$ORIGIN' -o "$LIB/liblqb_blocks_test.so"
