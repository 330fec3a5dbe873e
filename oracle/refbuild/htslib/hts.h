// TEST INFRASTRUCTURE ONLY -- stand-in for <htslib/hts.h> (HTSlib is absent from this image).
// src/rcpp_get_base_freqs.cpp uses one symbol of it: seq_nt16_int, HTSlib's table from the 4-bit BAM base code
// (=ACMGRSVTWYHKDBN) to 0..3 for A,C,G,T and 4 for everything else.  Defined in refbuild/ref_driver.cpp.
#pragma once
extern "C" const int seq_nt16_int[16];
