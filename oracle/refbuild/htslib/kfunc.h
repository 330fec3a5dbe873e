// TEST INFRASTRUCTURE ONLY -- stand-in for <htslib/kfunc.h> (HTSlib is absent from this image).
// src/rcpp_fep.cpp uses kt_fisher_exact; the definition linked here is oracle/refbuild/ref_driver.cpp's restatement
// of the published algorithm (hypergeometric tails summed from both ends while below the observed table's
// probability, with HTSlib's 1e-8 relative guard), pinned by the constants of inst/unitTests/test_generateVcfReport.R.
#pragma once
extern "C" double kt_fisher_exact(int n11, int n12, int n21, int n22, double* left, double* right, double* two);
