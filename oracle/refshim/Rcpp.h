// Rcpp.h — stand-in (TEST INFRASTRUCTURE): the Rcpp half of RcppArmadillo.h in this directory, so that
// shim/quilt_gpu_shim.cpp can be compiled, linked and EXECUTED without R (tests/test_shim_executes.py).
#ifndef REFSHIM_RCPP_H
#define REFSHIM_RCPP_H
#include "RcppArmadillo.h"
#endif
