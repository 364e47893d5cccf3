// shim.cpp — kernel_wrapper_ccdpp_NV / kernel_wrapper_als_NV against this build's host containers.
#include "pmf.h"
#include "shim_impl.h"
