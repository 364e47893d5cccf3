// dropin_shim.cpp — the same shim (cuda-recommender_b200/host/shim_impl.h) compiled against the
// REFERENCE's own containers: "pmf.h" below resolves to /root/reference/src/pmf.h through the -I of
// oracle/Makefile (this directory holds no pmf.h).  Linked with the reference's unmodified main.cpp
// and CPU sources it yields oracle/_ref/cuda_andre_dropin: the reference executable whose -CUDA path
// is this repo's library.  Test / demonstration artefact only.
#include "pmf.h"
#include "../cuda-recommender_b200/host/shim_impl.h"
