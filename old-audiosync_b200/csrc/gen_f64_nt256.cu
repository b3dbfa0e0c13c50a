// Runtime-radix four-step kernels, double arithmetic, 256 threads per CTA (see gen_impl.cuh).
#include "gen_impl.cuh"
template struct asc::GenStage<double, 256>;
