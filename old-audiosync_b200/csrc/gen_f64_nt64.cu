// Runtime-radix four-step kernels, double arithmetic, 64 threads per CTA (see gen_impl.cuh).
#include "gen_impl.cuh"
template struct asc::GenStage<double, 64>;
