// Runtime-radix four-step kernels, float arithmetic, 64 threads per CTA (see gen_impl.cuh).
#include "gen_impl.cuh"
template struct asc::GenStage<float, 64>;
