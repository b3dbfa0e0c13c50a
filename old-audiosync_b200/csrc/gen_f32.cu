// Runtime-radix four-step kernels, float arithmetic (see gen_impl.cuh).
#include "gen_impl.cuh"
template int asc::build_generic_plan_t<float>(asc::FftPlan*);
