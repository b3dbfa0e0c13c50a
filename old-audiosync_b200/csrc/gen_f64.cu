// Runtime-radix four-step kernels, double arithmetic (see gen_impl.cuh).
#include "gen_impl.cuh"
template int asc::build_generic_plan_t<double>(asc::FftPlan*);
