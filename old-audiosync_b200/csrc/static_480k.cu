// Static four-step kernels of the L = 480k plan (see static_plan_impl.cuh).
#include "static_plan_impl.cuh"
template int asc::build_static_plan<asc::Plan480k>(asc::FftPlan*);
