// Static four-step kernels of the L = 288k plan (see static_plan_impl.cuh).
#include "static_plan_impl.cuh"
template int asc::build_static_plan<asc::Plan288k>(asc::FftPlan*);
