// gen_impl.cuh -- launches and plan construction of the runtime-radix kernels; included by
// gen_f32.cu / gen_f64.cu, which instantiate build_generic_plan_t<float / double>.
#pragma once

#include "plan_host.h"
#include "reduce_kernels.cuh"

namespace asc {

// ------------------------------------------------- runtime-radix four-step plans (any length)
template <class K>
static int prepare_gen_kernel(size_t smem) {
    if (smem > 48 * 1024)
        ASC_CUDA_OK(cudaFuncSetAttribute(gen_kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

template <typename T, typename InT>
static int run_generic_wave(FftPlan* plan, audiosync_cuda_ctx* ctx, DeviceState& d, const void* src, const void* smp,
                            long long sp, long long mp, void* ws, PairPeak* peaks, int pairs, cudaStream_t st) {
    typedef typename GenTraits<T>::C C;
    const GenShape& sh = plan->gen;
    C* planes = static_cast<C*>(ws);
    const C* wcol = static_cast<const C*>(plan->g_wcol.p);
    const C* wrow = static_cast<const C*>(plan->g_wrow.p);
    const C* lo = static_cast<const C*>(plan->g_lo.p);
    const C* hi = static_cast<const C*>(plan->g_hi.p);
    const int* p2f_col = static_cast<const int*>(plan->g_p2f_col.p);
    const C* wpos = static_cast<const C*>(plan->g_wpos.p);
    const int* f2p_row = static_cast<const int*>(plan->g_f2p_row.p);
    auto col_fwd = [&](auto KK) -> int {
        using K = decltype(KK);
        // two elements per load where the pairs' bases allow it
        const size_t al = 2 * sizeof(InT);
        const int vec_src = reinterpret_cast<uintptr_t>(src) % al == 0 && sp % 2 == 0;
        const int vec_smp = reinterpret_cast<uintptr_t>(smp) % al == 0 && mp % 2 == 0;
        typename K::Params p{static_cast<const InT*>(src), static_cast<const InT*>(smp), planes, peaks, wcol, lo, hi,
                             p2f_col, sh, sp, mp, vec_src, vec_smp};
        const dim3 grid((sh.M2 + K::CT - 1) / K::CT, 2, pairs);
        return launch(ctx, d, KC_COL_FWD, st, [&] {
            launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
        });
    };
    if constexpr (sizeof(T) == 4) {
        if ((sh.ct == 16 ? col_fwd(GenColFwdKernel<T, InT, 16>{}) : col_fwd(GenColFwdKernel<T, InT, 8>{})) != 0) return -1;
    } else {
        if (col_fwd(GenColFwdKernel<T, InT, 8>{}) != 0) return -1;
    }
    {
        using K = GenRowFusedKernel<T>;
        typename K::Params p{planes, wrow, wpos, lo, hi, f2p_row, sh};
        const dim3 grid(sh.M1 / 2 + 1, 1, pairs);
        if (launch(ctx, d, KC_ROW_FUSED, st, [&] {
                launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
            }) != 0) return -1;
    }
    auto col_inv = [&](auto KK) -> int {
        using K = decltype(KK);
        typename K::Params p{planes, peaks, wcol, p2f_col, sh, reinterpret_cast<T*>(planes)};
        const dim3 grid(pairs, (sh.M2 + K::CT - 1) / K::CT, 1);
        return launch(ctx, d, KC_COL_INV, st, [&] {
            launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
        });
    };
    if constexpr (sizeof(T) == 4) {
        if ((sh.ct == 16 ? col_inv(GenColInvKernel<T, 16>{}) : col_inv(GenColInvKernel<T, 8>{})) != 0) return -1;
    } else {
        if (col_inv(GenColInvKernel<T, 8>{}) != 0) return -1;
    }
    if constexpr (sizeof(T) == 8) {
        // fp64: r[0 .. 2L) of pair i sits in its (dead) sample plane; full double keys
        const double* r = reinterpret_cast<const double*>(planes) + 2 * sh.M;
        if (launch(ctx, d, KC_ARGMAX_F64, st, [&] {
                argmax_f64_kernel<<<pairs, 1024, 0, st>>>(r, 2 * sh.L, 4 * sh.M, peaks);
            }) != 0) return -1;
    }
    return 0;
}

template <typename T>
int build_generic_plan_t(FftPlan* plan) {
    typedef typename GenTraits<T>::C C;
    const GenShape& sh = plan->gen;
    const GenTables<C> tb = gen_build_tables<C>(sh);
    if (upload(plan->g_wcol, tb.wcol) != 0 || upload(plan->g_wrow, tb.wrow) != 0 || upload(plan->g_lo, tb.m_lo) != 0 ||
        upload(plan->g_hi, tb.m_hi) != 0 || upload(plan->g_p2f_col, tb.p2f_col) != 0 ||
        upload(plan->g_wpos, tb.wpos) != 0 || upload(plan->g_f2p_row, tb.f2p_row) != 0)
        return -1;
    if (prepare_gen_kernel<GenRowFusedKernel<T>>(GenRowFusedKernel<T>::smem_bytes(sh)) != 0) return -1;
    auto prep_cols = [&](auto CTC) -> int {
        constexpr int CT = decltype(CTC)::value;
        return prepare_gen_kernel<GenColFwdKernel<T, float, CT>>(GenColFwdKernel<T, float, CT>::smem_bytes(sh)) != 0 ||
               prepare_gen_kernel<GenColFwdKernel<T, double, CT>>(GenColFwdKernel<T, double, CT>::smem_bytes(sh)) != 0 ||
               prepare_gen_kernel<GenColInvKernel<T, CT>>(GenColInvKernel<T, CT>::smem_bytes(sh)) != 0 ? -1 : 0;
    };
    if (sizeof(T) == 4 && sh.ct == 16) { if (prep_cols(IC<16>{}) != 0) return -1; }
    else if (prep_cols(IC<8>{}) != 0) return -1;
    plan->run_wave = [plan](audiosync_cuda_ctx* ctx, DeviceState& d, const void* src, const void* smp, int dtype,
                            long long sp, long long mp, void* ws, PairPeak* peaks, int pairs, cudaStream_t st) {
        return dtype == AUDIOSYNC_CUDA_F32 ? run_generic_wave<T, float>(plan, ctx, d, src, smp, sp, mp, ws, peaks, pairs, st)
                                           : run_generic_wave<T, double>(plan, ctx, d, src, smp, sp, mp, ws, peaks, pairs, st);
    };
    return 0;
}


}  // namespace asc
