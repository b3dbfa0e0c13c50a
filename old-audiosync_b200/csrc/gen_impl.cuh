// gen_impl.cuh -- launchers of the runtime-radix kernels, one set per (arithmetic type, CTA size);
// included by gen_f32_*.cu / gen_f64_*.cu, which instantiate GenStage<T, NT> (declared in plan_host.h).
#pragma once

#include "plan_host.h"
#include "reduce_kernels.cuh"

namespace asc {

// The runtime-radix kernels serve plans of many shapes: each is allowed the device's whole opt-in
// shared memory (a per-plan limit would be lowered again by the next, smaller plan of the same kernel).
template <class K>
static int prepare_gen_kernel(size_t smem) {
    int dev = 0, optin = 0;
    ASC_CUDA_OK(cudaGetDevice(&dev));
    ASC_CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)optin) {
        set_last_error("plan needs %zu bytes of shared memory per CTA, the device allows %d", smem, optin);
        return -1;
    }
    ASC_CUDA_OK(cudaFuncSetAttribute(gen_kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    return 0;
}

// fp32 arithmetic has 16- and 8-column tiles, fp64 only 8
template <typename T, class F>
static int gen_for_tile_width(int ct, F&& f) {
    if constexpr (sizeof(T) == 4) {
        if (ct == 16) return f(IC<16>{});
    }
    return f(IC<8>{});
}

template <typename T, int NT>
int GenStage<T, NT>::prepare_cols(const GenShape& sh) {
    return gen_for_tile_width<T>(sh.ct, [&](auto CTC) -> int {
        constexpr int CT = decltype(CTC)::value;
        return prepare_gen_kernel<GenColFwdKernel<T, float, CT, NT>>(GenColFwdKernel<T, float, CT, NT>::smem_bytes(sh)) != 0 ||
               prepare_gen_kernel<GenColFwdKernel<T, double, CT, NT>>(GenColFwdKernel<T, double, CT, NT>::smem_bytes(sh)) != 0 ||
               prepare_gen_kernel<GenColInvKernel<T, CT, NT>>(GenColInvKernel<T, CT, NT>::smem_bytes(sh)) != 0 ? -1 : 0;
    });
}

template <typename T, int NT>
int GenStage<T, NT>::prepare_rows(const GenShape& sh) {
    return prepare_gen_kernel<GenRowFusedKernel<T, NT>>(GenRowFusedKernel<T, NT>::smem_bytes(sh));
}

template <typename T, int NT>
int GenStage<T, NT>::col_fwd(FftPlan* plan, audiosync_cuda_ctx* ctx, DeviceState& d, const void* src, const void* smp,
                             int dtype, long long sp, long long mp, void* ws, PairPeak* peaks, int pairs, cudaStream_t st) {
    typedef typename GenTraits<T>::C C;
    const GenShape& sh = plan->gen;
    auto go = [&](auto IN) -> int {
        using InT = decltype(IN);
        return gen_for_tile_width<T>(sh.ct, [&](auto CTC) -> int {
            using K = GenColFwdKernel<T, InT, decltype(CTC)::value, NT>;
            // two elements per load where the pairs' bases allow it
            const size_t al = 2 * sizeof(InT);
            const int vec_src = reinterpret_cast<uintptr_t>(src) % al == 0 && sp % 2 == 0;
            const int vec_smp = reinterpret_cast<uintptr_t>(smp) % al == 0 && mp % 2 == 0;
            typename K::Params p{static_cast<const InT*>(src), static_cast<const InT*>(smp), static_cast<C*>(ws), peaks,
                                 static_cast<const C*>(plan->g_wcol.p), static_cast<const C*>(plan->g_lo.p),
                                 static_cast<const C*>(plan->g_hi.p), static_cast<const int*>(plan->g_p2f_col.p),
                                 sh, sp, mp, vec_src, vec_smp};
            const dim3 grid((sh.M2 + K::CT - 1) / K::CT, 2, pairs);
            return launch(ctx, d, KC_COL_FWD, st, [&] {
                launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
            });
        });
    };
    return dtype == AUDIOSYNC_CUDA_F32 ? go(float{}) : go(double{});
}

template <typename T, int NT>
int GenStage<T, NT>::rows(FftPlan* plan, audiosync_cuda_ctx* ctx, DeviceState& d, void* ws, int pairs, cudaStream_t st) {
    typedef typename GenTraits<T>::C C;
    const GenShape& sh = plan->gen;
    using K = GenRowFusedKernel<T, NT>;
    typename K::Params p{static_cast<C*>(ws), static_cast<const C*>(plan->g_wrow.p), static_cast<const C*>(plan->g_wpos.p),
                         static_cast<const C*>(plan->g_lo.p), static_cast<const C*>(plan->g_hi.p),
                         static_cast<const int*>(plan->g_f2p_row.p), sh};
    const dim3 grid(sh.M1 / 2 + 1, 1, pairs);
    return launch(ctx, d, KC_ROW_FUSED, st, [&] {
        launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
    });
}

template <typename T, int NT>
int GenStage<T, NT>::col_inv(FftPlan* plan, audiosync_cuda_ctx* ctx, DeviceState& d, void* ws, PairPeak* peaks, int pairs,
                             cudaStream_t st) {
    typedef typename GenTraits<T>::C C;
    const GenShape& sh = plan->gen;
    C* planes = static_cast<C*>(ws);
    const int rc = gen_for_tile_width<T>(sh.ct, [&](auto CTC) -> int {
        using K = GenColInvKernel<T, decltype(CTC)::value, NT>;
        typename K::Params p{planes, peaks, static_cast<const C*>(plan->g_wcol.p), static_cast<const int*>(plan->g_p2f_col.p),
                             sh, reinterpret_cast<T*>(planes)};
        const dim3 grid(pairs, (sh.M2 + K::CT - 1) / K::CT, 1);
        return launch(ctx, d, KC_COL_INV, st, [&] {
            launch_stage(gen_kernel_entry<K>, grid, dim3(K::THREADS), K::smem_bytes(sh), st, p);
        });
    });
    if (rc != 0) return -1;
    if constexpr (sizeof(T) == 8) {
        // fp64: r[0 .. 2L) of pair i sits in its (dead) sample plane; full double keys
        const double* r = reinterpret_cast<const double*>(planes) + 2 * sh.M;
        return launch(ctx, d, KC_ARGMAX_F64, st, [&] {
            argmax_f64_kernel<<<pairs, 1024, 0, st>>>(r, 2 * sh.L, 4 * sh.M, peaks);
        });
    }
    return 0;
}

}  // namespace asc
