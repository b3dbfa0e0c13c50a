// static_plan_impl.cuh -- launches and plan construction of the static four-step kernels; included
// by the per-length translation units (static_*.cu), which instantiate build_static_plan<P>.
#pragma once

#include "plan_host.h"

namespace asc {

template <class P>
static int run_static_wave(FftPlan* plan, audiosync_cuda_ctx* ctx, DeviceState& d,
                           const void* src, const void* smp, int dtype, long long src_pitch,
                           long long smp_pitch, void* ws, PairPeak* peaks, int pairs, cudaStream_t st) {
    using Col = typename P::Col;
    using Row = typename P::Row;
    constexpr int M1 = Col::n, M2 = Row::n;
    cplx* planes = static_cast<cplx*>(ws);
    const cplx* col_tw = static_cast<const cplx*>(plan->col_tw.p);
    const cplx* row_tw = static_cast<const cplx*>(plan->row_tw.p);
    const cplx* col_tc = static_cast<const cplx*>(plan->col_tc.p);
    const cplx* row_rev = static_cast<const cplx*>(plan->row_rev.p);
    const cplx* row_tab = static_cast<const cplx*>(plan->row_tab.p);
    const cplx* m_lo = static_cast<const cplx*>(plan->m_lo.p);
    const cplx* m_hi = static_cast<const cplx*>(plan->m_hi.p);
    const dim3 grid_a(M2 / COL_T, 2, pairs);
    auto col_fwd = [&](auto KK, const auto* s_in, const auto* m_in) -> int {
        using K = decltype(KK);
        typename K::Params p{s_in, m_in, planes, peaks, col_tw, col_tc, m_lo, m_hi, P::L, src_pitch, smp_pitch};
        return launch(ctx, d, KC_COL_FWD, st, [&] {
            launch_stage(fft_kernel_entry<K>, grid_a, dim3(K::THREADS), K::SMEM, st, p);
        });
    };
    if (dtype == AUDIOSYNC_CUDA_F32) {
        // cp.async staging needs 16-byte aligned rows: every pair / row offset is a multiple
        // of 16 bytes, so only the base pointers decide.
        static const bool no_async = getenv("AUDIOSYNC_CUDA_NOASYNC") != nullptr;   // experiment knob
        const bool aligned = !no_async && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(smp)) & 15u) == 0 &&
                             src_pitch % 4 == 0 && smp_pitch % 4 == 0;
        const float* s_in = static_cast<const float*>(src);
        const float* m_in = static_cast<const float*>(smp);
        int rc;
        using KT = ColFwdKernel<Col, Row::n, P::NT_COL, float, 2>;
        typename KT::Params pt{s_in, m_in, planes, peaks, col_tw, col_tc, m_lo, m_hi, P::L, src_pitch, smp_pitch};
        if (aligned && tensor_map_encoder() &&
            make_tile_map(&pt.tm_src, s_in, M2, M1, (size_t)src_pitch * sizeof(float), (size_t)pairs,
                          tile_box_rows(KT::SRC_ROWS)) == 0 &&
            make_tile_map(&pt.tm_smp, m_in, M2, M1 / 2, (size_t)smp_pitch * sizeof(float), (size_t)pairs,
                          tile_box_rows(KT::SMP_ROWS)) == 0) {
            rc = launch(ctx, d, KC_COL_FWD, st, [&] {
                launch_stage(fft_kernel_entry<KT>, grid_a, dim3(KT::THREADS), KT::SMEM, st, pt);
            });
        } else {
            rc = aligned ? col_fwd(ColFwdKernel<Col, Row::n, P::NT_COL, float, 1>{}, s_in, m_in)
                         : col_fwd(ColFwdKernel<Col, Row::n, P::NT_COL, float, 0>{}, s_in, m_in);
        }
        if (rc != 0) return -1;
    } else {
        if (col_fwd(ColFwdKernel<Col, Row::n, P::NT_COL, double, 0>{}, static_cast<const double*>(src),
                    static_cast<const double*>(smp)) != 0) return -1;
    }
    {
        using K = RowFusedKernel<Row, Col::n, P::NT_ROW>;
        typename K::Params p{planes, row_tw, row_rev, m_lo, m_hi, P::L, row_tab};
        const dim3 grid(M1 / 2 + 1, 1, pairs);
        if (launch(ctx, d, KC_ROW_FUSED, st, [&] {
                launch_stage(fft_kernel_entry<K>, grid, dim3(K::THREADS), K::SMEM, st, p);
            }) != 0) return -1;
    }
    using KCT = ColInvKernel<Col, Row::n, P::NT_COL, true>;
    typename KCT::Params pct{planes, peaks, col_tw, P::L};
    if (tensor_map_encoder() &&
        make_tile_map(&pct.tm, planes, M2, M1, (size_t)P::L * sizeof(cplx), (size_t)2 * pairs,
                      tile_box_rows(KCT::TMA_ROWS)) == 0) {
        const dim3 grid(pairs, M2 / COL_T, 1);
        if (launch(ctx, d, KC_COL_INV, st, [&] {
                launch_stage(fft_kernel_entry<KCT>, grid, dim3(KCT::THREADS), KCT::SMEM, st, pct);
            }) != 0) return -1;
    } else {
        using K = ColInvKernel<Col, Row::n, P::NT_COL>;
        typename K::Params p{planes, peaks, col_tw, P::L};
        const dim3 grid(pairs, M2 / COL_T, 1);
        if (launch(ctx, d, KC_COL_INV, st, [&] {
                launch_stage(fft_kernel_entry<K>, grid, dim3(K::THREADS), K::SMEM, st, p);
            }) != 0) return -1;
    }
    return 0;
}


template <class P>
int build_static_plan(FftPlan* plan) {
    using Col = typename P::Col;
    using Row = typename P::Row;
    plan->kind = PATH_STATIC_FFT;
    plan->L = P::L;
    plan->M1 = Col::n;
    plan->M2 = Row::n;
    plan->ws_bytes_per_pair = (size_t)2 * P::L * sizeof(cplx);
    std::string d = "fft L=" + std::to_string(P::L) + " M1=" + std::to_string(Col::n) +
                    " M2=" + std::to_string(Row::n) + " col=";
    for (int i = 0; i < Col::count; i++) d += (i ? "x" : "") + std::to_string(Col::r(i));
    d += " row=";
    for (int i = 0; i < Row::count; i++) d += (i ? "x" : "") + std::to_string(Row::r(i));
    d += " static four-step fp32";
    plan->desc = d;
    std::vector<cplx> m_lo, m_hi;
    build_two_level(P::L, P::L - 1, m_lo, m_hi);
    if (upload(plan->col_tw, build_pass_tables(radix_vector<Col>())) != 0 ||
        upload(plan->row_tw, build_pass_tables(radix_vector<Row>())) != 0 ||
        upload(plan->col_tc, build_col_tc(P::L, Col::weight(Col::count - 1))) != 0 ||
        upload(plan->row_rev, build_row_rev<Row>()) != 0 ||
        upload(plan->row_tab, build_row_tab<Row>(P::L, Col::n)) != 0 ||
        upload(plan->m_lo, m_lo) != 0 || upload(plan->m_hi, m_hi) != 0)
        return -1;
    if (prepare_kernel<ColFwdKernel<Col, Row::n, P::NT_COL, float, 2>>(ColFwdKernel<Col, Row::n, P::NT_COL, float, 2>::SMEM) != 0 ||
        prepare_kernel<ColFwdKernel<Col, Row::n, P::NT_COL, float, 1>>(ColFwdKernel<Col, Row::n, P::NT_COL, float, 1>::SMEM) != 0 ||
        prepare_kernel<ColFwdKernel<Col, Row::n, P::NT_COL, float, 0>>(ColFwdKernel<Col, Row::n, P::NT_COL, float, 0>::SMEM) != 0 ||
        prepare_kernel<ColFwdKernel<Col, Row::n, P::NT_COL, double, 0>>(ColFwdKernel<Col, Row::n, P::NT_COL, double, 0>::SMEM) != 0 ||
        prepare_kernel<RowFusedKernel<Row, Col::n, P::NT_ROW>>(RowFusedKernel<Row, Col::n, P::NT_ROW>::SMEM) != 0 ||
        prepare_kernel<ColInvKernel<Col, Row::n, P::NT_COL, true>>(ColInvKernel<Col, Row::n, P::NT_COL, true>::SMEM) != 0 ||
        prepare_kernel<ColInvKernel<Col, Row::n, P::NT_COL>>(ColInvKernel<Col, Row::n, P::NT_COL>::SMEM) != 0)
        return -1;
    plan->run_wave = [plan](audiosync_cuda_ctx* ctx, DeviceState& d, const void* src, const void* smp,
                            int dtype, long long sp, long long mp, void* ws, PairPeak* peaks, int pairs,
                            cudaStream_t st) {
        return run_static_wave<P>(plan, ctx, d, src, smp, dtype, sp, mp, ws, peaks, pairs, st);
    };
    return 0;
}


}  // namespace asc
