// pipeline.cuh -- the wave pipeline kernel: one launch runs the four stages of
// the path on four CONSECUTIVE waves of pairs at once,
//
//     launch k:   K_A (wave k)   K_B (wave k-1)   K_C (wave k-2)   K_P (wave k-3)
//
// as CTA roles interleaved over the grid.  The stages of one launch touch
// different waves, so there is no dependency inside a launch; the dependency
// between the stages of one wave is the launch boundary (stream order).
//
// Why: the stand-alone kernels are bounded by different units -- K_A, K_C and
// K_P by HBM, K_B by the fp32 pipe and shared memory -- and each leaves the
// others' units idle.  With the roles interleaved every SM holds a mix of
// HBM-bound and arithmetic-bound CTAs, so the column tiles stream from HBM
// while the row CTAs compute.  It also cuts the launches per wave from 4 to 1.
//
// The role bodies are the same K::run functions the stand-alone kernels (and
// the CPU emulator) execute; only the block coordinates come from a schedule
// table instead of blockIdx.
#pragma once

#include <vector>

#include "fft_kernels.cuh"
#include "reduce_kernels.cuh"

namespace asc {

enum PipeRole { ROLE_A = 0, ROLE_B = 1, ROLE_C = 2, ROLE_P = 3, ROLE_COUNT = 4 };

// Pearson arguments of the pipeline's K_P role (pitches are 2L / L).
template <typename InT>
struct PearsonArgs {
    const InT* sources;
    const InT* samples;
    long long L;
    const PairPeak* peaks;
    PearsonPartial* partials;
    unsigned int* tickets;
    int n_chunks;
    audiosync_cuda_result* results;
};

template <class KA, class KB, class KC, typename InT>
struct PipelineParams {
    typename KA::Params a;
    typename KB::Params b;
    typename KC::Params c;
    PearsonArgs<InT> p;
    int pairs[ROLE_COUNT];       // pairs of the wave each role works on (0: role idle in this launch)
    const uint16_t* sched;       // [period]: role << 12 | item, roles interleaved in proportion
    int period;                  // work items of one pair over all roles
};

// Items per pair of each role for a (M1, M2, L, NT) plan.
struct PipeShape {
    int items[ROLE_COUNT];
    int period() const { return items[0] + items[1] + items[2] + items[3]; }
};

inline PipeShape pipe_shape(int M1, int M2, long long L, int nt) {
    PipeShape s;
    s.items[ROLE_A] = 2 * (M2 / COL_T);
    s.items[ROLE_B] = M1 / 2 + 1;
    s.items[ROLE_C] = M2 / COL_T;
    s.items[ROLE_P] = (int)((L + (long long)nt * PEARSON_PER_THREAD - 1) / ((long long)nt * PEARSON_PER_THREAD));
    return s;
}

// Proportional interleave: item k of a role with c items sits at time (k + 1/2) / c.
inline std::vector<uint16_t> build_pipe_schedule(const PipeShape& s) {
    struct E { double t; int role, item; };
    std::vector<E> e;
    for (int r = 0; r < ROLE_COUNT; r++)
        for (int k = 0; k < s.items[r]; k++) e.push_back({(k + 0.5) / s.items[r], r, k});
    // stable order: time, then role
    for (size_t i = 1; i < e.size(); i++) {
        E v = e[i];
        size_t j = i;
        while (j > 0 && (e[j - 1].t > v.t || (e[j - 1].t == v.t && e[j - 1].role > v.role))) { e[j] = e[j - 1]; j--; }
        e[j] = v;
    }
    std::vector<uint16_t> out;
    for (const E& x : e) out.push_back((uint16_t)((x.role << 12) | x.item));
    return out;
}

#if defined(__CUDACC__)
// grid = (period * max(pairs[]), 1, 1); dynamic shared memory = the largest role's.
template <class KA, class KB, class KC, typename InT, int NT>
__global__ void __launch_bounds__(NT, 3) pipeline_entry(const PipelineParams<KA, KB, KC, InT> q) {
    static_assert(KA::THREADS == NT && KB::THREADS == NT && KC::THREADS == NT, "one block size for every role");
    extern __shared__ __align__(128) unsigned char asc_smem[];   // one declaration of the dynamic buffer for every entry
    const unsigned b = blockIdx.x;
    const unsigned g = b / (unsigned)q.period;          // pair of the wave
    const unsigned i = b - g * (unsigned)q.period;
    const unsigned code = __ldg(q.sched + i);
    const unsigned role = code >> 12, item = code & 0xfffu;
    DeviceExec ex;
    cplx* buf = reinterpret_cast<cplx*>(asc_smem);
    if (role == ROLE_A) {
        if ((int)g >= q.pairs[ROLE_A]) return;
        constexpr unsigned tiles = KA::M2 / COL_T;
        ex.x_ = (int)(item % tiles); ex.y_ = (int)(item / tiles); ex.z_ = (int)g;
        KA::run(ex, q.a, buf);
    } else if (role == ROLE_B) {
        if ((int)g >= q.pairs[ROLE_B]) return;
        ex.x_ = (int)item; ex.y_ = 0; ex.z_ = (int)g;
        KB::run(ex, q.b, buf);
    } else if (role == ROLE_C) {
        if ((int)g >= q.pairs[ROLE_C]) return;
        // pair index fastest, as in the stand-alone launch (the argmax threshold of a pair is
        // warm for its later tiles)
        constexpr unsigned tiles = KC::M2 / COL_T;
        const unsigned j = g * tiles + item;
        ex.x_ = (int)(j % (unsigned)q.pairs[ROLE_C]); ex.y_ = (int)(j / (unsigned)q.pairs[ROLE_C]); ex.z_ = 0;
        KC::run(ex, q.c, buf);
    } else {
        if ((int)g >= q.pairs[ROLE_P]) return;
        pearson_block<InT, NT>(q.p.sources, q.p.samples, 2 * q.p.L, q.p.L, q.p.L, q.p.peaks, 0, q.p.partials,
                               q.p.tickets, q.p.n_chunks, q.p.results, (int)g, (int)item, (int)threadIdx.x,
                               *reinterpret_cast<PearsonShared<NT>*>(asc_smem));
    }
}
#endif

}  // namespace asc
