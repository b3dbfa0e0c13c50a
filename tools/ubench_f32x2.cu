// Microbenchmark: scalar FADD/FFMA vs packed FADD2/FFMA2 issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 c; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float addf(float a, float b){ float c; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(c) : "f"(a), "f"(b)); return c; }
__device__ __forceinline__ float fmaf_(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
constexpr int CH = 8;      // independent chains per thread
constexpr int IT = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed) {
    float r = 0.f;
    if (MODE == 0) {         // FADD scalar: 2*CH chains (same number of float results as packed)
        float v[2 * CH]; for (int i = 0; i < 2 * CH; i++) v[i] = seed + i + threadIdx.x;
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < 2 * CH; i++) v[i] = addf(v[i], seed);
        for (int i = 0; i < 2 * CH; i++) r += v[i];
    } else if (MODE == 1) {  // FADD2
        u64 v[CH]; u64 s; float2 sf = make_float2(seed, seed); s = *reinterpret_cast<u64*>(&sf);
        for (int i = 0; i < CH; i++) { float2 f = make_float2(seed + i, seed + threadIdx.x); v[i] = *reinterpret_cast<u64*>(&f); }
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < CH; i++) v[i] = add2(v[i], s);
        for (int i = 0; i < CH; i++) { float2 f = *reinterpret_cast<float2*>(&v[i]); r += f.x + f.y; }
    } else if (MODE == 2) {  // FFMA scalar
        float v[2 * CH]; for (int i = 0; i < 2 * CH; i++) v[i] = seed + i + threadIdx.x;
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < 2 * CH; i++) v[i] = fmaf_(v[i], seed, seed);
        for (int i = 0; i < 2 * CH; i++) r += v[i];
    } else if (MODE == 3) {  // FFMA2
        u64 v[CH]; u64 s; float2 sf = make_float2(seed, seed); s = *reinterpret_cast<u64*>(&sf);
        for (int i = 0; i < CH; i++) { float2 f = make_float2(seed + i, seed + threadIdx.x); v[i] = *reinterpret_cast<u64*>(&f); }
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < CH; i++) v[i] = fma2(v[i], s, s);
        for (int i = 0; i < CH; i++) { float2 f = *reinterpret_cast<float2*>(&v[i]); r += f.x + f.y; }
    } else if (MODE == 4) {  // mix: FADD2 + IADD (alu pipe) interleaved, to see co-issue
        u64 v[CH]; u64 s; float2 sf = make_float2(seed, seed); s = *reinterpret_cast<u64*>(&sf);
        int w[CH];
        for (int i = 0; i < CH; i++) { float2 f = make_float2(seed + i, seed + threadIdx.x); v[i] = *reinterpret_cast<u64*>(&f); w[i] = i + threadIdx.x; }
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < CH; i++) { v[i] = add2(v[i], s); asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[i]) : "r"(it)); }
        for (int i = 0; i < CH; i++) { float2 f = *reinterpret_cast<float2*>(&v[i]); r += f.x + f.y + w[i]; }
    } else if (MODE == 5) {  // mix: FADD scalar x2 + xor
        float v[2 * CH]; int w[CH]; for (int i = 0; i < 2 * CH; i++) v[i] = seed + i + threadIdx.x;
        for (int i = 0; i < CH; i++) w[i] = i + threadIdx.x;
        for (int it = 0; it < IT; it++)
#pragma unroll
            for (int i = 0; i < CH; i++) { v[2*i] = addf(v[2*i], seed); v[2*i+1] = addf(v[2*i+1], seed); asm volatile("xor.b32 %0, %0, %1;" : "+r"(w[i]) : "r"(it)); }
        for (int i = 0; i < 2 * CH; i++) r += v[i];
        for (int i = 0; i < CH; i++) r += w[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 8;
    bench<MODE><<<blocks, 256>>>(d, 1.0f);
    cudaEventRecord(a);
    bench<MODE><<<blocks, 256>>>(d, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double flt = (double)blocks * 256 * IT * 2 * CH;   // float results produced
    printf("%-28s %8.3f ms  %8.2f T float-results/s  (per SM per clk @1.965GHz: %.1f)\n", name, ms, flt / ms / 1e9,
           flt / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FADD scalar", d); run<1>("FADD2 packed", d); run<2>("FFMA scalar", d); run<3>("FFMA2 packed", d);
    run<4>("FADD2 + XOR", d); run<5>("2xFADD + XOR", d);
    return 0;
}
