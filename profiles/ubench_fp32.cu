// FP32 pipe microbenchmarks for the search-kernel design (run on the GPU box, results in profiles/r01_ubench_fp32.txt):
//   1. FFMA, register operands, 16 independent chains                        -> scalar FP32 peak
//   2. FFMA2 (fma.rn.f32x2), 8 independent 64-bit chains                    -> packed FP32 peak
//   3. the same with ALU-pipe work (FMNMX) interleaved 1:8 and 1:4          -> does packed FMA free issue slots?
//   4. the same with broadcast LDS.128 interleaved                          -> shared-memory operand feed
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp32 ubench_fp32.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, int iters, float x, float y) {
    __shared__ float4 sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = make_float4(x, y, x, y);
    __syncthreads();
    if (MODE == 0) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = (float)(threadIdx.x + j) * 1e-3f;
        float xr = x + threadIdx.x * 1e-9f, yr = y + threadIdx.x * 1e-9f;   // registers, not constants
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], xr, yr);
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += a[j];
        if (s == 123.456f) sink[0] = s;
    } else {
        unsigned long long a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = pack((float)(threadIdx.x + j) * 1e-3f, (float)(threadIdx.x + j) * 2e-3f);
        unsigned long long xr = pack(x + threadIdx.x * 1e-9f, x), yr = pack(y + threadIdx.x * 1e-9f, y);
        float m = 1e30f;
        float xs = x + threadIdx.x * 1e-9f, ys = y + threadIdx.x * 1e-9f;
        float4 acc4 = make_float4(0, 0, 0, 0);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == 6 || MODE == 7) a[j] = fma2(a[j], pack(xs, xs), pack(ys, ys));     // scalar operands broadcast by the instruction
                else a[j] = fma2(a[j], xr, yr);
                if (MODE == 7 && (j & 1)) { float lo, hi; unpack(a[j], lo, hi); m = fminf(m, hi); if (j == 7) { float4 t = sh[(i + threadIdx.x / 32) & 63]; xs = t.x; ys = t.y; } }
                if (MODE == 2 && j == 7) { float lo, hi; unpack(a[j], lo, hi); m = fminf(m, lo); }
                if (MODE == 3 && (j & 1)) { float lo, hi; unpack(a[j], lo, hi); m = fminf(m, hi); }
                if (MODE == 4 && j == 7) { float4 t = sh[(i + threadIdx.x / 32) & 63]; acc4.x += 0.f; xr = pack(t.x, t.z); yr = pack(t.y, t.w); }
                if (MODE == 5 && (j & 1)) { float lo, hi; unpack(a[j], lo, hi); m = fminf(m, hi); if (j == 7) { float4 t = sh[(i + threadIdx.x / 32) & 63]; xr = pack(t.x, t.z); yr = pack(t.y, t.w); } }
            }
        }
        float s = m + acc4.x;
#pragma unroll
        for (int j = 0; j < 8; ++j) { float lo, hi; unpack(a[j], lo, hi); s += lo + hi; }
        if (s == 123.456f) sink[0] = s;
    }
}

template <int MODE> static void run(const char* name, float* sink, int sms, double clk_ghz) {
    const int iters = 1 << 14, grid = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<grid, 256>>>(sink, iters, 0.999f, 1e-3f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    printf("%-60s %7.2f TFLOP/s  (%.1f%% of %d SMs x 128 x 2 x %.3f GHz)\n", name, best, best / (sms * 128 * 2 * clk_ghz * 1e-3) * 100, sms, clk_ghz);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* sink;
    cudaMalloc(&sink, 256);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, max SM clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<0>("FFMA   3-register form, 16 chains", sink, p.multiProcessorCount, ghz);
    run<1>("FFMA2  fma.rn.f32x2, 8 x 64-bit chains", sink, p.multiProcessorCount, ghz);
    run<2>("FFMA2  + 1 FMNMX per 8 FFMA2", sink, p.multiProcessorCount, ghz);
    run<3>("FFMA2  + 1 FMNMX per 2 FFMA2", sink, p.multiProcessorCount, ghz);
    run<4>("FFMA2  + 1 LDS.128 per 8 FFMA2 (operands from smem)", sink, p.multiProcessorCount, ghz);
    run<5>("FFMA2  + 4 FMNMX + 1 LDS.128 per 8 FFMA2", sink, p.multiProcessorCount, ghz);
    run<6>("FFMA2  scalar .F32 broadcast operands (d.xy * a + b)", sink, p.multiProcessorCount, ghz);
    run<7>("FFMA2  scalar operands + 4 FMNMX + 1 LDS.128 per 8 FFMA2", sink, p.multiProcessorCount, ghz);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
