// Operand-pattern test for FFMA2 at 4 warps/SMSP, pure registers, no memory:
//  A: x = fma2(x, s, t) (scalar broadcast operands)                 16 chains
//  B: r = fma2(d, s, t); acc = fma2(r, r, acc)   (the chi^2 pattern)  NCH accumulator chains, 5 components
//  C: like B but the square-accumulate done with scalar FFMA on the halves
//  D: like B with acc = fma2(r, r2, acc) where r2 is a copy in another register (tests same-register double read)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE, int NCH>
__global__ void __launch_bounds__(256, 2) k(float* sink, int iters, float x, float y) {
    float s[5], t[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) { s[c] = x + threadIdx.x * 1e-9f * (c + 1); t[c] = y + threadIdx.x * 1e-9f * (c + 2); }
    u64 d[NCH][5], acc[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        acc[j] = pack(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 5; ++c) d[j][c] = pack((threadIdx.x + j + c) * 1e-3f, (threadIdx.x + j * 3 + c) * 2e-3f);
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (MODE == 0) { d[j][c] = fma2(d[j][c], pack(s[c], s[c]), pack(t[c], t[c])); acc[j] = fma2(acc[j], pack(s[c], s[c]), pack(t[c], t[c])); }
                else {
                    const u64 r = fma2(d[j][c], pack(s[c], s[c]), pack(t[c], t[c]));
                    if (MODE == 1) acc[j] = fma2(r, r, acc[j]);
                    if (MODE == 2) { float lo, hi, al, ah; unpack(r, lo, hi); unpack(acc[j], al, ah); al = fmaf(lo, lo, al); ah = fmaf(hi, hi, ah); acc[j] = pack(al, ah); }
                    if (MODE == 3) acc[j] = fma2(r, d[j][c], acc[j]);
                }
            }
#pragma unroll
        for (int c = 0; c < 5; ++c) { s[c] += 1e-9f; }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) { float lo, hi; unpack(acc[j], lo, hi); sum += lo + hi; if (MODE == 0) { for (int c = 0; c < 5; ++c) { unpack(d[j][c], lo, hi); sum += lo + hi; } } }
    if (sum == 123.456f) sink[0] = sum;
}
template <int MODE, int NCH> static void run(const char* name, float* sink, int sms, double ghz, int ctas) {
    const int iters = 1 << 12;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a); k<MODE, NCH><<<sms * ctas, 256>>>(sink, iters, 0.999f, 1e-3f); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep > 0 && ms < best) best = ms;
    }
    const double warps = 2.0 * ctas;     // per SMSP
    const double fma_cyc = warps * iters * 5.0 * NCH * 2 * (MODE == 2 ? 2.0 : 2.0);   // two FMA-pipe ops per (chain, component), 2 cycles each
    printf("%-58s NCH %d, %.0f warps/SMSP: FMA pipe %5.1f%% busy\n", name, NCH, warps, fma_cyc / (best * 1e-3 * ghz * 1e9) * 100);
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* sink; cudaMalloc(&sink, 256);
    const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
    run<0, 4>("A  x = fma2(x, s, t) everywhere", sink, sms, ghz, 2);
    run<1, 4>("B  r = fma2(d,s,t); acc = fma2(r,r,acc)", sink, sms, ghz, 2);
    run<1, 8>("B  r = fma2(d,s,t); acc = fma2(r,r,acc)", sink, sms, ghz, 2);
    run<1, 4>("B  (1 CTA/SM)", sink, sms, ghz, 1);
    run<2, 4>("C  r = fma2(d,s,t); acc.lo/hi by scalar FFMA", sink, sms, ghz, 2);
    run<3, 4>("D  r = fma2(d,s,t); acc = fma2(r,d,acc)", sink, sms, ghz, 2);
    run<3, 8>("D  r = fma2(d,s,t); acc = fma2(r,d,acc)", sink, sms, ghz, 2);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
