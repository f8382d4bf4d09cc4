// What does one extra non-FMA instruction cost next to a stream of packed FP32 FMAs (FFMA2) on sm_100a?
// Loop body: 16 independent FFMA2 (32 FMA-pipe cycles at 1 FFMA2 / 2 clk / SMSP) + NX instructions of type OP.
// Reports SMSP cycles per loop body (16 warps per SM = 4 per SMSP, as the search kernel runs) and the marginal
// cost of one OP in cycles.  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_coissue ubench_coissue.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

enum { OP_NONE, OP_FMNMX, OP_FMNMX3, OP_LOP3, OP_IADD3, OP_IMAD, OP_FSETP_SEL, OP_LDS128, OP_FADD, OP_FMUL, OP_HMNMX2, OP_VIMNMX, OP_SHF, OP_FFMA, OP_LDS32, OP_PRMT, OP_FSETP };

template <int OP, int NX>
__global__ void __launch_bounds__(256, 2) k(float* sink, int iters, float x, float y) {
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(x, y, x + threadIdx.x, y);
    __syncthreads();
    u64 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = pack((float)(threadIdx.x + j) * 1e-3f, (float)(threadIdx.x + j) * 2e-3f);
    float xs = x + threadIdx.x * 1e-9f, ys = y + threadIdx.x * 1e-9f, xs2 = xs, ys2 = ys;
    float f[4] = {1e30f, 2e30f, 3e30f, 4e30f};
    unsigned u[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
    const uint32_t shb = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 31) * 16;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            a[j] = (j & 1) ? fma2(a[j], pack(xs2, xs2), pack(ys2, ys2)) : fma2(a[j], pack(xs, xs), pack(ys, ys));
            // spread the NX extra instructions evenly; each reads a fresh FMA result so that it cannot be hoisted
            if ((j * NX) / 16 != ((j + 1) * NX) / 16) {
                const int q = ((j * NX) / 16) & 3;
                float lo, hi;
                unpack(a[j], lo, hi);
                if (OP == OP_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(hi));
                if (OP == OP_FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[q]) : "f"(hi), "f"(lo));
                if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xfe;" : "+r"(u[q]) : "r"(__float_as_uint(hi)), "r"(__float_as_uint(lo)));
                if (OP == OP_IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[q]) : "r"(__float_as_uint(hi)));
                if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(u[q]) : "r"(__float_as_uint(hi)), "r"(__float_as_uint(lo)));
                if (OP == OP_FSETP_SEL) asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; selp.u32 %0, 1, %0, p;}" : "+r"(u[q]) : "f"(hi), "f"(lo));
                if (OP == OP_FSETP) asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; @p bra.uni L%=; L%=: }" :: "r"(u[q]), "f"(hi), "f"(lo));
                if (OP == OP_LDS128) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(shb + ((i + j) & 7) * 512)); xs = t.x; ys = t.y; xs2 = t.z; ys2 = t.w; }
                if (OP == OP_LDS32) { float t; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(shb + ((i + j) & 7) * 512)); f[q] = t; }
                if (OP == OP_FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(hi));
                if (OP == OP_FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(hi));
                if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[q]) : "f"(hi), "f"(lo));
                if (OP == OP_HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u[q]) : "r"(__float_as_uint(hi)));
                if (OP == OP_VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(u[q]) : "r"(__float_as_uint(hi)));
                if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(u[q]) : "r"(__float_as_uint(hi)));
                if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[q]) : "r"(__float_as_uint(hi)));
            }
        }
    }
    float s = xs2 + ys2 + f[0] + f[1] + f[2] + f[3] + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
#pragma unroll
    for (int j = 0; j < 16; ++j) { float lo, hi; unpack(a[j], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}

static double g_base = 0;
template <int OP, int NX> static void run(const char* name, float* sink, int sms, double ghz) {
    const int iters = 1 << 13, grid = sms * 2;     // 2 CTAs x 8 warps per SM = 4 warps per SMSP
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<OP, NX><<<grid, 256>>>(sink, iters, 0.999f, 1e-3f);
        cudaEventRecord(b);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    // cycles per loop body per SMSP: 4 warps each run `iters` bodies
    const double cyc = best * 1e-3 * ghz * 1e9 / (4.0 * iters);
    if (OP == OP_NONE) g_base = cyc;
    printf("%-34s NX=%2d  %7.2f cyc/body  (+%5.2f)  %5.2f cyc per extra instruction\n", name, NX, cyc, cyc - g_base, NX ? (cyc - g_base) / NX : 0.0);
}
#define RUN3(OP, NAME) run<OP, 4>(NAME, sink, sms, ghz); run<OP, 8>(NAME, sink, sms, ghz); run<OP, 16>(NAME, sink, sms, ghz);

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* sink;
    cudaMalloc(&sink, 256);
    const double ghz = khz * 1e-6;
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz; body = 16 FFMA2 (ideal 32 cyc/SMSP) + NX x OP, 4 warps/SMSP\n", p.name, sms, ghz);
    run<OP_NONE, 0>("none", sink, sms, ghz);
    RUN3(OP_FMNMX, "FMNMX   min.f32 (2 in)")
    RUN3(OP_FMNMX3, "FMNMX3  min.f32 (3 in)")
    RUN3(OP_LOP3, "LOP3")
    RUN3(OP_IADD3, "IADD3")
    RUN3(OP_IMAD, "IMAD")
    RUN3(OP_FSETP_SEL, "FSETP+SEL (2 instr)")
    RUN3(OP_LDS128, "LDS.128")
    RUN3(OP_LDS32, "LDS.32")
    RUN3(OP_FADD, "FADD")
    RUN3(OP_FMUL, "FMUL")
    RUN3(OP_FFMA, "FFMA scalar")
    RUN3(OP_HMNMX2, "HMNMX2  min.f16x2")
    RUN3(OP_VIMNMX, "VIMNMX  min.u32")
    RUN3(OP_SHF, "SHF")
    RUN3(OP_PRMT, "PRMT")
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
