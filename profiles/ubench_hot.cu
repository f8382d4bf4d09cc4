// The search kernel's hot loop in isolation: per (tile, pixel) 3 broadcast LDS.128 of the pixel record, 5*EPL/2 * 2 FFMA2,
// sign-bit OR, one funnel shift.  No TMA, no rare pass, no bootstrap: what is the best the FP32 pipe does on this
// instruction mix, as a function of warps per SM, unrolling and entries per lane?
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_hot ubench_hot.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

constexpr int NPX = 32;

template <int EPL, int UNROLL, int MODE>
__global__ void k(uint32_t* out, int ntiles, float seed) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* recs = reinterpret_cast<float*>(smem) + warp * (NPX + 4) * 12;
    for (int i = lane; i < (NPX + 4) * 12; i += 32) recs[i] = seed * (float)(i % 7 + 1);
    __syncwarp();
    float d[EPL][5];
#pragma unroll
    for (int e = 0; e < EPL; ++e)
#pragma unroll
        for (int c = 0; c < 5; ++c) d[e][c] = seed * (float)(lane + e * 5 + c);
    const uint32_t rec_base = (uint32_t)__cvta_generic_to_shared(recs);
    uint32_t total = 0;
    for (int t = 0; t < ntiles; ++t) {
        uint32_t hmask = 0;
        uint32_t ra = rec_base;
        float4 n0 = lds128(ra), n1 = lds128(ra + 16), n2 = lds128(ra + 32);
#pragma unroll UNROLL
        for (int j = 0; j < NPX; ++j) {
            float4 x0 = n0, x1 = n1, x2 = n2;
            if (MODE != 2 || (j & 3) == 3) {
                ra += 48;
                n0 = lds128(ra); n1 = lds128(ra + 16); n2 = lds128(ra + 32);
            }
            if (MODE == 2) {     // rotate the coefficients so that the four pixels of a record are different computations
                const int q = j & 3;
                if (q == 1) { x0 = make_float4(x0.y, x0.z, x0.w, x1.x); x1 = make_float4(x0.x, x1.z, x1.w, x2.x); }
                if (q == 2) { x0 = make_float4(x0.z, x0.w, x1.x, x0.x); x1 = make_float4(x0.y, x1.w, x2.x, x2.y); }
                if (q == 3) { x0 = make_float4(x0.w, x1.x, x0.x, x0.y); x1 = make_float4(x0.z, x2.x, x2.y, x1.y); }
            }
            const float a[5] = {x0.x, x0.y, x0.z, x0.w, x1.x};
            const float nb[5] = {x1.y, x1.z, x1.w, x2.x, x2.y};
            uint32_t sg = 0;
            float mn = 1e30f;
#pragma unroll
            for (int e = 0; e < EPL; e += 2) {
                u64 acc = pack(x2.z, x2.z);
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const u64 r = fma2(pack(d[e][c], d[e + 1][c]), pack(a[c], a[c]), pack(nb[c], nb[c]));
                    acc = fma2(r, r, acc);
                }
                float lo, hi;
                unpack(acc, lo, hi);
                if (MODE == 0 || MODE == 2) sg |= __float_as_uint(lo) | __float_as_uint(hi);
                else if (MODE == 3) { if (e == 0) sg = __float_as_uint(lo); if (e == EPL - 2) sg ^= __float_as_uint(hi); if (e > 0 && e < EPL - 2) asm volatile("" :: "f"(lo), "f"(hi)); }
                else mn = fminf(mn, fminf(lo, hi));
            }
            if (MODE == 0 || MODE == 2) hmask = __funnelshift_l(sg, hmask, 1);
            else if (MODE == 3) hmask ^= sg;
            else hmask = (hmask << 1) | (mn <= x2.w ? 1u : 0u);
        }
        total += __popc(hmask);
#pragma unroll
        for (int c = 0; c < 5; ++c) d[0][c] += 1e-7f;      // the next "tile" differs
    }
    if (total == 0xffffffffu) out[0] = total;
}

template <int EPL, int UNROLL, int MODE>
static void run(int warps_per_cta, int ctas_per_sm, int sms, double ghz, uint32_t* out) {
    const int ntiles = 400;
    const size_t smem = (size_t)warps_per_cta * (NPX + 4) * 48;
    const size_t pad = (size_t)(227 * 1024) / ctas_per_sm - 1024;       // force the residency with shared memory
    const size_t dyn = smem > pad ? smem : pad;
    cudaFuncSetAttribute(k<EPL, UNROLL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<EPL, UNROLL, MODE>, warps_per_cta * 32, dyn);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<EPL, UNROLL, MODE>);
    if (occ < ctas_per_sm) { printf("EPL %2d unroll %d mode %d  %2d warps x %d CTAs: does not fit (occ %d, %d regs)\n", EPL, UNROLL, MODE, warps_per_cta, ctas_per_sm, occ, fa.numRegs); return; }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<EPL, UNROLL, MODE><<<sms * ctas_per_sm, warps_per_cta * 32, dyn>>>(out, ntiles, 1e-3f);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    const double warps_per_smsp = warps_per_cta * ctas_per_sm / 4.0;
    const double fma_cycles = warps_per_smsp * ntiles * NPX * (5.0 * EPL / 2 * 2) * 2;     // FFMA2 at 2 cycles each
    const double cyc = best * 1e-3 * ghz * 1e9;
    printf("EPL %2d unroll %d %s  %2d warps/CTA x %d CTAs/SM (%4.1f warps/SMSP, %3d regs): FMA pipe %5.1f%% busy\n", EPL, UNROLL,
           MODE == 0 ? "signbit" : MODE == 1 ? "mintree" : MODE == 2 ? "lds/4  " : "no-or  ", warps_per_cta, ctas_per_sm, warps_per_smsp, fa.numRegs, fma_cycles / cyc * 100);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* out;
    cudaMalloc(&out, 256);
    const double ghz = khz * 1e-6;
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz\n", p.name, sms, ghz);
    run<8, 2, 0>(8, 2, sms, ghz, out);
    run<8, 2, 1>(8, 2, sms, ghz, out);
    run<8, 4, 2>(8, 2, sms, ghz, out);
    run<8, 2, 3>(8, 2, sms, ghz, out);
    run<8, 1, 0>(8, 2, sms, ghz, out);
    run<8, 4, 0>(8, 2, sms, ghz, out);
    run<8, 8, 0>(8, 2, sms, ghz, out);
    run<8, 2, 0>(4, 1, sms, ghz, out);
    run<8, 2, 0>(8, 1, sms, ghz, out);
    run<8, 2, 0>(12, 1, sms, ghz, out);
    run<8, 2, 0>(10, 2, sms, ghz, out);
    run<8, 2, 0>(12, 2, sms, ghz, out);
    run<8, 2, 0>(16, 2, sms, ghz, out);
    run<8, 4, 0>(12, 2, sms, ghz, out);
    run<16, 1, 0>(8, 1, sms, ghz, out);
    run<16, 2, 0>(8, 1, sms, ghz, out);
    run<16, 1, 0>(12, 1, sms, ghz, out);
    run<16, 2, 0>(12, 1, sms, ghz, out);
    run<16, 1, 0>(16, 1, sms, ghz, out);
    run<16, 2, 0>(8, 2, sms, ghz, out);
    run<16, 2, 0>(14, 1, sms, ghz, out);
    run<16, 1, 0>(14, 1, sms, ghz, out);
    run<16, 2, 0>(7, 2, sms, ghz, out);
    run<4, 2, 0>(8, 2, sms, ghz, out);
    run<4, 4, 0>(12, 2, sms, ghz, out);
    run<4, 4, 0>(16, 2, sms, ghz, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
