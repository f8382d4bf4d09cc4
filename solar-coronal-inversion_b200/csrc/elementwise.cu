// The two elementwise, HBM-bandwidth-bound neighbours of the search:
//   k_blos      1-line analytic B_LOS, blos_proc_1pix of the reference (CLEDB_PROC.py:874-914)
//   k_qurotate  linear-polarisation frame rotation + height map, obs_qurotate + obs_calcheight
//               (CLEDB_PREPINV.py:1277-1350, 897-926)
// One thread per pixel, 128-bit coalesced loads/stores, grid sized from the pixel count.
#include <math_constants.h>

#include <algorithm>

#include "cledb_internal.h"

namespace cledb {

// 34 B/pixel: 16 in, 16 out, 2 flag bytes.  Four pixels per thread, 256 apart: four independent coalesced 16-byte loads
// in flight per thread before the first use.
//
// The kernel is only bandwidth-bound if a pixel costs few instructions (ncu, round 2: with IEEE divisions, IEEE square root
// and the library atan2f it ran at 172 instructions per pixel and 82 % issue-slot utilisation - bound by instruction issue at
// 0.70 of the HBM copy bandwidth).  So: the two quotients that DECIDE the integer issue flags (Q/I, U/I against 1e-6) keep the
// IEEE division the reference's float32 arithmetic has; the three field strengths and the azimuth are floating-point results
// with a stated tolerance (tests: 2e-6 relative, 1e-6 rad absolute) and use reciprocal / reciprocal-square-root
// approximations (2 ulp) and a degree-7 minimax polynomial for atan (3e-7 absolute, the accuracy of NumPy's own float32
// arctan2).

// atan2(y, x) for finite arguments: |error| <= 4e-7.  atan(t)/t on [0, 1] as a polynomial in t^2.
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = __fdividef(mn, mx);
    if (mn == mx) t = mx > 0.f ? 1.f : 0.f;                 // 0/0 (atan2(0, 0) = 0 or pi), inf/inf
    const float q = t * t;
    float p = -0.00405456f;
    p = fmaf(p, q, 0.02186293f);
    p = fmaf(p, q, -0.05591229f);
    p = fmaf(p, q, 0.09642195f);
    p = fmaf(p, q, -0.13908629f);
    p = fmaf(p, q, 0.19946566f);
    p = fmaf(p, q, -0.33329861f);
    p = fmaf(p, q, 0.99999934f);
    float r = t * p;
    if (ay > ax) r = 1.5707963267948966f - r;
    if (signbit(x)) r = 3.141592653589793f - r;
    return copysignf(r, y);
}

__device__ __forceinline__ void blos_one(const float4 s, float kfac, float geff, float c66f, float4& o, uchar2& f) {
    o = make_float4(0.f, 0.f, 0.f, 0.f);
    f = make_uchar2(0, 0);
    const bool bad = (s.x == 0.f) || isnan(s.x) || isnan(s.y) || isnan(s.z) || isnan(s.w);   // CLEDB_PROC.py:882
    if (bad) return;
    // float32 arithmetic in the reference: NumPy float32 scalars with Python-float constants (weak promotion)
    const float l2 = __fadd_rn(__fmul_rn(s.y, s.y), __fmul_rn(s.z, s.z));
    const float lpol = l2 > 0.f ? l2 * rsqrtf(l2) : 0.f;
    const float nv = -s.w;
    const float d0 = __fsub_rn(__fmul_rn(geff, __fsub_rn(s.x, lpol)), __fmul_rn(c66f, lpol));
    const float d1 = __fadd_rn(__fmul_rn(geff, __fadd_rn(s.x, lpol)), __fmul_rn(c66f, lpol));
    const float d2 = __fmul_rn(geff, s.x);
    o.x = __fmul_rn(kfac, __fdividef(nv, d0));
    o.y = __fmul_rn(kfac, __fdividef(nv, d1));
    o.z = __fmul_rn(kfac, __fdividef(nv, d2));
    o.w = 0.5f * atan2_poly(s.z, s.y);
    const float qi = __fdiv_rn(s.y, s.x), ui = __fdiv_rn(s.z, s.x);
    if (fabsf(qi) < 1e-6f || fabsf(ui) < 1e-6f) {            // :907
        f.x = 1;
        if (qi < 1e-6f && ui < 1e-6f) f.y = 1;                // :909 (signed comparison, as in the reference)
    }
}

__global__ void __launch_bounds__(256) k_blos(int64_t npix, const float4* __restrict__ iquv, float kfac, float geff, float c66f,
                                              float4* __restrict__ out, uchar2* __restrict__ flags) {
    const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;      // pixels base + i*256: every access is coalesced
    float4 s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = base + i * 256;
        s[i] = p < npix ? __ldcs(iquv + p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = base + i * 256;
        float4 o;
        uchar2 f;
        blos_one(s[i], kfac, geff, c66f, o, f);
        if (p < npix) {
            __stcs(out + p, o);
            flags[p] = f;
        }
    }
}

// 72 B/pixel (+8 with a coordinate grid): 32 in, 32 out, aobs, yobs
__global__ void __launch_bounds__(256) k_qurotate(int nx, int ny, const double* __restrict__ xs, const double* __restrict__ ys,
                                                  float linpolref, const float* __restrict__ hpxy, const float4* __restrict__ in,
                                                  float4* __restrict__ out, float* __restrict__ aobs, float* __restrict__ yobs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npix = (int64_t)nx * ny;
    if (p >= npix) return;
    float a, y;
    if (hpxy) {   // Cryo-NIRSP coordinate grid: float32 arithmetic (CLEDB_PREPINV.py:923, 1326)
        const float px = hpxy[p], py = hpxy[npix + p];
        y = __fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
        a = -atan2f(px, py);
    } else {      // header keywords are Python floats: float64 until the store into the float32 arrays
        const double px = xs[p / ny], py = ys[p % ny];
        y = (float)sqrt(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)));
        a = (float)(-atan2(px, py));
    }
    const float twopi = 6.283185307179586f;
    if (a < 0.f) a = __fadd_rn(twopi, a);
    a = (a < linpolref) ? __fsub_rn(twopi, a) : __fadd_rn(a, linpolref);
    float sn, cs;
    sincosf(__fmul_rn(2.0f, a), &sn, &cs);
    const float4 l1 = in[2 * p], l2 = in[2 * p + 1];
    float4 o1 = l1, o2 = l2;
    o1.y = __fsub_rn(__fmul_rn(l1.y, cs), __fmul_rn(l1.z, sn));
    o1.z = __fadd_rn(__fmul_rn(l1.y, sn), __fmul_rn(l1.z, cs));
    o2.y = __fsub_rn(__fmul_rn(l2.y, cs), __fmul_rn(l2.z, sn));
    o2.z = __fadd_rn(__fmul_rn(l2.y, sn), __fmul_rn(l2.z, cs));
    out[2 * p] = o1;
    out[2 * p + 1] = o2;
    aobs[p] = a;
    yobs[p] = y;
}

// FP32 FFMA-pipe peak: 16 independent FMA chains per thread, no memory traffic.  The denominator of the search
// kernel's roofline (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
__global__ void __launch_bounds__(256) k_ffma_peak(float* sink, int iters, float x, float y) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (float)(threadIdx.x + j) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    if (s == 123.456f) sink[0] = s;    // never true; keeps the chains alive
}

int measure_fp32_peak(cledb_ctx* ctx, double* tflops) {
    float* sink = nullptr;
    CLEDB_CUDA(ctx, cudaMalloc(&sink, 256));
    const int iters = 1 << 14, grid = ctx->prop.multiProcessorCount * 8;
    cudaEvent_t a, b;
    CLEDB_CUDA(ctx, cudaEventCreate(&a));
    CLEDB_CUDA(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CLEDB_CUDA(ctx, cudaEventRecord(a, ctx->stream));
        k_ffma_peak<<<grid, 256, 0, ctx->stream>>>(sink, iters, 0.999f, 1e-3f);
        CLEDB_CUDA(ctx, cudaEventRecord(b, ctx->stream));
        CLEDB_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        CLEDB_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        const double tf = 2.0 * 16.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        ctx->launches++;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *tflops = best;
    return CLEDB_OK;
}

int launch_blos(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac, float* out,
                uint8_t* flags, cudaStream_t st) {
    if (npix == 0) return CLEDB_OK;
    // Python evaluates 0.66*F_factor in float64; NumPy then rounds that product, g_eff and the constant K to
    // float32 when they meet float32 scalars (CLEDB_PROC.py:889-896)
    k_blos<<<(int)((npix + 1023) / 1024), 256, 0, st>>>(npix, reinterpret_cast<const float4*>(iquv), (float)kfac, (float)g_eff,
                                                      (float)(0.66 * ffac), reinterpret_cast<float4*>(out),
                                                      reinterpret_cast<uchar2*>(flags));
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

int launch_qurotate(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref, const float* hpxy,
                    const float* in, float* out, float* aobs, float* yobs, cudaStream_t st) {
    const int64_t npix = (int64_t)nx * ny;
    if (npix == 0) return CLEDB_OK;
    k_qurotate<<<(int)((npix + 255) / 256), 256, 0, st>>>(nx, ny, xs, ys, (float)linpolref, hpxy,
                                                          reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), aobs, yobs);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

}  // namespace cledb
