// Spectral integration of the observation: obs_integrate_work_1pix + obs_cdf of the reference
// (CLEDB_PREPINV.py:1164-1270, 1356-1364), one warp per spectrum.  Everything is float64 in NumPy's operation order -
// including numpy.sum's pairwise summation, numpy.mean along axis 0 (row after row), numpy.gradient with caller-made
// coefficients and numpy.median - so the results are bit-identical to the reference for float64 input.
#include <math_constants.h>

#include "cledb_internal.h"

namespace cledb {

constexpr int kIntWarps = 4;

// numpy's pairwise summation of n terms f(0..n-1) (numpy/_core/src/umath/loops_utils.h.src, blocksize 128)
template <class F> __device__ double pw_sum(F f, int i0, int n) {
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, f(i0 + i));
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = f(i0 + j);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(i0 + i + j));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, f(i0 + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(pw_sum(f, i0, n2), pw_sum(f, i0 + n2, n - n2));
}

__device__ __forceinline__ double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(kIntWarps * 32) k_integrate(int64_t npix, const double* __restrict__ spec, cledb_spectral sp,
                                                              float* __restrict__ tot_out, float* __restrict__ snr_out,
                                                              float* __restrict__ bkg_out, int32_t* __restrict__ iss_out,
                                                              uint8_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * kIntWarps + warp;
    if (p >= npix) return;
    const int nw = sp.nw;
    double* fn = reinterpret_cast<double*>(smem) + (size_t)warp * 2 * nw;       // [nw] I/Itot, later squares
    double* vs = fn + nw;                                                        // [nw] finite V / slope values
    float* cdf = reinterpret_cast<float*>(vs);                                   // [nw] float32 running sums (before vs is used)
    const double* s = spec + p * (int64_t)nw * 4;

    auto zeros = [&](int st) {
        if (lane < 4) { tot_out[p * 4 + lane] = 0.f; snr_out[p * 4 + lane] = 0.f; bkg_out[p * 4 + lane] = 0.f; }
        if (lane == 0) { iss_out[p] = 0; status[p] = (uint8_t)st; }
    };
    // ---- counts and finiteness (:1168)
    bool finite = true, nonzero = false;
    for (int i = lane; i < nw * 4; i += 32) {
        const double v = s[i];
        finite = finite && isfinite(v);
        if ((i & 3) == 0 && v != 0.0) nonzero = true;
    }
    if (!__all_sync(0xffffffffu, finite) || !__any_sync(0xffffffffu, nonzero)) { zeros(0); return; }

    // ---- running sum of Stokes I, every prefix its own pairwise sum, float32 (obs_cdf)
    for (int i = lane; i < nw; i += 32) cdf[i] = (float)pw_sum([&](int q) { return s[4 * q]; }, 0, i + 1);
    __syncwarp();
    const float last = cdf[nw - 1];
    int l = 0x7fffffff, r = -1, cnt = 0;
    for (int i = lane; i < nw; i += 32) {
        const float c = fabsf(__fdiv_rn(cdf[i], last));
        if (c > 0.003f && c < 0.997f) { l = min(l, i); r = max(r, i); ++cnt; }
    }
    l = __reduce_min_sync(0xffffffffu, l);
    r = __reduce_max_sync(0xffffffffu, r);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt == 0) { zeros(1); return; }                     // lr0[0,0] raises IndexError in the reference
    __syncwarp();

    // ---- background: mean of (up to) three continuum rows, numpy.mean along axis 0 (:1192)
    const int m = max(0, min(3, nw - sp.wcont));
    double bkg = 0.0;
    if (lane < 4) {
        for (int q = 0; q < m; ++q) bkg = q == 0 ? s[(sp.wcont + q) * 4 + lane] : __dadd_rn(bkg, s[(sp.wcont + q) * 4 + lane]);
        bkg = m > 0 ? __ddiv_rn(bkg, (double)m) : CUDART_NAN;
    }
    const double b0 = warp_bcast(bkg, 0), b3 = warp_bcast(bkg, 3);
    auto ti = [&](int q) { const double v = __dsub_rn(s[4 * q], b0); return v < 0.0 ? 0.0 : v; };      // tmp2[:,0], negatives clipped
    // ---- integrals of I, Q, U (:1201, :1222-1223): lanes 0..2
    double tot = 0.0;
    if (lane == 0) tot = pw_sum([&](int q) { return __dmul_rn(ti(q), sp.dwv); }, 0, nw);
    if (lane == 1 || lane == 2) tot = pw_sum([&](int q) { return __dmul_rn(__dsub_rn(s[4 * q + lane], bkg), sp.dwv); }, 0, nw);
    const double tot0 = warp_bcast(tot, 0);
    // ---- Stokes V: median of V / (d(I/Itot)/dlambda + 1e-36) over the inner half (:1209-1216)
    for (int i = lane; i < nw; i += 32) fn[i] = __ddiv_rn(ti(i), tot0);
    __syncwarp();
    int nfin = 0;
    for (int base = sp.lo; base < sp.hi; base += 32) {
        const int i = base + lane;
        double v = 0.0;
        bool ok = false;
        if (i < sp.hi) {
            double g;
            if (i == 0) g = __ddiv_rn(__dsub_rn(fn[1], fn[0]), sp.dx_first);
            else if (i == nw - 1) g = __ddiv_rn(__dsub_rn(fn[nw - 1], fn[nw - 2]), sp.dx_last);
            else if (sp.uniform) g = __ddiv_rn(__dsub_rn(fn[i + 1], fn[i - 1]), sp.h2);
            else g = __dadd_rn(__dadd_rn(__dmul_rn(sp.ga[i - 1], fn[i - 1]), __dmul_rn(sp.gb[i - 1], fn[i])), __dmul_rn(sp.gc[i - 1], fn[i + 1]));
            v = __ddiv_rn(__dsub_rn(s[4 * i + 3], b3), __dadd_rn(g, 1e-36));
            ok = isfinite(v);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (ok) vs[nfin + __popc(bal & ((1u << lane) - 1u))] = v;
        nfin += __popc(bal);
    }
    __syncwarp();
    double med = CUDART_NAN;                                 // numpy.median of an empty array
    if (nfin > 0) {
        const int ra = (nfin - 1) / 2, rb = nfin / 2;        // the two middle ranks (equal when nfin is odd)
        double va = 0.0, vb = 0.0;
        for (int i = lane; i < nfin; i += 32) {
            const double x = vs[i];
            int rank = 0;
            for (int j = 0; j < nfin; ++j) rank += (vs[j] < x || (vs[j] == x && j < i)) ? 1 : 0;
            if (rank == ra) va = x;
            if (rank == rb) vb = x;
        }
        // exactly one lane holds each middle value: a sum over lanes of (value or +0) recovers it exactly
        for (int o = 16; o > 0; o >>= 1) { va += __shfl_xor_sync(0xffffffffu, va, o); vb += __shfl_xor_sync(0xffffffffu, vb, o); }
        med = ra == rb ? va : __ddiv_rn(__dadd_rn(va, vb), 2.0);
    }
    // ---- signal-to-noise (:1255-1258)
    double noise = 0.0;
    if (lane == 0) {
        const int nl = (l + 1) * 4, nr = (nw - r) * 4;       // rows 0..l and r..nw-1, flattened row-major, squared
        const double ss = pw_sum([&](int q) { const double v = q < nl ? s[q] : s[r * 4 + (q - nl)]; return __dmul_rn(v, v); }, 0, nl + nr);
        noise = sqrt(__ddiv_rn(ss, (double)(nl + nr)));
    }
    noise = warp_bcast(noise, 0);
    const int a = (int)(int16_t)rint((double)l + (double)cnt / 3.0), b = (int)(int16_t)rint((double)l + (double)(cnt * 2) / 3.0);
    double snr = 0.0;
    if (lane < 4) {
        const int n = max(0, min(b, nw) - max(a, 0));
        double acc = 0.0;
        for (int q = 0; q < n; ++q) {
            const double v = s[(a + q) * 4 + lane], sq = __dmul_rn(v, v);
            acc = q == 0 ? sq : __dadd_rn(acc, sq);
        }
        const double sig = n > 0 ? sqrt(__ddiv_rn(acc, (double)n)) : CUDART_NAN;
        snr = __ddiv_rn(sig, noise);
    }
    const uint32_t low = __ballot_sync(0xffffffffu, lane < 4 && snr < 1.0);
    if (lane < 4) {
        const double t = lane == 0 ? tot0 : (lane == 3 ? med : tot);
        tot_out[p * 4 + lane] = (float)t;
        snr_out[p * 4 + lane] = (float)snr;
        bkg_out[p * 4 + lane] = (float)bkg;
    }
    if (lane == 0) {
        iss_out[p] = ((low & 7u) ? 1 : 0) + ((low & 8u) ? 2 : 0);
        status[p] = 0;
    }
}

// ---- density from the line ratio: obs_dens_work_1pix (CLEDB_PREPINV.py:1436-1479), one thread per pixel
__global__ void __launch_bounds__(256) k_obs_dens(int64_t npix, const float* __restrict__ a_obs, const float* __restrict__ b_obs,
                                                  const float* __restrict__ yobs, int nh, const double* __restrict__ heights, int nknots,
                                                  const double* __restrict__ knots, int ncoef, const double* __restrict__ coefs,
                                                  float* __restrict__ dens, int32_t* __restrict__ issue) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float a = a_obs[p], b = b_obs[p];
    if (b == 0.f) { dens[p] = 1.f; issue[p] = 0; return; }                       // :1444
    const float rat = __fdiv_rn(a, b);
    if (isnan(rat) || isinf(rat)) { dens[p] = 0.f; issue[p] = 0; return; }       // :1450
    const double y = (double)yobs[p];
    int row = -1;
    for (int i = 0; i < nh; ++i)
        if (heights[i] > y) { row = i; break; }                                  // first table height above the pixel, :1454
    int iss = 0;
    if (row < 0) { row = nh - 1; iss = 4; }                                      // :1457-1459
    const double* t = knots + (size_t)row * nknots;
    const double* c = coefs + (size_t)row * ncoef;
    const double x = (double)rat;
    // scipy's find_interval with extrapolation: t[l] <= x < t[l+1], clamped to [k, n-1]
    const int k = 2, n = nknots - k - 1;
    int l = k;
    while (x < t[l] && l != k) --l;
    ++l;
    while (x >= t[l] && l != n) ++l;
    const int ell = l - 1;
    // de Boor's recurrence for the k+1 non-zero B-splines at x (scipy _deBoor_D, no derivative)
    double h[3] = {1.0, 0.0, 0.0}, hh[3];
    for (int j = 1; j <= k; ++j) {
        for (int q = 0; q < j; ++q) hh[q] = h[q];
        h[0] = 0.0;
        for (int m = 1; m <= j; ++m) {
            const int ind = ell + m;
            const double xb = t[ind], xa = t[ind - j];
            if (xb == xa) { h[m] = 0.0; continue; }
            const double w = __ddiv_rn(hh[m - 1], __dsub_rn(xb, xa));
            h[m - 1] = __dadd_rn(h[m - 1], __dmul_rn(w, __dsub_rn(xb, x)));
            h[m] = __dmul_rn(w, __dsub_rn(x, xa));
        }
    }
    double out = 0.0;
    for (int q = 0; q <= k; ++q) out = __dadd_rn(out, __dmul_rn(c[ell + q - k], h[q]));
    dens[p] = (float)out;
    issue[p] = iss;
}

}  // namespace cledb

using namespace cledb;

extern "C" int cledb_obs_dens(cledb_ctx* ctx, int64_t npix, const float* a_obs, const float* b_obs, const float* yobs, int nh,
                              const double* heights, int nknots, const double* knots, int ncoef, const double* coefs, float* dens_out,
                              int32_t* issue_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || !a_obs || !b_obs || !yobs || nh < 1 || !heights || nknots < 6 || !knots || ncoef != nknots - 3 || !coefs || !dens_out ||
        !issue_out) {
        ctx->err = "cledb_obs_dens: bad argument (degree-2 splines: ncoef = nknots - 3)";
        return CLEDB_ERR_ARG;
    }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up(off + b); return o; };
    const size_t o_a = take(n * 4), o_b = take(n * 4), o_y = take(n * 4), o_h = take((size_t)nh * 8), o_t = take((size_t)nh * nknots * 8),
                 o_c = take((size_t)nh * ncoef * 8), o_d = take(n * 4), o_i = take(n * 4);
    int rc = ensure_workspace(ctx, ctx->io, off);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_a, a_obs, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_b, b_obs, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_y, yobs, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_h, heights, (size_t)nh * 8, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_t, knots, (size_t)nh * nknots * 8, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_c, coefs, (size_t)nh * ncoef * 8, cudaMemcpyHostToDevice, st));
    k_obs_dens<<<(int)((npix + 255) / 256), 256, 0, st>>>(npix, reinterpret_cast<float*>(w + o_a), reinterpret_cast<float*>(w + o_b),
                                                         reinterpret_cast<float*>(w + o_y), nh, reinterpret_cast<double*>(w + o_h), nknots,
                                                         reinterpret_cast<double*>(w + o_t), ncoef, reinterpret_cast<double*>(w + o_c),
                                                         reinterpret_cast<float*>(w + o_d), reinterpret_cast<int32_t*>(w + o_i));
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    CLEDB_CUDA(ctx, cudaMemcpyAsync(dens_out, w + o_d, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(issue_out, w + o_i, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return CLEDB_OK;
}


extern "C" int cledb_obs_integrate(cledb_ctx* ctx, int64_t npix, const double* spec, const cledb_spectral* sp, float* tot, float* snr,
                                   float* background, int32_t* issue, uint8_t* status) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || !spec || !sp || !tot || !snr || !background || !issue || !status || sp->nw < 3 || sp->wcont < 0 || sp->lo < 0 ||
        sp->hi > sp->nw || (!sp->uniform && (!sp->ga || !sp->gb || !sp->gc))) {
        ctx->err = "cledb_obs_integrate: bad argument";
        return CLEDB_ERR_ARG;
    }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix, nw = (size_t)sp->nw;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up(off + b); return o; };
    const size_t o_spec = take(n * nw * 32), o_g = take(3 * nw * 8), o_tot = take(n * 16), o_snr = take(n * 16), o_bkg = take(n * 16),
                 o_iss = take(n * 4), o_st = take(n);
    int rc = ensure_workspace(ctx, ctx->io, off);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_spec, spec, n * nw * 32, cudaMemcpyHostToDevice, st));
    cledb_spectral d = *sp;
    double* g = reinterpret_cast<double*>(w + o_g);
    if (!sp->uniform) {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(g, sp->ga, (nw - 2) * 8, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(g + nw, sp->gb, (nw - 2) * 8, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(g + 2 * nw, sp->gc, (nw - 2) * 8, cudaMemcpyHostToDevice, st));
    }
    d.ga = g; d.gb = g + nw; d.gc = g + 2 * nw;
    const size_t smem = (size_t)kIntWarps * 2 * nw * 8;
    if (smem > 200 * 1024) { ctx->err = "cledb_obs_integrate: too many wavelength samples for shared memory"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_integrate<<<(int)((npix + kIntWarps - 1) / kIntWarps), kIntWarps * 32, smem, st>>>(
        npix, reinterpret_cast<const double*>(w + o_spec), d, reinterpret_cast<float*>(w + o_tot), reinterpret_cast<float*>(w + o_snr),
        reinterpret_cast<float*>(w + o_bkg), reinterpret_cast<int32_t*>(w + o_iss), w + o_st);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    CLEDB_CUDA(ctx, cudaMemcpyAsync(tot, w + o_tot, n * 16, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(snr, w + o_snr, n * 16, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(background, w + o_bkg, n * 16, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(issue, w + o_iss, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(status, w + o_st, n, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return CLEDB_OK;
}
