// Solution building on the device: what cledb_matchiquv / cledb_matchiqud do with the nsearch winners after the search
// (CLEDB_PROC.py:1010-1082 / 1185-1247 of the reference) for every pixel of a map at once, and the fused entry points
// cledb_invert / cledb_invert_dev (search + solutions in one call).
//
// One thread per pixel.  Every floating-point operation is written with explicit rounding intrinsics in the
// reference's dtype and order (NumPy float32 scalars, float64 where an int64 index meets float32 header values), so
// nothing is contracted into FMAs; sines and cosines come from caller-evaluated tables (see the header), so the
// outputs are bit-identical to the NumPy host path of cledb_b200/solution.py.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cledb_internal.h"

namespace cledb {

constexpr int kMaxK = CLEDB_MAX_NSEARCH;

// The reference's partial selection sort reports argmin(tmp[i:]) + i of an array it is swapping in place
// (CLEDB_PROC.py:1620-1625): an early candidate that is itself a winner, but not at its turn, is swapped away and found
// again at the old position of an earlier winner, so that earlier winner is reported twice.  Only the winners'
// positions q[] and the first k positions are ever touched, so the sort is replayed on that sparse set.
// sel[i] = which of the true top-k the reference shows in slot i.
__device__ void replay_partsort(int k, const int32_t* q, const double* chi, int* sel) {
    int pos[2 * kMaxK];
    double work[2 * kMaxK];
    int slot[2 * kMaxK];
    int n = 0;
    // sorted union of 0..k-1 and q[]: q is not sorted by position, so insert one by one
    for (int i = 0; i < k; ++i) { pos[n] = i; work[n] = CUDART_INF; slot[n] = -1; ++n; }
    for (int m = 0; m < k; ++m) {
        const int p = q[m];
        if (p < k) { work[p] = chi[m]; slot[p] = m; continue; }
        int at = n;
        while (at > k && pos[at - 1] > p) { pos[at] = pos[at - 1]; work[at] = work[at - 1]; slot[at] = slot[at - 1]; --at; }
        pos[at] = p; work[at] = chi[m]; slot[at] = m;
        ++n;
    }
    for (int i = 0; i < k; ++i) {
        int a = i;
        for (int j = i + 1; j < n; ++j)
            if (work[j] < work[a]) a = j;            // first minimum, like np.argmin
        // the winner's identity stays with its ORIGINAL position: the reference reads chisq[asrt[i]] from the
        // unswapped array (CLEDB_PROC.py:1043), so slot[] is not swapped along with work[]
        sel[i] = slot[a] >= 0 ? slot[a] : i;
        const double t = work[i]; work[i] = work[a]; work[a] = t;
    }
}

template <bool DOPP64>
__global__ void __launch_bounds__(128) k_build(int mode, int64_t npix, int k, const float* __restrict__ sobs,
                                               const float* __restrict__ yobs, const float* __restrict__ dobs,
                                               const float* __restrict__ rot_cos, const float* __restrict__ rot_sin,
                                               const void* __restrict__ dopp, int64_t dopp_stride, cledb_dbgrid g, double maxchisq,
                                               int bcalc, int strict_topk, const int32_t* __restrict__ idx,
                                               const float* __restrict__ chi, const double* __restrict__ chi64,
                                               const int32_t* __restrict__ kpos, const float* __restrict__ rows,
                                               const uint8_t* __restrict__ status, const int32_t* __restrict__ nan_slots,
                                               float* __restrict__ invout, float* __restrict__ sfound,
                                               const float* __restrict__ snr, int32_t* __restrict__ issue_out) {
    // one thread per (pixel, output slot); the 44- and 32-byte records of a block are staged in shared memory and stored
    // with coalesced 16-byte writes
    __shared__ __align__(16) float s_inv[128 * 11];
    __shared__ __align__(16) float s_sf[128 * 8];
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x, t = t0 + threadIdx.x;
    float o[11], f[8];
    float v_sth = 0.f, v_cth = 1.f, v_sph = 0.f, v_cph = 1.f;       // sin / cos of the solution's theta_B, phi_B (0, 0 without a solution)
    o[0] = -1.0f;
#pragma unroll
    for (int c = 1; c < 11; ++c) o[c] = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = 0.0f;
    do {
    if (t >= npix * k) break;
    const int64_t p = t / k;
    const int s = (int)(t - p * k);
    const int st = status[p];
    if (st == CLEDB_PIX_ALLINF) {                       // every chi^2 is +inf in the reference: rejected with that value
        o[1] = CUDART_INF_F;
        break;
    }
    if (st != CLEDB_PIX_SEARCHED) break;

    const int32_t* ix = idx + p * k;
    const int32_t* kp = kpos + p * k;
    bool all_valid = true, touched = false;
    for (int q = 0; q < k; ++q) {
        all_valid = all_valid && ix[q] >= 0;
        touched = touched || (ix[q] >= 0 && kp[q] >= 0 && kp[q] < k);
    }
    // reduced mode: candidates with a NaN chi^2 come first in the reference's selection sort and leave their slots empty
    const int shift = nan_slots ? nan_slots[p] : 0;
    int w = s - shift, w0 = 0;                          // winners shown in this slot and in slot 0
    if (!strict_topk && touched && all_valid && shift == 0) {
        double c64[kMaxK];
        int sel[kMaxK];
        for (int q = 0; q < k; ++q) c64[q] = chi64 ? chi64[p * k + q] : (double)chi[p * k + q];
        replay_partsort(k, kp, c64, sel);
        w = sel[s];
        w0 = sel[0];
    }
    auto chi_of = [&](int q) { return chi64 ? chi64[p * k + q] : (double)chi[p * k + q]; };

    const double limit = mode == CLEDB_MODE_IQUV ? maxchisq * 10.0 : maxchisq;          // CLEDB_PROC.py:1027 / 1202
    if (shift == 0 && ix[w0] >= 0 && chi_of(w0) > limit) {
        o[1] = (float)chi_of(w0);
        break;
    }
    if (w < 0 || ix[w] < 0) break;
    const double cw = chi_of(w);
    if (!(cw <= maxchisq)) break;                                                        // :1044 / :1219
    float so[8];
    {
        const float4 a = reinterpret_cast<const float4*>(sobs)[2 * p], b = reinterpret_cast<const float4*>(sobs)[2 * p + 1];
        so[0] = a.x; so[1] = a.y; so[2] = a.z; so[3] = a.w; so[4] = b.x; so[5] = b.y; so[6] = b.z; so[7] = b.w;
    }
    float nf = so[0];                                                                    // :980
    if (mode == CLEDB_MODE_IQUD) {
#pragma unroll
        for (int c = 1; c < 8; ++c) nf = fmaxf(nf, so[c]);                              // :1155 (no NaN in a searched IQUD pixel)
    }
    const float yo = yobs[p], ed = dobs[p], rc = rot_cos[p], rs = rot_sin[p];
    const int64_t n3 = g.nbtheta, n4 = (int64_t)g.nbphi * g.nbtheta;
    const double dgx = (double)__fsub_rn(g.gxmax, g.gxmin), dphi = (double)__fsub_rn(g.bphimax, g.bphimin),
                 dth = (double)__fsub_rn(g.bthetamax, g.bthetamin);
    {
        const float* r = rows + ((size_t)p * k + w) * 8;
        // ---- field strength
        float b32 = 0.f;
        double b64 = 0.0;
        if (mode == CLEDB_MODE_IQUV) {                                                   // :1049-1059
            const float r3 = fabsf(r[3]) > 1e-7f ? __fdiv_rn(__fdiv_rn(so[3], nf), r[3]) : 0.f;
            const float r7 = fabsf(r[7]) > 1e-7f ? __fdiv_rn(__fdiv_rn(so[7], nf), r[7]) : 0.f;
            b32 = bcalc == 0 ? r3 : (bcalc == 1 ? r7 : __fmul_rn(0.5f, __fadd_rn(r7, r3)));
        } else if (DOPP64) {
            b64 = reinterpret_cast<const double*>(dopp)[p * dopp_stride];                // :1225
        } else {
            b32 = reinterpret_cast<const float*>(dopp)[p * dopp_stride];
        }
        // ---- grid position of the entry, cledb_params + cledb_physcle (:1634-1659, :1727-1749); PyCELP type: i = 0
        const int64_t index = ix[w];
        const int64_t j = index / n4, kk = (index - j * n4) / n3, l = index - j * n4 - kk * n3;
        const float gx = (float)__dadd_rn((double)g.gxmin, __ddiv_rn(__dmul_rn((double)j, dgx), (double)(g.ngx - 1)));
        const float bphi = (float)__dadd_rn((double)g.bphimin, __ddiv_rn(__dmul_rn((double)kk, dphi), (double)(g.nbphi - 1)));
        const float bth = (float)__dadd_rn((double)g.bthetamin, __ddiv_rn(__dmul_rn((double)l, dth), (double)(g.nbtheta - 1)));
        const float sp = g.sin_bphi[kk], cp = g.cos_bphi[kk], sth = g.sin_btheta[l], cth = g.cos_btheta[l];
        v_sth = sth; v_cth = cth; v_sph = sp; v_cph = cp;
        float bo, bx, by, bz;
        if (DOPP64 && mode == CLEDB_MODE_IQUD) {                                         // float64 B_POS: NumPy promotes to float64
            double b = b64;
            if (bcalc == 3) {
                const float den = __fsqrt_rn(__fadd_rn(__fmul_rn(__fmul_rn(sth, sth), __fmul_rn(sp, sp)), __fmul_rn(cth, cth)));
                b = __ddiv_rn(b, (double)den);
            }
            const double ab = fabs(b);
            bo = (float)b;
            bx = (float)__dmul_rn(__dmul_rn(ab, (double)sth), (double)cp);
            by = (float)__dmul_rn(__dmul_rn(ab, (double)sth), (double)sp);
            bz = (float)__dmul_rn(ab, (double)cth);
        } else {
            float b = b32;
            if (bcalc == 3) {                                                            // :1712-1713
                const float den = __fsqrt_rn(__fadd_rn(__fmul_rn(__fmul_rn(sth, sth), __fmul_rn(sp, sp)), __fmul_rn(cth, cth)));
                b = __fdiv_rn(b, den);
            }
            const float ab = fabsf(b);
            bo = b;
            bx = __fmul_rn(__fmul_rn(ab, sth), cp);
            by = __fmul_rn(__fmul_rn(ab, sth), sp);
            bz = __fmul_rn(ab, cth);
        }
        o[0] = (float)index;                       // the reference stores the index in a float32 array (exact below 2^24)
        o[1] = (float)cw;
        o[2] = ed; o[3] = yo; o[4] = gx; o[5] = bo; o[6] = bphi; o[7] = bth; o[8] = bx; o[9] = by; o[10] = bz;
        // ---- matched profile, rotated back: cledb_quderotate (:1575-1588)
        const float q1 = __fmul_rn(r[1], nf), u1 = __fmul_rn(r[2], nf), q2 = __fmul_rn(r[5], nf), u2 = __fmul_rn(r[6], nf);
        f[0] = __fmul_rn(r[0], nf);
        f[1] = __fsub_rn(__fmul_rn(q1, rc), __fmul_rn(u1, rs));
        f[2] = __fadd_rn(__fmul_rn(q1, rs), __fmul_rn(u1, rc));
        f[3] = __fmul_rn(r[3], nf);
        f[4] = __fmul_rn(r[4], nf);
        f[5] = __fsub_rn(__fmul_rn(q2, rc), __fmul_rn(u2, rs));
        f[6] = __fadd_rn(__fmul_rn(q2, rs), __fmul_rn(u2, rc));
        f[7] = __fmul_rn(r[7], nf);
    }
    } while (0);
    // ---- issue codes of cledb_invproc (CLEDB_PROC.py:323-355): one int32 per pixel (zeroed before the launch), OR-ed together by
    //   the thread of the LAST solution (`flg = 0` sits inside the reference's loop over solutions, :324, so only slot
    //          nsearch-1 decides code 512):  512 = Van-Vleck proximity 53 deg <= vartheta_B <= 56 deg, and bit 0 = the angle
    //          lies within 2e-3 deg of either limit: the reference evaluates the chain in float32 with NumPy's own
    //          sin/cos/arctan2, whose last bits differ between CPUs; here it runs in float64 without a single transcendental
    //          (cos/sin of alpha = -atan2(a, b) are b/r and -a/r, the sines and cosines of the grid angles are the caller's
    //          tables, and 53 deg <= atan2(s, z) <= 56 deg is tan(53 deg) z <= s <= tan(56 deg) z for z > 0), which decides every
    //          pixel that is not that close to a limit; the binding re-evaluates the marked ones (about 1 in 25 000) with NumPy
    //   the thread of slot 0:  1024 = chi^2 of the best solution > 10 (:351), 2048 = snr[p,0] < 1 (:353)
    if (issue_out && t < npix * k) {
        const int64_t p = t / k;
        const int s = (int)(t - p * k);
        int code = 0;
        if (s == k - 1) {
            const double a = (double)o[4], b = (double)o[3];
            const double r2 = a * a + b * b;
            double ca, sa;
            if (r2 > 0.0 && r2 < 1e300) {
                const double ir = 1.0 / sqrt(r2);
                ca = b * ir;
                sa = -a * ir;
            } else if (a == 0.0 && b == 0.0) {                                  // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
                ca = signbit(b) ? -1.0 : 1.0;
                sa = 0.0;
            } else {                                                            // infinities, NaN: the library's conventions
                const double alpha = -atan2(a, b);
                ca = cos(alpha);
                sa = sin(alpha);
            }
            const double cb = 6.123233995736766e-17;                            // cos of the float64 pi/2 (its sine is 1)
            const double xb = (double)v_sth * (double)v_cph, yb = (double)v_sth * (double)v_sph, zb = (double)v_cth;
            const double yyb = cb * yb - zb, yzb = yb + cb * zb;
            const double xxb = ca * xb + sa * yzb, zzb = -sa * xb + ca * yzb;
            const double sq = sqrt(xxb * xxb + yyb * yyb);
            constexpr double t53 = 1.3270448216204098, t56 = 1.4825609685127403;          // tan(53 deg), tan(56 deg)
            constexpr double t53lo = 1.326948447329313, t53hi = 1.3271412048405367;       // tan((53 -+ 2e-3) deg)
            constexpr double t56lo = 1.4824493434833603, t56hi = 1.4826726050961634;      // tan((56 -+ 2e-3) deg)
            if (zzb > 0.0 && sq >= t53 * zzb && sq <= t56 * zzb) code |= 512;
            if (zzb > 0.0 && ((sq > t53lo * zzb && sq < t53hi * zzb) || (sq > t56lo * zzb && sq < t56hi * zzb))) code |= 1;
        }
        if (s == 0) code |= (o[1] > 10.0f ? 1024 : 0) | ((snr && snr[p * 8] < 1.0f) ? 2048 : 0);
        if (code) atomicOr(issue_out + p, code);
    }
#pragma unroll
    for (int c = 0; c < 11; ++c) s_inv[threadIdx.x * 11 + c] = o[c];
#pragma unroll
    for (int c = 0; c < 8; ++c) s_sf[threadIdx.x * 8 + c] = f[c];
    __syncthreads();
    const int64_t nthr = min((int64_t)blockDim.x, npix * k - t0);                       // output slots of this block
    if (nthr <= 0) return;
    float4* gi = reinterpret_cast<float4*>(invout + t0 * 11);                           // t0 * 44 B is a multiple of 16
    float4* gs = reinterpret_cast<float4*>(sfound + t0 * 8);
    const int n4i = (int)(nthr * 11 / 4), n4s = (int)(nthr * 2);
    for (int i = threadIdx.x; i < n4i; i += blockDim.x) gi[i] = reinterpret_cast<const float4*>(s_inv)[i];
    for (int i = n4i * 4 + threadIdx.x; i < nthr * 11; i += blockDim.x) invout[t0 * 11 + i] = s_inv[i];      // tail of a partial block
    for (int i = threadIdx.x; i < n4s; i += blockDim.x) gs[i] = reinterpret_cast<const float4*>(s_sf)[i];
}

int launch_build(cledb_ctx* ctx, int mode, int precision, int64_t npix, int nsearch, const float* sobs, const float* yobs,
                 const float* dobs, const float* rot_cos, const float* rot_sin, const void* dopp, int dopp_is_f64,
                 int64_t dopp_stride, const cledb_dbgrid& g, double maxchisq, int bcalc, int strict_topk, const int32_t* idx,
                 const float* chi, const double* chi64, const int32_t* kpos, const float* rows, const uint8_t* status,
                 const int32_t* nan_slots, float* invout, float* sfound, cudaStream_t st, const float* snr, int32_t* issue_out) {
    (void)precision;
    if (issue_out) CLEDB_CUDA(ctx, cudaMemsetAsync(issue_out, 0, (size_t)npix * 4, st));
    const int nb = (int)((npix * nsearch + 127) / 128);
    if (dopp_is_f64)
        k_build<true><<<nb, 128, 0, st>>>(mode, npix, nsearch, sobs, yobs, dobs, rot_cos, rot_sin, dopp, dopp_stride, g, maxchisq,
                                          bcalc, strict_topk, idx, chi, chi64, kpos, rows, status, nan_slots, invout, sfound, snr, issue_out);
    else
        k_build<false><<<nb, 128, 0, st>>>(mode, npix, nsearch, sobs, yobs, dobs, rot_cos, rot_sin, dopp, dopp_stride, g, maxchisq,
                                           bcalc, strict_topk, idx, chi, chi64, kpos, rows, status, nan_slots, invout, sfound, snr, issue_out);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

int grid_to_device(cledb_ctx* ctx, const cledb_dbgrid* grid, cledb_dbgrid* g_dev, cudaStream_t st) {
    const size_t nphi = (size_t)grid->nbphi, nth = (size_t)grid->nbtheta, n = 2 * nphi + 2 * nth;
    std::vector<float> now(n);
    std::memcpy(now.data(), grid->sin_bphi, nphi * 4);
    std::memcpy(now.data() + nphi, grid->cos_bphi, nphi * 4);
    std::memcpy(now.data() + 2 * nphi, grid->sin_btheta, nth * 4);
    std::memcpy(now.data() + 2 * nphi + nth, grid->cos_btheta, nth * 4);
    const bool same = ctx->d_tab && now.size() == ctx->tab_host.size() && std::memcmp(now.data(), ctx->tab_host.data(), n * 4) == 0;
    if (!same) {
        if (!ctx->d_tab || now.size() > ctx->tab_host.capacity() || now.size() != ctx->tab_host.size()) {
            CLEDB_CUDA(ctx, cudaStreamSynchronize(st));                 // nothing reads the old tables any more
            if (ctx->d_tab) cudaFree(ctx->d_tab);
            ctx->d_tab = nullptr;
            ctx->tab_host.clear();
            if (cudaMalloc(&ctx->d_tab, n * 4) != cudaSuccess) { cudaGetLastError(); ctx->err = "out of device memory (grid tables)"; return CLEDB_ERR_NOMEM; }
        }
        ctx->tab_host.swap(now);                                        // the copy below reads the context's own buffer
        CLEDB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tab, ctx->tab_host.data(), n * 4, cudaMemcpyHostToDevice, st));
    }
    *g_dev = *grid;
    g_dev->sin_bphi = ctx->d_tab;
    g_dev->cos_bphi = ctx->d_tab + nphi;
    g_dev->sin_btheta = ctx->d_tab + 2 * nphi;
    g_dev->cos_btheta = ctx->d_tab + 2 * nphi + nth;
    return CLEDB_OK;
}

}  // namespace cledb

using namespace cledb;

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static int check_invert_args(cledb_ctx* ctx, int mode, int64_t npix, const void* sobs, const void* snr, const void* db_id,
                             int nsearch, const void* yobs, const void* dobs, const void* rc, const void* rs, const void* dopp,
                             const cledb_dbgrid* g, int bcalc, const void* invout, const void* sfound) {
    if (npix < 0 || !sobs || !snr || !db_id || !yobs || !dobs || !rc || !rs || !g || !invout || !sfound) {
        ctx->err = "cledb_invert: null argument";
        return CLEDB_ERR_ARG;
    }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "cledb_invert: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (mode != CLEDB_MODE_IQUV && mode != CLEDB_MODE_IQUD) { ctx->err = "cledb_invert: unknown mode"; return CLEDB_ERR_ARG; }
    if (g->dbtype != 1) { ctx->err = "cledb_invert: only PyCELP-type databases (dbtype 1) are built on the device"; return CLEDB_ERR_ARG; }
    if (g->ngx < 1 || g->nbphi < 1 || g->nbtheta < 1 || !g->sin_bphi || !g->cos_bphi || !g->sin_btheta || !g->cos_btheta) {
        ctx->err = "cledb_invert: incomplete grid description";
        return CLEDB_ERR_ARG;
    }
    if (bcalc < 0 || bcalc > 3 || (mode == CLEDB_MODE_IQUD) != (bcalc == 3)) {           // CLEDB_PROC.py:174-178
        ctx->err = "cledb_invert: bcalc 3 goes with IQUD and only with IQUD";
        return CLEDB_ERR_ARG;
    }
    if (mode == CLEDB_MODE_IQUD && !dopp) { ctx->err = "cledb_invert: IQUD needs the Doppler field strengths"; return CLEDB_ERR_ARG; }
    return CLEDB_OK;
}

// raw search outputs of the fused call live in their own workspace
struct RawBufs { int32_t* idx; float* chi; double* chi64; int32_t* kpos; float* rows; uint8_t* status; size_t bytes; };
static RawBufs carve_raw(unsigned char* base, size_t n, size_t k, bool want64) {
    RawBufs r;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_idx = take(n * k * 4), o_chi = take(n * k * 4), o_c64 = take(want64 ? n * k * 8 : 0), o_kp = take(n * k * 4),
                 o_rows = take(n * k * 32), o_st = take(n);
    r.idx = reinterpret_cast<int32_t*>(base + o_idx);
    r.chi = reinterpret_cast<float*>(base + o_chi);
    r.chi64 = want64 ? reinterpret_cast<double*>(base + o_c64) : nullptr;
    r.kpos = reinterpret_cast<int32_t*>(base + o_kp);
    r.rows = reinterpret_cast<float*>(base + o_rows);
    r.status = base + o_st;
    r.bytes = off;
    return r;
}

extern "C" int cledb_invert_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                                const int32_t* db_id, int nsearch, const float* yobs, const float* dobs, const float* rot_cos,
                                const float* rot_sin, const void* dopp, int dopp_is_f64, int64_t dopp_stride,
                                const cledb_dbgrid* grid, double maxchisq, int bcalc, int strict_topk, float* invout,
                                float* sfound, uint8_t* status_out, int32_t* issue_out, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    int rc = check_invert_args(ctx, mode, npix, sobs, snr, db_id, nsearch, yobs, dobs, rot_cos, rot_sin, dopp, grid, bcalc, invout, sfound);
    if (rc) return rc;
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const bool want64 = precision == CLEDB_PREC_FP64;
    RawBufs probe = carve_raw(nullptr, (size_t)npix, (size_t)nsearch, want64);
    if ((rc = ensure_workspace(ctx, ctx->raw, probe.bytes))) return rc;
    RawBufs raw = carve_raw(reinterpret_cast<unsigned char*>(ctx->raw.ptr), (size_t)npix, (size_t)nsearch, want64);
    uint8_t* status = status_out ? status_out : raw.status;
    rc = match_device(ctx, mode, precision, npix, sobs, snr, db_id, nsearch, raw.idx, raw.chi, raw.chi64, raw.kpos, raw.rows, status,
                      nullptr, st);
    if (rc) return rc;
    return launch_build(ctx, mode, precision, npix, nsearch, sobs, yobs, dobs, rot_cos, rot_sin, dopp, dopp_is_f64,
                        dopp_stride > 0 ? dopp_stride : 1, *grid, maxchisq, bcalc, strict_topk, raw.idx, raw.chi, raw.chi64, raw.kpos,
                        raw.rows, status, nullptr, invout, sfound, st, snr, issue_out);
}

extern "C" int cledb_invert(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                            const int32_t* db_id, int nsearch, const float* yobs, const float* dobs, const float* rot_cos,
                            const float* rot_sin, const void* dopp, int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid,
                            double maxchisq, int bcalc, int strict_topk, float* invout, float* sfound, uint8_t* status_out,
                            int32_t* issue_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    int rc = check_invert_args(ctx, mode, npix, sobs, snr, db_id, nsearch, yobs, dobs, rot_cos, rot_sin, dopp, grid, bcalc, invout, sfound);
    if (rc) return rc;
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix, k = (size_t)nsearch;
    const size_t dsz = dopp ? (dopp_is_f64 ? 8 : 4) : 0;
    if (dopp_stride <= 0) dopp_stride = 1;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_sobs = take(n * 32), o_snr = take(n * 32), o_id = take(n * 4), o_y = take(n * 4), o_d = take(n * 4),
                 o_rc = take(n * 4), o_rs = take(n * 4), o_dopp = take(n * dsz), o_inv = take(n * k * 44), o_sf = take(n * k * 32),
                 o_st = take(n), o_iss = take(n * 4);
    if ((rc = ensure_workspace(ctx, ctx->io, off))) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_sobs, sobs, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_snr, snr, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_id, db_id, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_y, yobs, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_d, dobs, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_rc, rot_cos, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_rs, rot_sin, n * 4, cudaMemcpyHostToDevice, st));
    const void* d_dopp = nullptr;
    if (dopp) {
        if (dopp_stride == 1) CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_dopp, dopp, n * dsz, cudaMemcpyHostToDevice, st));
        else CLEDB_CUDA(ctx, cudaMemcpy2DAsync(w + o_dopp, dsz, dopp, (size_t)dopp_stride * dsz, dsz, n, cudaMemcpyHostToDevice, st));
        d_dopp = w + o_dopp;
    }
    cledb_dbgrid g;
    if ((rc = grid_to_device(ctx, grid, &g, st))) return rc;
    // Big maps go through in chunks of 128 Ki pixels: the search of chunk c+1 runs while the finished maps of chunk c
    // travel to the host on a second stream (608 B per pixel at nsearch 8 is a quarter of the search time on PCIe 5).
    int64_t chunk = npix;
    if (!getenv("CLEDB_B200_NO_CHUNK")) {
        if (npix >= (int64_t)1 << 18) chunk = (int64_t)1 << 17;      // measured on 512^2 and 1024^2 maps; smaller maps lose more
                                                                      // to the per-chunk host round trip than the overlap gains
        if (const char* e = getenv("CLEDB_B200_CHUNK")) chunk = std::max<int64_t>(1024, std::atoll(e));
    }
    if (chunk < npix && !ctx->copy_stream) CLEDB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int64_t c0 = 0; c0 < npix; c0 += chunk) {
        const int64_t nc = std::min(chunk, npix - c0);
        const size_t o = (size_t)c0;
        rc = cledb_invert_dev(ctx, mode, precision, nc, reinterpret_cast<float*>(w + o_sobs) + o * 8, reinterpret_cast<float*>(w + o_snr) + o * 8,
                              reinterpret_cast<int32_t*>(w + o_id) + o, nsearch, reinterpret_cast<float*>(w + o_y) + o,
                              reinterpret_cast<float*>(w + o_d) + o, reinterpret_cast<float*>(w + o_rc) + o,
                              reinterpret_cast<float*>(w + o_rs) + o, d_dopp ? reinterpret_cast<const unsigned char*>(d_dopp) + o * dsz : nullptr,
                              dopp_is_f64, 1, &g, maxchisq, bcalc, strict_topk, reinterpret_cast<float*>(w + o_inv) + o * k * 11,
                              reinterpret_cast<float*>(w + o_sf) + o * k * 8, w + o_st + o, issue_out ? reinterpret_cast<int32_t*>(w + o_iss) + o : nullptr, st);
        if (rc) return rc;
        cudaStream_t cs = st;
        if (chunk < npix) {                                    // hand the chunk's results to the copy stream
            cudaEvent_t ev;
            CLEDB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CLEDB_CUDA(ctx, cudaEventRecord(ev, st));
            CLEDB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
            CLEDB_CUDA(ctx, cudaEventDestroy(ev));             // released once the wait has consumed it
            cs = ctx->copy_stream;
        }
        CLEDB_CUDA(ctx, cudaMemcpyAsync(invout + o * k * 11, w + o_inv + o * k * 44, (size_t)nc * k * 44, cudaMemcpyDeviceToHost, cs));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(sfound + o * k * 8, w + o_sf + o * k * 32, (size_t)nc * k * 32, cudaMemcpyDeviceToHost, cs));
        if (status_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(status_out + o, w + o_st + o, (size_t)nc, cudaMemcpyDeviceToHost, cs));
        if (issue_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(issue_out + o, w + o_iss + o * 4, (size_t)nc * 4, cudaMemcpyDeviceToHost, cs));
    }
    if (chunk < npix) CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return check_async_error(ctx);
}
