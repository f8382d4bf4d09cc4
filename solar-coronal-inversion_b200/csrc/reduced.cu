// Reduced-database mode of the 2-line search (reduced = True): cledb_getsubsetiquv / cledb_getsubsetiqud + the search over
// the per-pixel subset (CLEDB_PROC.py:1401-1568, 951-963, 1125-1134 of the reference).
//
// The reference builds, for every pixel, a private copy of 2*nsearch theta rows per phi (the nsearch rows whose
// tan(theta)*sin(phi) is closest to the observed tan(Phi_B), and the nsearch closest to the degenerate branch), for all
// gx, and searches that copy.  Here one warp owns one pixel: it derives the theta lists in shared memory (phase 1) and
// then ENUMERATES the subset in the reference's order without materialising it (phase 2): position q of the subset maps
// to one row of the resident, decoded database (a 32-byte gather that hits L2), duplicates and the all-zero rows of the
// phi = 0 plane included, so positions, ties and the cledb_partsort quirk see exactly the reference's list.
#include <math_constants.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "cledb_internal.h"

namespace cledb {

constexpr int kRedWarps = 4;

struct RedParams {
    int mode, k;
    int outer, nbphi, nbtheta;
    int64_t npix;
    const float* sobs;
    const float* snr;
    const int32_t* db_id;
    const DbSlot* slots;
    const int32_t* slot_of_id;
    int id_cap;
    const float* const* rows8;     // [slots] decoded [n,8] databases
    const double* tan_btheta;      // [nbtheta]
    const double* sin_bphi;        // [nbphi]
    const int32_t* tie_rank;       // [nbtheta]
    const double* tan_phib;        // [npix,2]
    int32_t* idx;
    float* chi;
    double* chi64;
    int32_t* kpos;
    float* rows;
    uint8_t* status;
    int32_t* nan_slots;
    int32_t* error;
};

struct RedCand { double v; int32_t kpos; int32_t q; };

__device__ __forceinline__ int sgn(float x) { return x > 0.f ? 1 : (x < 0.f ? -1 : (x == 0.f ? 0 : 2)); }   // 2 = NaN: equals nothing

template <int PREC>
__global__ void __launch_bounds__(kRedWarps * 32) k_reduced(RedParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * kRedWarps + warp;
    if (p >= P.npix) return;
    const int k = P.k, k2 = 2 * k;
    const size_t per_warp = ((size_t)P.nbphi * k2 * 2 + 15) / 16 * 16 + (size_t)CLEDB_MAX_NSEARCH * sizeof(RedCand);
    uint16_t* lists = reinterpret_cast<uint16_t*>(smem + warp * per_warp);                                  // [nbphi][2k]
    RedCand* top = reinterpret_cast<RedCand*>(smem + warp * per_warp + ((size_t)P.nbphi * k2 * 2 + 15) / 16 * 16);   // [k]

    auto finish = [&](int status, int nfound, int nans) {
        for (int i = lane; i < k; i += 32) {
            const int64_t o = p * k + i;
            if (i >= nfound) {
                P.idx[o] = -1;
                P.chi[o] = 0.f;
                if (P.chi64) P.chi64[o] = 0.0;
                P.kpos[o] = -1;
                for (int c = 0; c < 8; ++c) P.rows[o * 8 + c] = 0.f;
            }
        }
        if (lane == 0) { P.status[p] = (uint8_t)status; P.nan_slots[p] = nans; }
    };

    // ---- the pixel: null tests and normalisation exactly as in the full search (k_classify)
    float s[8], n[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { s[c] = P.sobs[p * 8 + c]; n[c] = P.snr[p * 8 + c]; }
    int nnan = 0, nzero = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { nnan += isnan(s[c]) ? 1 : 0; nzero += s[c] == 0.f ? 1 : 0; }
    const bool null_pix = P.mode == CLEDB_MODE_IQUV ? (nnan > 0 || nzero == 8) : (nnan == 8 || nzero == 8);   // :944 / :1118
    if (null_pix) { finish(CLEDB_PIX_NULL, 0, 0); return; }
    const int id = P.db_id[p];
    const int slot = (id >= 0 && id < P.id_cap) ? P.slot_of_id[id] : -1;
    if (slot < 0 || !P.slots[slot].live || P.slots[slot].n != (int64_t)P.outer * P.nbphi * P.nbtheta) {
        if (lane == 0) atomicExch(P.error, slot < 0 ? 1 : 2);
        finish(CLEDB_PIX_NULL, 0, 0);
        return;
    }
    if (P.mode == CLEDB_MODE_IQUD && nnan > 0) { finish(CLEDB_PIX_REFRAISES, 0, 0); return; }                  // np.max is NaN
    float nf = s[0];
    if (P.mode == CLEDB_MODE_IQUD) {
#pragma unroll
        for (int c = 1; c < 8; ++c) nf = fmaxf(nf, s[c]);
    }
    float st8[8], sig8[8];
    bool allnan = false, degenerate = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        st8[c] = __fdiv_rn(s[c], nf);
        sig8[c] = __fdiv_rn(st8[c], n[c]);
        if (!isfinite(st8[c]) || isnan(sig8[c])) allnan = true;
    }
    if (sig8[3] == 0.f || sig8[7] == 0.f) allnan = true;             // x/0 times weight 0 is NaN in every row
    if (allnan) { finish(CLEDB_PIX_ALLNAN, 0, 0); return; }
    float st[kNC], sig[kNC], a[kNC], nb[kNC];
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const int oc = chi_comp(c);
        st[c] = st8[oc];
        sig[c] = sig8[oc];
        if (sig[c] == 0.f || !isfinite(sig[c])) degenerate = true;   // +-inf / NaN per row: evaluate in the reference's order
        const double w = oc == 4 ? 0.01 : 1.0;
        a[c] = (float)(w / (double)sig[c]);
        nb[c] = (float)(-(double)a[c] * (double)st[c]);
    }
    if (sig8[0] == 0.f || !isfinite(sig8[0])) degenerate = true;
    // component 0 of a database row is 1, of an all-zero row 0
    const float r0_one = __fdiv_rn(__fsub_rn(1.0f, st8[0]), sig8[0]), r0_zero = __fdiv_rn(__fsub_rn(0.0f, st8[0]), sig8[0]);
    const double c0_one = (double)r0_one * (double)r0_one, c0_zero = (double)r0_zero * (double)r0_zero;

    // ---- phase 1: theta lists.  Lane ir handles phi index ir, ir+32, ...; for each branch nsearch rounds of "smallest
    //      (difference, tie rank) above the previous one" (NumPy argsort order; NaN differences sort last)
    const double ta = P.tan_phib[2 * p], tb = P.tan_phib[2 * p + 1];
    for (int ir = lane; ir < P.nbphi; ir += 32) {
        const double sp = P.sin_bphi[ir];
        for (int br = 0; br < 2; ++br) {
            const double t = br ? tb : ta;
            double pd = -1.0;
            int pr = -1;
            for (int r = 0; r < k; ++r) {
                double bd = CUDART_INF;
                int brk = 0x7fffffff, bl = 0;
                for (int l = 0; l < P.nbtheta; ++l) {
                    double d = fabs(__dsub_rn(t, __dmul_rn(P.tan_btheta[l], sp)));
                    if (isnan(d)) d = CUDART_INF;
                    const int rk = P.tie_rank[l];
                    const bool above = d > pd || (d == pd && rk > pr);
                    const bool below = d < bd || (d == bd && rk < brk);
                    if (above && below) { bd = d; brk = rk; bl = l; }
                }
                lists[ir * k2 + br * k + r] = (uint16_t)bl;
                pd = bd;
                pr = brk;
            }
        }
    }
    __syncwarp();
    const bool zero0 = lists[0] == 0 && lists[k] == 0;               // CLEDB_PROC.py:1446: the phi = 0 plane stays zero

    // ---- phase 2: enumerate the subset in the reference's order
    const float* __restrict__ db8 = P.rows8[slot];
    const int64_t M = (int64_t)P.outer * P.nbphi * k2;
    const int ssign = sgn(s[3]);
    int running = 0, nans = 0, count = 0;
    double tau = CUDART_INF;
    for (int i = lane; i < k; i += 32) { top[i].v = CUDART_INF; top[i].kpos = -1; top[i].q = -1; }
    __syncwarp();
    for (int64_t base = 0; base < M; base += 32) {
        const int64_t q = base + lane;
        const bool inside = q < M;
        float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool zr = false;
        if (inside) {
            const int o = (int)(q / ((int64_t)P.nbphi * k2));
            const int rem = (int)(q - (int64_t)o * P.nbphi * k2);
            const int ir = rem / k2, jj = rem - ir * k2;
            zr = ir == 0 && zero0;
            if (!zr) {
                const int64_t r = ((int64_t)o * P.nbphi + ir) * P.nbtheta + lists[ir * k2 + jj];
                const float4 x = reinterpret_cast<const float4*>(db8)[2 * r], y = reinterpret_cast<const float4*>(db8)[2 * r + 1];
                row[0] = x.x; row[1] = x.y; row[2] = x.z; row[3] = x.w; row[4] = y.x; row[5] = y.y; row[6] = y.z; row[7] = y.w;
            }
        }
        const bool cand = inside && (P.mode == CLEDB_MODE_IQUV ? sgn(row[3]) == ssign : row[3] == row[3]);   // :959 / :1131
        const uint32_t bal = __ballot_sync(0xffffffffu, cand);
        const int mypos = running + __popc(bal & ((1u << lane) - 1u));
        running += __popc(bal);
        double v = CUDART_INF;
        bool isn = false;
        if (cand) {
            const float d[kNC] = {row[1], row[2], row[4], row[5], row[6]};
            if (PREC == CLEDB_PREC_FP64 || degenerate) v = eval_fp64(d, st, sig, zr ? c0_zero : c0_one);
            else v = (double)(0.5f * eval_fp32(d, a, nb, (float)(zr ? c0_zero : c0_one)));
            isn = isnan(v);
            if (isn) v = CUDART_INF;
        }
        nans += __popc(__ballot_sync(0xffffffffu, isn));
        uint32_t hits = __ballot_sync(0xffffffffu, cand && !isn && (v < tau || count < k));
        while (hits) {
            const int src = __ffs(hits) - 1;
            hits &= hits - 1;
            const double cv = __shfl_sync(0xffffffffu, v, src);
            const int cp = __shfl_sync(0xffffffffu, mypos, src);
            const int cq = __shfl_sync(0xffffffffu, (int)(base + src), src);
            if (!(cv < tau || count < k)) continue;                   // tau moved while this chunk was being inserted
            // sorted insertion, stable: a new value goes behind equal ones (np.argmin = first occurrence)
            double lv = CUDART_INF;
            int lk = -1, lq = -1;
            if (lane < k) { lv = top[lane].v; lk = top[lane].kpos; lq = top[lane].q; }
            const bool before = lane < count && lv <= cv;
            const int at = __popc(__ballot_sync(0xffffffffu, before));
            const double uv = __shfl_up_sync(0xffffffffu, lv, 1);
            const int uk = __shfl_up_sync(0xffffffffu, lk, 1), uq = __shfl_up_sync(0xffffffffu, lq, 1);
            if (at < k) {
                if (lane == at) { lv = cv; lk = cp; lq = cq; }
                else if (lane > at) { lv = uv; lk = uk; lq = uq; }
                if (lane >= at && lane < k) { top[lane].v = lv; top[lane].kpos = lk; top[lane].q = lq; }
                if (count < k) ++count;
            }
            __syncwarp();
            tau = count == k ? top[k - 1].v : CUDART_INF;
        }
    }
    // ---- what the reference would have done with this list
    if (running == 0 || running < k) { finish(CLEDB_PIX_REFRAISES, 0, 0); return; }       // argmin of an empty slice raises
    const int nn = min(nans, k);                                      // NaN chi^2 come first in the reference's selection sort
    const int nfound = min(count, k - nn);
    for (int i = lane; i < nfound; i += 32) {
        const RedCand c = top[i];
        const int o = (int)(c.q / ((int64_t)P.nbphi * k2));
        const int rem = (int)(c.q - (int64_t)o * P.nbphi * k2);
        const int ir = rem / k2, jj = rem - ir * k2;
        const bool zr = ir == 0 && zero0;
        const int th = zr ? 0 : lists[ir * k2 + jj];                  // outredindex keeps 0 for the zero plane
        const int64_t r = ((int64_t)o * P.nbphi + ir) * P.nbtheta + th;
        const int64_t oo = p * k + i;
        P.idx[oo] = (int32_t)r;
        P.chi[oo] = (float)c.v;
        if (P.chi64) P.chi64[oo] = c.v;
        P.kpos[oo] = c.kpos;
        for (int cc = 0; cc < 8; ++cc) P.rows[oo * 8 + cc] = zr ? 0.f : db8[r * 8 + cc];
    }
    finish(CLEDB_PIX_SEARCHED, nfound, nn);
}

// decoded [n,8] copy of a resident database, built on first use of the reduced mode
static int ensure_rows8(cledb_ctx* ctx, int slot, cudaStream_t st) {
    HostDb& h = ctx->dbs[slot];
    if (h.d_rows8) return CLEDB_OK;
    if (cudaMalloc(&h.d_rows8, (size_t)h.slot.n * 32) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "reduced search: out of device memory for the decoded database";
        return CLEDB_ERR_NOMEM;
    }
    return launch_decode_db(ctx, slot, h.d_rows8, st);
}

int reduced_device(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr, const int32_t* db_id,
                   int nsearch, const cledb_reduced& red, int32_t* idx, float* chi, double* chi64, int32_t* kpos, float* rows,
                   uint8_t* status, int32_t* nan_slots, cudaStream_t st) {
    if (red.outer < 1 || red.nbphi < 1 || red.nbtheta < 1 || red.nbtheta > 65535 || !red.tan_btheta || !red.sin_bphi || !red.tie_rank ||
        !red.tan_phib) {
        ctx->err = "reduced search: incomplete cledb_reduced description";
        return CLEDB_ERR_ARG;
    }
    if (nsearch > red.nbtheta) { ctx->err = "reduced search: nsearch exceeds the number of theta entries"; return CLEDB_ERR_ARG; }
    int rc = sync_tables(ctx, st);
    if (rc) return rc;
    const int nslots = (int)ctx->dbs.size();
    std::vector<const float*> h_rows8(std::max(1, nslots), nullptr);
    for (int sl = 0; sl < nslots; ++sl) {
        if (!ctx->dbs[sl].slot.live) continue;
        if (ctx->dbs[sl].slot.n != (int64_t)red.outer * red.nbphi * red.nbtheta) continue;     // pixels using it are reported
        if ((rc = ensure_rows8(ctx, sl, st))) return rc;
        h_rows8[sl] = ctx->dbs[sl].d_rows8;
    }
    const size_t tab_bytes = h_rows8.size() * sizeof(float*) + 16;
    if ((rc = ensure_workspace(ctx, ctx->red, tab_bytes))) return rc;
    int32_t* d_err = reinterpret_cast<int32_t*>(ctx->red.ptr);
    const float** d_rows8 = reinterpret_cast<const float**>(reinterpret_cast<unsigned char*>(ctx->red.ptr) + 16);
    CLEDB_CUDA(ctx, cudaMemsetAsync(d_err, 0, 16, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(d_rows8, h_rows8.data(), h_rows8.size() * sizeof(float*), cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));          // h_rows8 is a stack object
    RedParams P;
    P.mode = mode; P.k = nsearch; P.outer = red.outer; P.nbphi = red.nbphi; P.nbtheta = red.nbtheta; P.npix = npix;
    P.sobs = sobs; P.snr = snr; P.db_id = db_id; P.slots = ctx->d_slots; P.slot_of_id = ctx->d_slot_of_id; P.id_cap = ctx->id_cap;
    P.rows8 = d_rows8; P.tan_btheta = red.tan_btheta; P.sin_bphi = red.sin_bphi; P.tie_rank = red.tie_rank; P.tan_phib = red.tan_phib;
    P.idx = idx; P.chi = chi; P.chi64 = chi64; P.kpos = kpos; P.rows = rows; P.status = status; P.nan_slots = nan_slots; P.error = d_err;
    const size_t per_warp = ((size_t)red.nbphi * 2 * nsearch * 2 + 15) / 16 * 16 + (size_t)CLEDB_MAX_NSEARCH * sizeof(RedCand);
    const size_t smem = per_warp * kRedWarps;
    if (smem > 200 * 1024) { ctx->err = "reduced search: nbphi * nsearch too large for shared memory"; return CLEDB_ERR_ARG; }
    const int grid = (int)((npix + kRedWarps - 1) / kRedWarps);
    if (precision == CLEDB_PREC_FP64) {
        CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_reduced<CLEDB_PREC_FP64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_reduced<CLEDB_PREC_FP64><<<grid, kRedWarps * 32, smem, st>>>(P);
    } else {
        CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_reduced<CLEDB_PREC_FP32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_reduced<CLEDB_PREC_FP32><<<grid, kRedWarps * 32, smem, st>>>(P);
    }
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    int32_t herr = 0;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    if (herr == 1) { ctx->err = "reduced search: a pixel refers to a database id that was never put"; return CLEDB_ERR_NODB; }
    if (herr == 2) { ctx->err = "reduced search: a database does not have outer*nbphi*nbtheta rows"; return CLEDB_ERR_ARG; }
    return CLEDB_OK;
}

}  // namespace cledb

using namespace cledb;

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

// Host-buffer entry point shared by the raw and the fused call: uploads, runs the reduced search and (if grid != NULL)
// the solution building, downloads.
static int reduced_host(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                        const int32_t* db_id, int nsearch, const cledb_reduced* red, int32_t* idx_out, float* chi_out,
                        double* chi64_out, int32_t* kpos_out, float* rows_out, uint8_t* status_out, int32_t* nan_out,
                        const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin, const void* dopp,
                        int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid, double maxchisq, int bcalc, int strict_topk,
                        float* invout, float* sfound, int32_t* issue_out = nullptr) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || !sobs || !snr || !db_id || !red) { ctx->err = "reduced search: null argument"; return CLEDB_ERR_ARG; }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "reduced search: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (mode != CLEDB_MODE_IQUV && mode != CLEDB_MODE_IQUD) { ctx->err = "reduced search: unknown mode"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix, k = (size_t)nsearch;
    const bool want64 = precision == CLEDB_PREC_FP64;
    const size_t dsz = (grid && dopp) ? (dopp_is_f64 ? 8 : 4) : 0;
    if (dopp_stride <= 0) dopp_stride = 1;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_sobs = take(n * 32), o_snr = take(n * 32), o_id = take(n * 4), o_tp = take(n * 16),
                 o_tt = take((size_t)red->nbtheta * 8), o_sp = take((size_t)red->nbphi * 8), o_tr = take((size_t)red->nbtheta * 4),
                 o_idx = take(n * k * 4), o_chi = take(n * k * 4), o_c64 = take(want64 ? n * k * 8 : 0), o_kp = take(n * k * 4),
                 o_rows = take(n * k * 32), o_st = take(n), o_nan = take(n * 4),
                 o_y = take(grid ? n * 4 : 0), o_d = take(grid ? n * 4 : 0), o_rc = take(grid ? n * 4 : 0), o_rs = take(grid ? n * 4 : 0),
                 o_dopp = take(n * dsz), o_tab = take(grid ? (size_t)(2 * grid->nbphi + 2 * grid->nbtheta) * 4 : 0),
                 o_inv = take(grid ? n * k * 44 : 0), o_sf = take(grid ? n * k * 32 : 0),
                 o_iss = take(grid && issue_out ? n * 4 : 0);
    int rc = ensure_workspace(ctx, ctx->io, off);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_sobs, sobs, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_snr, snr, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_id, db_id, n * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_tp, red->tan_phib, n * 16, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_tt, red->tan_btheta, (size_t)red->nbtheta * 8, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_sp, red->sin_bphi, (size_t)red->nbphi * 8, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_tr, red->tie_rank, (size_t)red->nbtheta * 4, cudaMemcpyHostToDevice, st));
    cledb_reduced dr = *red;
    dr.tan_phib = reinterpret_cast<const double*>(w + o_tp);
    dr.tan_btheta = reinterpret_cast<const double*>(w + o_tt);
    dr.sin_bphi = reinterpret_cast<const double*>(w + o_sp);
    dr.tie_rank = reinterpret_cast<const int32_t*>(w + o_tr);
    int32_t* d_idx = reinterpret_cast<int32_t*>(w + o_idx);
    float* d_chi = reinterpret_cast<float*>(w + o_chi);
    double* d_c64 = want64 ? reinterpret_cast<double*>(w + o_c64) : nullptr;
    int32_t* d_kp = reinterpret_cast<int32_t*>(w + o_kp);
    float* d_rows = reinterpret_cast<float*>(w + o_rows);
    int32_t* d_nan = reinterpret_cast<int32_t*>(w + o_nan);
    rc = reduced_device(ctx, mode, precision, npix, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_snr),
                        reinterpret_cast<int32_t*>(w + o_id), nsearch, dr, d_idx, d_chi, d_c64, d_kp, d_rows, w + o_st, d_nan, st);
    if (rc) return rc;
    if (grid) {
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
        cledb_dbgrid g = *grid;
        float* tab = reinterpret_cast<float*>(w + o_tab);
        CLEDB_CUDA(ctx, cudaMemcpyAsync(tab, grid->sin_bphi, (size_t)g.nbphi * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + g.nbphi, grid->cos_bphi, (size_t)g.nbphi * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + 2 * g.nbphi, grid->sin_btheta, (size_t)g.nbtheta * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + 2 * g.nbphi + g.nbtheta, grid->cos_btheta, (size_t)g.nbtheta * 4, cudaMemcpyHostToDevice, st));
        g.sin_bphi = tab; g.cos_bphi = tab + g.nbphi; g.sin_btheta = tab + 2 * g.nbphi; g.cos_btheta = tab + 2 * g.nbphi + g.nbtheta;
        rc = launch_build(ctx, mode, precision, npix, nsearch, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_y),
                          reinterpret_cast<float*>(w + o_d), reinterpret_cast<float*>(w + o_rc), reinterpret_cast<float*>(w + o_rs), d_dopp,
                          dopp_is_f64, 1, g, maxchisq, bcalc, strict_topk, d_idx, d_chi, d_c64, d_kp, d_rows, w + o_st, d_nan,
                          reinterpret_cast<float*>(w + o_inv), reinterpret_cast<float*>(w + o_sf), st, reinterpret_cast<float*>(w + o_snr),
                          issue_out ? reinterpret_cast<int32_t*>(w + o_iss) : nullptr);
        if (rc) return rc;
        if (issue_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(issue_out, w + o_iss, n * 4, cudaMemcpyDeviceToHost, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(invout, w + o_inv, n * k * 44, cudaMemcpyDeviceToHost, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(sfound, w + o_sf, n * k * 32, cudaMemcpyDeviceToHost, st));
    }
    if (idx_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(idx_out, d_idx, n * k * 4, cudaMemcpyDeviceToHost, st));
    if (chi_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(chi_out, d_chi, n * k * 4, cudaMemcpyDeviceToHost, st));
    if (chi64_out && want64) CLEDB_CUDA(ctx, cudaMemcpyAsync(chi64_out, d_c64, n * k * 8, cudaMemcpyDeviceToHost, st));
    if (kpos_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(kpos_out, d_kp, n * k * 4, cudaMemcpyDeviceToHost, st));
    if (rows_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(rows_out, d_rows, n * k * 32, cudaMemcpyDeviceToHost, st));
    if (status_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(status_out, w + o_st, n, cudaMemcpyDeviceToHost, st));
    if (nan_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(nan_out, d_nan, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return CLEDB_OK;
}

extern "C" int cledb_match_reduced(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                                   const int32_t* db_id, int nsearch, const cledb_reduced* red, int32_t* idx_out, float* chi_out,
                                   double* chi64_out, int32_t* kpos_out, float* rows_out, uint8_t* status_out, int32_t* nan_slots_out) {
    if (ctx && (!idx_out || !chi_out)) { ctx->err = "cledb_match_reduced: null argument"; return CLEDB_ERR_ARG; }
    return reduced_host(ctx, mode, precision, npix, sobs, snr, db_id, nsearch, red, idx_out, chi_out, chi64_out, kpos_out, rows_out,
                        status_out, nan_slots_out, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 1, nullptr, 0.0, 0, 0, nullptr, nullptr);
}

// device-pointer variant of the raw reduced search: every array - the four tables inside *red included - lives on the
// context's device; enqueued on the stream, no copies, no wait
extern "C" int cledb_match_reduced_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                                       const int32_t* db_id, int nsearch, const cledb_reduced* red, int32_t* idx_out, float* chi_out,
                                       double* chi64_out, int32_t* kpos_out, float* rows_out, uint8_t* status_out, int32_t* nan_slots_out,
                                       void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || !sobs || !snr || !db_id || !red || !idx_out || !chi_out || !kpos_out || !rows_out || !status_out || !nan_slots_out) {
        ctx->err = "cledb_match_reduced_dev: null argument (every output is required)";
        return CLEDB_ERR_ARG;
    }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "reduced search: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (mode != CLEDB_MODE_IQUV && mode != CLEDB_MODE_IQUD) { ctx->err = "reduced search: unknown mode"; return CLEDB_ERR_ARG; }
    if (precision == CLEDB_PREC_FP64 && !chi64_out) { ctx->err = "cledb_match_reduced_dev: the fp64 search needs chi64_out"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    return reduced_device(ctx, mode, precision, npix, sobs, snr, db_id, nsearch, *red, idx_out, chi_out,
                          precision == CLEDB_PREC_FP64 ? chi64_out : nullptr, kpos_out, rows_out, status_out, nan_slots_out,
                          stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int cledb_invert_reduced(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                                    const int32_t* db_id, int nsearch, const cledb_reduced* red, const float* yobs, const float* dobs,
                                    const float* rot_cos, const float* rot_sin, const void* dopp, int dopp_is_f64, int64_t dopp_stride,
                                    const cledb_dbgrid* grid, double maxchisq, int bcalc, int strict_topk, float* invout, float* sfound,
                                    uint8_t* status_out, int32_t* issue_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!yobs || !dobs || !rot_cos || !rot_sin || !grid || !invout || !sfound) { ctx->err = "cledb_invert_reduced: null argument"; return CLEDB_ERR_ARG; }
    if (grid->dbtype != 1) { ctx->err = "cledb_invert_reduced: only PyCELP-type databases (dbtype 1) are built on the device"; return CLEDB_ERR_ARG; }
    if (bcalc < 0 || bcalc > 3 || (mode == CLEDB_MODE_IQUD) != (bcalc == 3)) { ctx->err = "cledb_invert_reduced: bcalc 3 goes with IQUD and only with IQUD"; return CLEDB_ERR_ARG; }
    if (mode == CLEDB_MODE_IQUD && !dopp) { ctx->err = "cledb_invert_reduced: IQUD needs the Doppler field strengths"; return CLEDB_ERR_ARG; }
    return reduced_host(ctx, mode, precision, npix, sobs, snr, db_id, nsearch, red, nullptr, nullptr, nullptr, nullptr, nullptr, status_out,
                        nullptr, yobs, dobs, rot_cos, rot_sin, dopp, dopp_is_f64, dopp_stride, grid, maxchisq, bcalc, strict_topk, invout,
                        sfound, issue_out);
}
