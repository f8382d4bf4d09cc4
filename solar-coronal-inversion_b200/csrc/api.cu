// C-ABI entry points of libcledb_b200.so (include/cledb_b200.h): context management and the host-buffer
// wrappers around the device paths.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cledb_internal.h"

using namespace cledb;

static thread_local std::string g_create_error;

extern "C" const char* cledb_version(void) { return "cledb_b200 0.1.0 (sm_100a)"; }

extern "C" const char* cledb_last_error(cledb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int cledb_create(int device, cledb_ctx** out) {
    if (!out) return CLEDB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("cledb_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        cudaGetLastError();
        return CLEDB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "cledb_create: device ordinal out of range"; return CLEDB_ERR_ARG; }
    cledb_ctx* ctx = new cledb_ctx();
    ctx->device = device;
    if (const char* e = std::getenv("CLEDB_B200_PRUNING")) ctx->prune = std::max(0, std::min(2, std::atoi(e)));   // default: 1
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cledb_create: ") + cudaGetErrorString(e);
        delete ctx;
        return CLEDB_ERR_CUDA;
    }
    if (ctx->prop.major < 10) {
        g_create_error = "cledb_create: device is sm_" + std::to_string(ctx->prop.major * 10 + ctx->prop.minor) +
                         "; this library is built for sm_100a (B200) only";
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return CLEDB_ERR_CUDA;
    }
    *out = ctx;
    return CLEDB_OK;
}

extern "C" int cledb_destroy(cledb_ctx* ctx) {
    if (!ctx) return CLEDB_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cledb_comm_destroy(ctx);
    std::vector<int> ids;
    for (auto& kv : ctx->slot_of_id) ids.push_back(kv.first);
    for (int id : ids) cledb_db_drop(ctx, id);
    if (ctx->d_slots) cudaFree(ctx->d_slots);
    if (ctx->d_slot_of_id) cudaFree(ctx->d_slot_of_id);
    if (ctx->ws.ptr) cudaFree(ctx->ws.ptr);
    if (ctx->ws2.ptr) cudaFree(ctx->ws2.ptr);
    if (ctx->raw.ptr) cudaFree(ctx->raw.ptr);
    if (ctx->red.ptr) cudaFree(ctx->red.ptr);
    if (ctx->d_prune_stats) cudaFree(ctx->d_prune_stats);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->d_tab) cudaFree(ctx->d_tab);
    if (ctx->io.ptr) cudaFree(ctx->io.ptr);
    for (auto& pr : ctx->prof) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return CLEDB_OK;
}

extern "C" int cledb_device_props(cledb_ctx* ctx, int64_t props[8]) {
    if (!ctx || !props) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->occ == 0) { int rc = match_occupancy(ctx); if (rc) return rc; }
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
    props[0] = ctx->prop.multiProcessorCount;
    props[1] = clock_khz;
    props[2] = (int64_t)ctx->prop.totalGlobalMem;
    props[3] = ctx->prop.l2CacheSize;
    props[4] = ctx->prop.major * 10 + ctx->prop.minor;
    props[5] = ctx->occ;
    props[6] = ctx->regs;
    props[7] = ctx->smem;
    return CLEDB_OK;
}

extern "C" int cledb_synchronize(cledb_ctx* ctx) {
    if (!ctx) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    CLEDB_CUDA(ctx, cudaDeviceSynchronize());
    return check_async_error(ctx);
}

extern "C" int64_t cledb_launch_count(cledb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cledb_profile_begin(cledb_ctx* ctx) {
    if (!ctx) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    while (ctx->prof.size() < 256) {
        cudaEvent_t a, b;
        CLEDB_CUDA(ctx, cudaEventCreate(&a));
        CLEDB_CUDA(ctx, cudaEventCreate(&b));
        ctx->prof.push_back({a, b});
    }
    ctx->prof_used = 0;
    ctx->prof_on = true;
    return CLEDB_OK;
}

extern "C" int cledb_profile_collect(cledb_ctx* ctx, float* ms, int cap) {
    if (!ctx || !ms) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    CLEDB_CUDA(ctx, cudaDeviceSynchronize());
    int n = 0;
    for (int i = 0; i < ctx->prof_used && n < cap; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->prof[i].first, ctx->prof[i].second) == cudaSuccess) ms[n++] = t;
        else cudaGetLastError();
    }
    ctx->prof_on = false;
    return n;
}

extern "C" int cledb_measure_fp32_peak(cledb_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    return measure_fp32_peak(ctx, tflops);
}

// pinned (page-locked) host memory for the binding's staging buffers: copies to and from it run at full PCIe speed and
// asynchronously; pageable memory goes through the driver's bounce buffers at about half of that
extern "C" int cledb_host_alloc(int64_t bytes, void** out) {
    if (!out || bytes <= 0) return CLEDB_ERR_ARG;
    *out = nullptr;
    if (cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return CLEDB_ERR_NOMEM;
    }
    return CLEDB_OK;
}
extern "C" int cledb_host_free(void* p) {
    if (!p) return CLEDB_OK;
    return cudaFreeHost(p) == cudaSuccess ? CLEDB_OK : CLEDB_ERR_CUDA;
}

extern "C" int cledb_set_tuning(cledb_ctx* ctx, int pixels_per_item, int split) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (pixels_per_item < 0) {   // undocumented: timing experiments (results invalid), see Plan::dbg
        ctx->dbg = -pixels_per_item;
        return CLEDB_OK;
    }
    ctx->dbg = 0;
    if (pixels_per_item > kMaxP || split < 0 || split > 8) { ctx->err = "cledb_set_tuning: out of range"; return CLEDB_ERR_ARG; }
    ctx->pixels_per_item = pixels_per_item;
    ctx->split = split;
    return CLEDB_OK;
}

extern "C" int cledb_set_pruning(cledb_ctx* ctx, int on) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (on < 0 || on > 2) { ctx->err = "cledb_set_pruning: 0 (plain layout, brute force), 1 (exact tile pruning) or 2 (k-d layout, seeded brute force)"; return CLEDB_ERR_ARG; }
    ctx->prune = on;
    return CLEDB_OK;
}

extern "C" int cledb_pruning_stats(cledb_ctx* ctx, int64_t evaluated[2]) {
    if (!ctx || !evaluated) return CLEDB_ERR_ARG;
    evaluated[0] = evaluated[1] = 0;
    if (!ctx->d_prune_stats) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    CLEDB_CUDA(ctx, cudaDeviceSynchronize());
    unsigned long long h[2] = {0, 0};
    CLEDB_CUDA(ctx, cudaMemcpy(h, ctx->d_prune_stats, 16, cudaMemcpyDeviceToHost));
    evaluated[0] = (int64_t)h[0];
    evaluated[1] = (int64_t)h[1];
    return CLEDB_OK;
}

// timing experiments only (not in the header): raw words of the pruned search's statistics buffer
extern "C" int cledb_debug_stats(cledb_ctx* ctx, unsigned long long* out, int n) {
    if (!ctx || !out || n < 1 || n > (1 << 17) || !ctx->d_prune_stats) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    CLEDB_CUDA(ctx, cudaDeviceSynchronize());
    CLEDB_CUDA(ctx, cudaMemcpy(out, ctx->d_prune_stats, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return CLEDB_OK;
}

// ---------------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------------
extern "C" int cledb_match_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                               const int32_t* db_id, int nsearch, int32_t* idx_out, float* chi_out, double* chi64_out,
                               int32_t* kpos_out, float* rows_out, uint8_t* status_out, int64_t* ncand_out, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    return match_device(ctx, mode, precision, npix, sobs, snr, db_id, nsearch, idx_out, chi_out, chi64_out, kpos_out, rows_out,
                        status_out, ncand_out, st);
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int cledb_match(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                           const int32_t* db_id, int nsearch, int32_t* idx_out, float* chi_out, double* chi64_out,
                           int32_t* kpos_out, float* rows_out, uint8_t* status_out, int64_t* ncand_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || !sobs || !snr || !db_id || !idx_out || !chi_out) { ctx->err = "cledb_match: null argument"; return CLEDB_ERR_ARG; }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "cledb_match: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t k = (size_t)nsearch, n = (size_t)npix;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_sobs = take(n * 32), o_snr = take(n * 32), o_id = take(n * 4);
    const size_t o_idx = take(n * k * 4), o_chi = take(n * k * 4), o_chi64 = take(chi64_out ? n * k * 8 : 0);
    const size_t o_kpos = take(kpos_out ? n * k * 4 : 0), o_rows = take(rows_out ? n * k * 32 : 0);
    const size_t o_st = take(n), o_nc = take(ncand_out ? n * 8 : 0);
    int rc = ensure_workspace(ctx, ctx->io, off);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_sobs, sobs, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_snr, snr, n * 32, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_id, db_id, n * 4, cudaMemcpyHostToDevice, st));
    rc = match_device(ctx, mode, precision, npix, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_snr),
                      reinterpret_cast<int32_t*>(w + o_id), nsearch, reinterpret_cast<int32_t*>(w + o_idx),
                      reinterpret_cast<float*>(w + o_chi), chi64_out ? reinterpret_cast<double*>(w + o_chi64) : nullptr,
                      kpos_out ? reinterpret_cast<int32_t*>(w + o_kpos) : nullptr,
                      rows_out ? reinterpret_cast<float*>(w + o_rows) : nullptr, w + o_st,
                      ncand_out ? reinterpret_cast<int64_t*>(w + o_nc) : nullptr, st);
    if (rc) return rc;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(idx_out, w + o_idx, n * k * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(chi_out, w + o_chi, n * k * 4, cudaMemcpyDeviceToHost, st));
    if (chi64_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(chi64_out, w + o_chi64, n * k * 8, cudaMemcpyDeviceToHost, st));
    if (kpos_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(kpos_out, w + o_kpos, n * k * 4, cudaMemcpyDeviceToHost, st));
    if (rows_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(rows_out, w + o_rows, n * k * 32, cudaMemcpyDeviceToHost, st));
    if (status_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(status_out, w + o_st, n, cudaMemcpyDeviceToHost, st));
    if (ncand_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(ncand_out, w + o_nc, n * 8, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return check_async_error(ctx);
}

// ---------------------------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------------------------
extern "C" int cledb_blos_dev(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac,
                              float* out, uint8_t* flags, void* stream) {
    if (!ctx || npix < 0 || !iquv || !out || !flags) { if (ctx) ctx->err = "cledb_blos: bad argument"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_blos(ctx, npix, iquv, kfac, g_eff, ffac, out, flags, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int cledb_blos(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac, float* out,
                          uint8_t* flags) {
    if (!ctx || npix < 0 || !iquv || !out || !flags) { if (ctx) ctx->err = "cledb_blos: bad argument"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix;
    int rc = ensure_workspace(ctx, ctx->io, up256(n * 16) * 2 + up256(n * 2));
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    float* d_in = reinterpret_cast<float*>(w);
    float* d_out = reinterpret_cast<float*>(w + up256(n * 16));
    uint8_t* d_fl = w + 2 * up256(n * 16);
    CLEDB_CUDA(ctx, cudaMemcpyAsync(d_in, iquv, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_blos(ctx, npix, d_in, kfac, g_eff, ffac, d_out, d_fl, ctx->stream))) return rc;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(out, d_out, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(flags, d_fl, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CLEDB_OK;
}

extern "C" int cledb_qurotate_dev(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref,
                                  const float* hpxy, const float* in, float* out, float* aobs, float* yobs, void* stream) {
    if (!ctx || nx < 0 || ny < 0 || !in || !out || !aobs || !yobs || (!hpxy && (!xs || !ys))) {
        if (ctx) ctx->err = "cledb_qurotate: bad argument";
        return CLEDB_ERR_ARG;
    }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_qurotate(ctx, nx, ny, xs, ys, linpolref, hpxy, in, out, aobs, yobs, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int cledb_qurotate(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref,
                              const float* hpxy, const float* in, float* out, float* aobs, float* yobs) {
    if (!ctx || nx < 0 || ny < 0 || !in || !out || !aobs || !yobs || (!hpxy && (!xs || !ys))) {
        if (ctx) ctx->err = "cledb_qurotate: bad argument";
        return CLEDB_ERR_ARG;
    }
    const size_t n = (size_t)nx * ny;
    if (n == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_in = take(n * 32), o_out = take(n * 32), o_a = take(n * 4), o_y = take(n * 4);
    const size_t o_xs = take((size_t)nx * 8), o_ys = take((size_t)ny * 8), o_hp = take(hpxy ? n * 8 : 0);
    int rc = ensure_workspace(ctx, ctx->io, off);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    cudaStream_t st = ctx->stream;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_in, in, n * 32, cudaMemcpyHostToDevice, st));
    if (hpxy) CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_hp, hpxy, n * 8, cudaMemcpyHostToDevice, st));
    else {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_xs, xs, (size_t)nx * 8, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_ys, ys, (size_t)ny * 8, cudaMemcpyHostToDevice, st));
    }
    rc = launch_qurotate(ctx, nx, ny, reinterpret_cast<double*>(w + o_xs), reinterpret_cast<double*>(w + o_ys), linpolref,
                         hpxy ? reinterpret_cast<float*>(w + o_hp) : nullptr, reinterpret_cast<float*>(w + o_in),
                         reinterpret_cast<float*>(w + o_out), reinterpret_cast<float*>(w + o_a), reinterpret_cast<float*>(w + o_y), st);
    if (rc) return rc;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(out, w + o_out, n * 32, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(aobs, w + o_a, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(yobs, w + o_y, n * 4, cudaMemcpyDeviceToHost, st));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return CLEDB_OK;
}
