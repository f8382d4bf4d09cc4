// Multi-GPU entry points: one process and one context per GPU, pixels sharded, databases replicated.
//
// The search has no exchange step - every pixel is independent (the reference dispatches one pool task per pixel,
// CLEDB_PROC.py:304-313) - so the only collective of the path is the gather of the finished result maps.  Each rank
// builds the packed records of its shard on the device (cledb_invert_dev writes invout | sfound | status | issue codes
// straight into the send buffer), ONE ncclAllGather moves them between the GPUs over NVLink / NVSwitch, and k_unshard
// puts every record at its pixel's place in the full maps.  Nothing travels through host memory in between.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 that is already loaded into the process - the one torch brought
// along - else the system's), so single-GPU users need no NCCL at all and the library never carries a second copy of it
// into a process that already has one.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cledb_internal.h"

using namespace cledb;

// ------------------------------------------------------------------------------------------------------
// the handful of NCCL entry points used, declared here (stable C ABI since NCCL 2.0)
// ------------------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[CLEDB_COMM_ID_BYTES]; };
typedef void* NcclComm;
typedef int (*fn_get_version)(int*);
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*fn_error_string)(int);
constexpr int kNcclUint8 = 1;           // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct Nccl {
    void* handle = nullptr;
    fn_get_version get_version = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
    std::string err;
};
Nccl g_nccl;

bool load_nccl() {
    if (g_nccl.handle) return true;
    void* h = nullptr;
    if (const char* p = std::getenv("CLEDB_B200_NCCL")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy this process has already loaded, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.err = std::string("NCCL not found (") + dlerror() + "); set CLEDB_B200_NCCL to the path of libnccl.so.2"; return false; }
    Nccl n;
    n.handle = h;
    n.get_version = (fn_get_version)dlsym(h, "ncclGetVersion");
    n.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    n.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    n.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    n.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    n.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
    if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_gather) {
        g_nccl.err = "the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
        return false;
    }
    g_nccl = n;
    return true;
}
std::string nccl_text(int rc) { return g_nccl.error_string ? g_nccl.error_string(rc) : ("NCCL error " + std::to_string(rc)); }
}  // namespace

struct cledb_comm {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    Workspace send, recv, full, order;          // device: packed records of this rank, of every rank, full maps, pixel order
};

extern "C" int cledb_comm_unique_id(void* id128) {
    if (!id128) return CLEDB_ERR_ARG;
    if (!load_nccl()) return CLEDB_ERR_COMM;
    NcclId id;
    std::memset(&id, 0, sizeof(id));
    const int rc = g_nccl.get_unique_id(&id);
    if (rc != 0) { g_nccl.err = "ncclGetUniqueId: " + nccl_text(rc); return CLEDB_ERR_COMM; }
    std::memcpy(id128, &id, sizeof(id));
    return CLEDB_OK;
}

extern "C" int cledb_comm_init(cledb_ctx* ctx, int rank, int nranks, const void* id128) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) { ctx->err = "cledb_comm_init: bad rank / nranks / id"; return CLEDB_ERR_ARG; }
    if (ctx->comm) { ctx->err = "cledb_comm_init: the context already has a communicator (cledb_comm_destroy it first)"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cledb_comm* c = new cledb_comm();
    c->rank = rank;
    c->nranks = nranks;
    if (nranks > 1) {
        if (!load_nccl()) { ctx->err = "cledb_comm_init: " + g_nccl.err; delete c; return CLEDB_ERR_COMM; }
        NcclId id;
        std::memcpy(&id, id128, sizeof(id));
        const int rc = g_nccl.comm_init_rank(&c->comm, nranks, id, rank);
        if (rc != 0) { ctx->err = "cledb_comm_init: ncclCommInitRank: " + nccl_text(rc); delete c; return CLEDB_ERR_COMM; }
    }
    ctx->comm = c;
    return CLEDB_OK;
}

extern "C" int cledb_comm_destroy(cledb_ctx* ctx) {
    if (!ctx) return CLEDB_ERR_ARG;
    cledb_comm* c = ctx->comm;
    if (!c) return CLEDB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (c->comm) g_nccl.comm_destroy(c->comm);
    for (Workspace* w : {&c->send, &c->recv, &c->full, &c->order})
        if (w->ptr) cudaFree(w->ptr);
    delete c;
    ctx->comm = nullptr;
    return CLEDB_OK;
}

extern "C" int cledb_comm_info(cledb_ctx* ctx, int32_t info[4]) {
    if (!ctx || !info) return CLEDB_ERR_ARG;
    info[0] = ctx->comm ? ctx->comm->rank : 0;
    info[1] = ctx->comm ? ctx->comm->nranks : 1;
    int v = 0;
    if (g_nccl.handle && g_nccl.get_version) g_nccl.get_version(&v);
    info[2] = v;
    info[3] = ctx->comm && ctx->comm->comm ? 1 : 0;
    return CLEDB_OK;
}

extern "C" int cledb_allgather_dev(cledb_ctx* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!ctx->comm) { ctx->err = "cledb_allgather: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    if (!send || !recv || bytes < 0) { ctx->err = "cledb_allgather: bad argument"; return CLEDB_ERR_ARG; }
    if (bytes == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    cledb_comm* c = ctx->comm;
    if (c->nranks == 1) {
        if (send != recv) CLEDB_CUDA(ctx, cudaMemcpyAsync(recv, send, (size_t)bytes, cudaMemcpyDeviceToDevice, st));
        return CLEDB_OK;
    }
    const int rc = g_nccl.all_gather(send, recv, (size_t)bytes, kNcclUint8, c->comm, st);
    if (rc != 0) { ctx->err = "cledb_allgather: ncclAllGather: " + nccl_text(rc); return CLEDB_ERR_COMM; }
    return CLEDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// pixel partition
// ------------------------------------------------------------------------------------------------------
// Pixels are ordered by `key` (stable; the caller passes db_enc, the same on every rank), so that a rank touches few
// databases, and cut into nranks contiguous ranges of equal total weight; weight of a pixel = weight_of_key[key] (the rows
// of its database: what one search of that pixel costs; NULL = every pixel weighs the same).
extern "C" int cledb_partition(int64_t npix, const int32_t* key, int nkeys, const double* weight_of_key, int nranks,
                               int32_t* order_out, int64_t* bounds_out) {
    if (npix < 0 || nranks < 1 || nkeys < 1 || (npix > 0 && (!key || !order_out)) || !bounds_out) return CLEDB_ERR_ARG;
    for (int64_t p = 0; p < npix; ++p)
        if (key[p] < 0 || key[p] >= nkeys) return CLEDB_ERR_ARG;
    std::vector<int64_t> start((size_t)nkeys + 1, 0);
    for (int64_t p = 0; p < npix; ++p) start[(size_t)key[p] + 1]++;
    double total = 0.0;
    for (int id = 0; id < nkeys; ++id) total += (weight_of_key ? weight_of_key[id] : 1.0) * (double)start[(size_t)id + 1];
    for (size_t i = 1; i < start.size(); ++i) start[i] += start[i - 1];
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t p = 0; p < npix; ++p) order_out[cur[key[p]]++] = (int32_t)p;
    }
    // cut r = first position whose cumulative weight reaches total * r / nranks (np.searchsorted(csum, x, 'left'))
    bounds_out[0] = 0;
    bounds_out[nranks] = npix;
    int r = 1;
    double csum = 0.0;
    for (int64_t i = 0; i < npix && r < nranks; ++i) {
        csum += weight_of_key ? weight_of_key[key[order_out[i]]] : 1.0;
        while (r < nranks && csum >= total * (double)r / (double)nranks) bounds_out[r++] = i;
    }
    while (r < nranks) bounds_out[r++] = npix;
    return CLEDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// packed records of a shard and their scatter into the full maps
// ------------------------------------------------------------------------------------------------------
namespace {
struct Packed { size_t o_inv, o_sf, o_st, o_iss, bytes; };
Packed packed_layout(int64_t maxcount, int k) {
    Packed L;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    L.o_inv = off; off = up(off + (size_t)maxcount * k * 44);
    L.o_sf = off;  off = up(off + (size_t)maxcount * k * 32);
    L.o_st = off;  off = up(off + (size_t)maxcount);
    L.o_iss = off; off = up(off + (size_t)maxcount * 4);
    L.bytes = off;
    return L;
}
}  // namespace

namespace cledb {
// One warp per pixel: the pixel's invout row (k*44 B) and sfound row (k*32 B) move as 16-byte vectors when k is a multiple of
// 4 (352 + 256 B at nsearch 8: 22 + 16 vectors, one or two per lane), as 4-byte words otherwise; lane 0 adds the status
// byte and the two issue bytes.  Reads are contiguous (the packed records of a rank), writes are whole rows.
template <bool VEC>
__global__ void __launch_bounds__(256) k_unshard(int nranks, const int64_t* __restrict__ bounds, const int32_t* __restrict__ order,
                                                 int64_t npix, int k, const unsigned char* __restrict__ recv, size_t rank_bytes,
                                                 size_t o_inv, size_t o_sf, size_t o_st, size_t o_iss, float* __restrict__ invout,
                                                 float* __restrict__ sfound, uint8_t* __restrict__ status, int32_t* __restrict__ issue) {
    const int lane = threadIdx.x & 31;
    const int64_t spos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // position in the database-sorted order
    if (spos >= npix) return;
    int r = 0;
    while (r + 1 < nranks && spos >= bounds[r + 1]) ++r;
    const int64_t j = spos - bounds[r];
    const int64_t p = order[spos];
    const unsigned char* blk = recv + (size_t)r * rank_bytes;
    const int wi = k * 11, ws = k * 8;
    if (VEC) {
        const float4* si = reinterpret_cast<const float4*>(blk + o_inv) + j * (wi / 4);
        const float4* ss = reinterpret_cast<const float4*>(blk + o_sf) + j * (ws / 4);
        float4* di = reinterpret_cast<float4*>(invout) + p * (wi / 4);
        float4* ds = reinterpret_cast<float4*>(sfound) + p * (ws / 4);
        for (int w = lane; w < wi / 4; w += 32) di[w] = __ldcs(si + w);
        for (int w = lane; w < ws / 4; w += 32) ds[w] = __ldcs(ss + w);
    } else {
        const float* si = reinterpret_cast<const float*>(blk + o_inv) + j * wi;
        const float* ss = reinterpret_cast<const float*>(blk + o_sf) + j * ws;
        for (int w = lane; w < wi; w += 32) invout[p * wi + w] = si[w];
        for (int w = lane; w < ws; w += 32) sfound[p * ws + w] = ss[w];
    }
    if (lane == 0) {
        if (status) status[p] = blk[o_st + j];
        if (issue) issue[p] = reinterpret_cast<const int32_t*>(blk + o_iss)[j];
    }
}
}  // namespace cledb

extern "C" int cledb_shard_layout(int64_t maxcount, int nsearch, int64_t layout[5]) {
    if (!layout || maxcount < 0 || nsearch < 1) return CLEDB_ERR_ARG;
    const Packed L = packed_layout(maxcount, nsearch);
    layout[0] = (int64_t)L.o_inv; layout[1] = (int64_t)L.o_sf; layout[2] = (int64_t)L.o_st; layout[3] = (int64_t)L.o_iss; layout[4] = (int64_t)L.bytes;
    return CLEDB_OK;
}

extern "C" int cledb_unshard_dev(cledb_ctx* ctx, int nranks, const int64_t* bounds_dev, const int32_t* order_dev, int64_t npix,
                                 int64_t maxcount, int nsearch, const void* recv, float* invout, float* sfound, uint8_t* status,
                                 int32_t* issue, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (nranks < 1 || !bounds_dev || !order_dev || !recv || !invout || !sfound || nsearch < 1 || npix < 0) { ctx->err = "cledb_unshard: bad argument"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const Packed L = packed_layout(maxcount, nsearch);
    const unsigned grid = (unsigned)((npix * 32 + 255) / 256);
    if (nsearch % 4 == 0)
        k_unshard<true><<<grid, 256, 0, st>>>(nranks, bounds_dev, order_dev, npix, nsearch, reinterpret_cast<const unsigned char*>(recv),
                                              L.bytes, L.o_inv, L.o_sf, L.o_st, L.o_iss, invout, sfound, status, issue);
    else
        k_unshard<false><<<grid, 256, 0, st>>>(nranks, bounds_dev, order_dev, npix, nsearch, reinterpret_cast<const unsigned char*>(recv),
                                               L.bytes, L.o_inv, L.o_sf, L.o_st, L.o_iss, invout, sfound, status, issue);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// one map over all ranks: every rank passes ITS pixels (order[bounds[rank] : bounds[rank+1]] of the map)
// ------------------------------------------------------------------------------------------------------
static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int cledb_invert_sharded(cledb_ctx* ctx, int mode, int precision, int64_t nlocal, const float* sobs, const float* snr,
                                    const int32_t* db_id, int nsearch, const float* yobs, const float* dobs, const float* rot_cos,
                                    const float* rot_sin, const void* dopp, int dopp_is_f64, int64_t dopp_stride,
                                    const cledb_dbgrid* grid, double maxchisq, int bcalc, int strict_topk, int64_t npix,
                                    const int32_t* order, const int64_t* bounds, int root, float* invout, float* sfound,
                                    uint8_t* status_out, int32_t* issue_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!ctx->comm) { ctx->err = "cledb_invert_sharded: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    cledb_comm* c = ctx->comm;
    const int R = c->nranks, me = c->rank;
    const bool want = root < 0 || root == me;
    if (npix < 0 || nlocal < 0 || !grid || !order || !bounds || (want && (!invout || !sfound)) ||
        (nlocal > 0 && (!sobs || !snr || !db_id || !yobs || !dobs || !rot_cos || !rot_sin))) {
        ctx->err = "cledb_invert_sharded: null argument";
        return CLEDB_ERR_ARG;
    }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "cledb_invert_sharded: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (mode == CLEDB_MODE_IQUD && nlocal > 0 && !dopp) { ctx->err = "cledb_invert_sharded: IQUD needs the Doppler field strengths"; return CLEDB_ERR_ARG; }
    if (bounds[0] != 0 || bounds[R] != npix || bounds[me + 1] - bounds[me] != nlocal) {
        ctx->err = "cledb_invert_sharded: the partition does not match this rank's pixel count";
        return CLEDB_ERR_ARG;
    }
    if (npix == 0) return CLEDB_OK;
    if (npix > (int64_t)1 << 30) { ctx->err = "cledb_invert_sharded: more than 2^30 pixels in one call"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t k = (size_t)nsearch, n = (size_t)nlocal;
    const size_t dsz = dopp ? (dopp_is_f64 ? 8 : 4) : 0;
    if (dopp_stride <= 0) dopp_stride = 1;
    int64_t maxcount = 0;
    for (int r = 0; r < R; ++r) maxcount = std::max(maxcount, bounds[r + 1] - bounds[r]);
    const Packed L = packed_layout(maxcount, nsearch);

    // ---- device buffers: this rank's inputs (ctx->io), packed send / recv, pixel order + bounds, full maps
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t nn = std::max<size_t>(n, 1);
    const size_t o_sobs = take(nn * 32), o_snr = take(nn * 32), o_id = take(nn * 4), o_y = take(nn * 4), o_d = take(nn * 4),
                 o_rc = take(nn * 4), o_rs = take(nn * 4), o_dopp = take(nn * dsz),
                 o_tab = take((size_t)(2 * grid->nbphi + 2 * grid->nbtheta) * 4);
    int rc;
    if ((rc = ensure_workspace(ctx, ctx->io, off))) return rc;
    if ((rc = ensure_workspace(ctx, c->send, L.bytes))) return rc;
    if ((rc = ensure_workspace(ctx, c->recv, L.bytes * (size_t)R))) return rc;
    const size_t o_bounds = up256((size_t)npix * 4);
    if ((rc = ensure_workspace(ctx, c->order, o_bounds + ((size_t)R + 1) * 8))) return rc;
    size_t foff = 0;
    auto ftake = [&](size_t b) { size_t o = foff; foff = up256(foff + b); return o; };
    const size_t f_inv = ftake(want ? (size_t)npix * k * 44 : 0), f_sf = ftake(want ? (size_t)npix * k * 32 : 0),
                 f_st = ftake(want ? (size_t)npix : 0), f_iss = ftake(want ? (size_t)npix * 4 : 0);
    if (want && (rc = ensure_workspace(ctx, c->full, foff))) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    if (n > 0) {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_sobs, sobs, n * 32, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_snr, snr, n * 32, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_id, db_id, n * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_y, yobs, n * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_d, dobs, n * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_rc, rot_cos, n * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_rs, rot_sin, n * 4, cudaMemcpyHostToDevice, st));
        if (dsz) {
            if (dopp_stride == 1) CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_dopp, dopp, n * dsz, cudaMemcpyHostToDevice, st));
            else CLEDB_CUDA(ctx, cudaMemcpy2DAsync(w + o_dopp, dsz, dopp, (size_t)dopp_stride * dsz, dsz, n, cudaMemcpyHostToDevice, st));
        }
    }
    cledb_dbgrid g = *grid;
    float* tab = reinterpret_cast<float*>(w + o_tab);
    CLEDB_CUDA(ctx, cudaMemcpyAsync(tab, grid->sin_bphi, (size_t)g.nbphi * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + g.nbphi, grid->cos_bphi, (size_t)g.nbphi * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + 2 * g.nbphi, grid->sin_btheta, (size_t)g.nbtheta * 4, cudaMemcpyHostToDevice, st));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(tab + 2 * g.nbphi + g.nbtheta, grid->cos_btheta, (size_t)g.nbtheta * 4, cudaMemcpyHostToDevice, st));
    g.sin_bphi = tab;
    g.cos_bphi = tab + g.nbphi;
    g.sin_btheta = tab + 2 * g.nbphi;
    g.cos_btheta = tab + 2 * g.nbphi + g.nbtheta;
    unsigned char* dord = reinterpret_cast<unsigned char*>(c->order.ptr);
    if (want) {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(dord, order, (size_t)npix * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(dord + o_bounds, bounds, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st));
    }

    // ---- search + solutions of the shard, written straight into the packed send buffer
    unsigned char* send = reinterpret_cast<unsigned char*>(c->send.ptr);
    if (n > 0) {
        rc = cledb_invert_dev(ctx, mode, precision, nlocal, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_snr),
                              reinterpret_cast<int32_t*>(w + o_id), nsearch, reinterpret_cast<float*>(w + o_y), reinterpret_cast<float*>(w + o_d),
                              reinterpret_cast<float*>(w + o_rc), reinterpret_cast<float*>(w + o_rs), dsz ? w + o_dopp : nullptr, dopp_is_f64, 1,
                              &g, maxchisq, bcalc, strict_topk, reinterpret_cast<float*>(send + L.o_inv), reinterpret_cast<float*>(send + L.o_sf),
                              send + L.o_st, reinterpret_cast<int32_t*>(send + L.o_iss), st);
        if (rc) return rc;        // a rank that fails here leaves the others waiting in the collective: the launcher aborts the job
    }
    // ---- the one collective of the path, then every record to its pixel
    if ((rc = cledb_allgather_dev(ctx, send, c->recv.ptr, (int64_t)L.bytes, st))) return rc;
    if (want) {
        unsigned char* f = reinterpret_cast<unsigned char*>(c->full.ptr);
        rc = cledb_unshard_dev(ctx, R, reinterpret_cast<int64_t*>(dord + o_bounds), reinterpret_cast<int32_t*>(dord), npix, maxcount, nsearch,
                               c->recv.ptr, reinterpret_cast<float*>(f + f_inv), reinterpret_cast<float*>(f + f_sf), f + f_st, reinterpret_cast<int32_t*>(f + f_iss), st);
        if (rc) return rc;
        CLEDB_CUDA(ctx, cudaMemcpyAsync(invout, f + f_inv, (size_t)npix * k * 44, cudaMemcpyDeviceToHost, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(sfound, f + f_sf, (size_t)npix * k * 32, cudaMemcpyDeviceToHost, st));
        if (status_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(status_out, f + f_st, (size_t)npix, cudaMemcpyDeviceToHost, st));
        if (issue_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(issue_out, f + f_iss, (size_t)npix * 4, cudaMemcpyDeviceToHost, st));
    }
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return check_async_error(ctx);
}
