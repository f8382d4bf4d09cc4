// Multi-GPU entry points: one process and one context per GPU, pixels sharded, databases replicated.
//
// The search has no exchange step - every pixel is independent (the reference dispatches one pool task per pixel,
// CLEDB_PROC.py:304-313) - so the only collective of the path is the gather of the finished result maps.  Each rank
// builds the packed records of its shard on the device (cledb_invert_dev writes invout | sfound | status | issue codes
// straight into the send buffer); then, by transport:
//   peer memory (default)  the receive windows of all ranks are mapped into every process (CUDA IPC) and k_push_records
//                          stores every record once, at its pixel's place, in the maps of every rank over NVLink / NVSwitch
//   NCCL                   ONE ncclAllGather moves the packed records, k_unshard puts every record at its pixel's place
// Nothing travels through host memory in between.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 that is already loaded into the process - the one torch brought
// along - else the system's), so single-GPU users need no NCCL at all and the library never carries a second copy of it
// into a process that already has one.  With the peer transport it only bootstraps (the exchange of the IPC handles).
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

// ------------------------------------------------------------------------------------------------------
// peer memory: every rank's window mapped into every process (CUDA IPC), collectives as plain stores over NVLink
// ------------------------------------------------------------------------------------------------------
// NCCL's all-gather costs the path three passes over the finished maps: k_build writes the packed records, NCCL copies
// them into every rank's receive buffer, k_unshard reads that buffer and writes the maps.  With the maps of every rank
// mapped into every process the push kernels below store each record ONCE, straight at its pixel's place in every rank's
// maps (16-byte vectors, whole rows, NVLink / NVSwitch as the store target), and the only things left of the collective
// are two rounds of flags inside that one kernel:
//   free[e]  "my window may be overwritten by operation e"      sent by block 0 when the push kernel starts; every block waits for all
//   done[e]  "all my stores of operation e have been performed" sent by the block that finishes last, which then waits for all
// Flags are 32-bit epochs in a small IPC allocation of their own; every rank writes its own slot in every peer's array
// (st.release.sys after __threadfence_system) and spins on its local array (ld.acquire.sys).  Collective calls come in the
// same order on every rank, so the host-side epoch counters agree without any exchange.
constexpr int kMaxRanks = 16;
constexpr int kMaxPieces = 8;         // pieces of a shard in the pipelined gather
constexpr int kFlagFree = 0, kFlagDone = 32, kFlagCount = 64, kFlagWords = 128;      // word offsets inside the flag array
constexpr long long kSpinLimitNs = 60LL * 1000 * 1000 * 1000;                         // a peer that never arrives: give up, report

struct Peers {
    unsigned char* base[kMaxRanks];     // window of rank r as seen from this process (base[me] = the local allocation)
    unsigned*      flags[kMaxRanks];    // flag array of rank r
    int n, me;
};

struct cledb_comm {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    Workspace send, order, scratch;             // device: packed records of this rank, pixel order + bounds, small exchange buffer
    Workspace recv;                             // NCCL transport only: packed records of every rank
    // peer transport
    bool peer_ok = false;                       // flag arrays of all ranks are mapped: windows can be exchanged
    int transport = 0;                          // 0 = NCCL collectives, 1 = stores into the peers' windows
    std::string peer_why;                       // why peer_ok is false
    Peers peers;                                // window bases (null until cledb_comm_window) and flag arrays
    unsigned* flags_local = nullptr;
    void* flags_opened[kMaxRanks] = {};         // what cudaIpcOpenMemHandle returned (block bases), for cudaIpcCloseMemHandle
    void* win_opened[kMaxRanks] = {};
    cudaStream_t side = nullptr;                // second stream + events of the pipelined gather (cledb_invert_gather_dev)
    cudaEvent_t ev[kMaxPieces + 1] = {};
    void* win = nullptr;                        // local window (plain allocation when peer_ok is false)
    std::vector<size_t> win_bytes;              // size of every rank's window, tracked identically on all ranks
    unsigned epoch = 0;
    int push_blocks = 0;                        // > 0: bound of the grid of the push kernels (cledb_comm_push_blocks)
};

namespace {
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long now_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// thread r < n waits until rank r's slot of the local flag array has reached `epoch` (wrap-safe); a peer that does not
// arrive within kSpinLimitNs leaves code 2 in the context's mapped error word instead of hanging the GPU
__device__ __forceinline__ void wait_flags(const Peers& P, int which, unsigned epoch, int32_t* err_mapped) {
    if ((int)threadIdx.x < P.n) {
        const unsigned* f = P.flags[P.me] + which + threadIdx.x;
        const long long t0 = now_ns();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(64);
            if (now_ns() - t0 > kSpinLimitNs) { if (err_mapped) *err_mapped = 2; break; }
        }
    }
}
__device__ __forceinline__ void signal_flags(const Peers& P, int which, unsigned epoch) {
    if ((int)threadIdx.x < P.n) st_release_sys(P.flags[threadIdx.x] + which + P.me, epoch);
}
// A push kernel opens with this: block 0 tells every peer that this rank's window may be overwritten by operation `epoch`
// (everything that read the window was enqueued before this kernel), every block waits until all peers have said the same.
__device__ __forceinline__ void block_open(const Peers& P, unsigned epoch, int32_t* err_mapped) {
    if (blockIdx.x == 0) {
        __threadfence_system();
        signal_flags(P, kFlagFree, epoch);
    }
    wait_flags(P, kFlagFree, epoch, err_mapped);
    __syncthreads();
}
// ... and closes with this: every thread has fenced its stores; the block that arrives last at the counter tells every peer
// that this rank is done and then waits for the done flags of all peers, so the kernel's completion on the stream means
// that every record of every rank is in place in this rank's window.
__device__ __forceinline__ void block_done(const Peers& P, unsigned epoch, int32_t* err_mapped) {
    __shared__ int last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(P.flags[P.me] + kFlagCount, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) P.flags[P.me][kFlagCount] = 0;
        __threadfence_system();
        signal_flags(P, kFlagDone, epoch);
        wait_flags(P, kFlagDone, epoch, err_mapped);
    }
}


// all-gather of contiguous blocks: the 16-byte vectors of this rank's block go to the same place of every rank's window
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_push_block(Peers P, const uint4* __restrict__ send, size_t nvec, size_t dst_off, unsigned epoch,
                                                        int32_t* err_mapped) {
    block_open(P, epoch, err_mapped);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 v = __ldcs(send + i);
        for (int q = 0; q < P.n; ++q) {
            int r = P.me + 1 + q + (int)(blockIdx.x % (unsigned)P.n);      // blocks start at different peers
            r %= P.n;
            reinterpret_cast<uint4*>(P.base[r] + dst_off)[i] = v;
        }
    }
    block_done(P, epoch, err_mapped);
}
}  // namespace

static int ensure_err_word(cledb_ctx* ctx) {
    if (!ctx->h_err) {
        CLEDB_CUDA(ctx, cudaHostAlloc(&ctx->h_err, 64, cudaHostAllocMapped));
        std::memset(ctx->h_err, 0, 64);
        CLEDB_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_err_mapped, ctx->h_err, 0));
    }
    return CLEDB_OK;
}
static inline int32_t* comm_err_word(cledb_ctx* ctx) { return ctx->d_err_mapped ? ctx->d_err_mapped + 1 : nullptr; }

// `bytes` per rank from `mine` (host) to `all` (host, nranks * bytes) through the communicator; doubles as a barrier
static int exchange_host(cledb_ctx* ctx, const void* mine, void* all, size_t bytes) {
    cledb_comm* c = ctx->comm;
    if (c->nranks == 1) { std::memcpy(all, mine, bytes); return CLEDB_OK; }
    int rc;
    if ((rc = ensure_workspace(ctx, c->scratch, bytes * ((size_t)c->nranks + 1)))) return rc;
    unsigned char* d = reinterpret_cast<unsigned char*>(c->scratch.ptr);
    CLEDB_CUDA(ctx, cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int nrc = g_nccl.all_gather(d, d + bytes, bytes, kNcclUint8, c->comm, ctx->stream);
    if (nrc != 0) { ctx->err = "ncclAllGather: " + nccl_text(nrc); return CLEDB_ERR_COMM; }
    CLEDB_CUDA(ctx, cudaMemcpyAsync(all, d + bytes, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CLEDB_OK;
}

// cudaMalloc may carve a small allocation out of a larger block, and an IPC handle always names the whole block: export
// the block's base and ship the offset of `mine` inside it (cuMemGetAddressRange through the runtime's driver entry points).
static void base_of(void* p, void** base, size_t* offset) {
    typedef int (*fn_range)(unsigned long long*, size_t*, unsigned long long);
    static fn_range range = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            range = (fn_range)f;
        else
            cudaGetLastError();
    }
    *base = p;
    *offset = 0;
    unsigned long long b = 0;
    size_t sz = 0;
    if (range && range(&b, &sz, (unsigned long long)(uintptr_t)p) == 0 && b) {
        *base = (void*)(uintptr_t)b;
        *offset = (size_t)((uintptr_t)p - (uintptr_t)b);
    }
}

// Maps the `mine` allocation of every rank into this process: out[r] (out[me] = mine), opened[r] = what cudaIpcCloseMemHandle
// wants back.  Collective.  Every rank stamps the first word of its allocation and reads the stamps of the others through the
// new mappings (word at stamp_off); ok = 0 on return (with CLEDB_OK) when some rank could not export, open or verify: every rank then closes
// what it opened.
static int map_peers(cledb_ctx* ctx, void* mine, size_t stamp_off, void** out, void** opened, int* ok, std::string* why) {
    cledb_comm* c = ctx->comm;
    const int R = c->nranks, me = c->rank;
    for (int r = 0; r < R; ++r) out[r] = opened[r] = nullptr;
    out[me] = mine;
    *ok = 1;
    if (R == 1) return CLEDB_OK;
    struct Msg { cudaIpcMemHandle_t h; unsigned long long offset; int good; int pad; };
    Msg m;
    std::memset(&m, 0, sizeof(m));
    void* base = nullptr;
    size_t offset = 0;
    base_of(mine, &base, &offset);
    m.offset = offset;
    const unsigned stamp = 0xC1EDB000u + (unsigned)me;
    cudaError_t e = cudaMemcpy(reinterpret_cast<unsigned char*>(mine) + stamp_off, &stamp, 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&m.h, base);
    m.good = e == cudaSuccess;
    if (!m.good) { *why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); }
    if (std::getenv("CLEDB_B200_NO_IPC")) { m.good = 0; *why = "CUDA IPC switched off by CLEDB_B200_NO_IPC"; }     // exercises the fallback
    std::vector<Msg> all((size_t)R);
    int rc;
    if ((rc = exchange_host(ctx, &m, all.data(), sizeof(Msg)))) return rc;
    int good = 1;
    for (int r = 0; r < R; ++r) good &= all[(size_t)r].good;
    if (!good && why->empty()) *why = "another rank could not export its memory (CUDA IPC)";
    if (good)
        for (int r = 0; r < R; ++r) {
            if (r == me) continue;
            e = cudaIpcOpenMemHandle(&opened[r], all[(size_t)r].h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                *why = "cudaIpcOpenMemHandle (rank " + std::to_string(r) + "): " + cudaGetErrorString(e);
                cudaGetLastError();
                opened[r] = nullptr;
                good = 0;
                break;
            }
            out[r] = reinterpret_cast<unsigned char*>(opened[r]) + all[(size_t)r].offset;
            unsigned seen = 0;
            e = cudaMemcpy(&seen, reinterpret_cast<unsigned char*>(out[r]) + stamp_off, 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess || seen != 0xC1EDB000u + (unsigned)r) {
                *why = "the mapping of rank " + std::to_string(r) + "'s memory does not show that rank's stamp" +
                       (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string());
                cudaGetLastError();
                good = 0;
                break;
            }
        }
    // second round: did every rank open and verify every handle?
    std::vector<int> oks((size_t)R);
    if ((rc = exchange_host(ctx, &good, oks.data(), sizeof(int)))) return rc;
    for (int r = 0; r < R; ++r) good &= oks[(size_t)r];
    if (!good) {
        if (why->empty()) *why = "another rank could not map this rank's memory (CUDA IPC)";
        for (int r = 0; r < R; ++r) {
            if (r != me && opened[r]) cudaIpcCloseMemHandle(opened[r]);
            if (r != me) out[r] = opened[r] = nullptr;
        }
        *ok = 0;
    }
    return CLEDB_OK;
}

static void unmap_peers(cledb_comm* c, void** opened) {
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && opened[r]) { cudaIpcCloseMemHandle(opened[r]); opened[r] = nullptr; }
}

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
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) {
        ctx->err = "cledb_comm_init: bad rank / nranks (1..16) / id";
        return CLEDB_ERR_ARG;
    }
    if (ctx->comm) { ctx->err = "cledb_comm_init: the context already has a communicator (cledb_comm_destroy it first)"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cledb_comm* c = new cledb_comm();
    c->rank = rank;
    c->nranks = nranks;
    std::memset(&c->peers, 0, sizeof(c->peers));
    c->peers.n = nranks;
    c->peers.me = rank;
    c->win_bytes.assign((size_t)nranks, 0);
    if (nranks > 1) {
        if (!load_nccl()) { ctx->err = "cledb_comm_init: " + g_nccl.err; delete c; return CLEDB_ERR_COMM; }
        NcclId id;
        std::memcpy(&id, id128, sizeof(id));
        const int rc = g_nccl.comm_init_rank(&c->comm, nranks, id, rank);
        if (rc != 0) { ctx->err = "cledb_comm_init: ncclCommInitRank: " + nccl_text(rc); delete c; return CLEDB_ERR_COMM; }
    }
    ctx->comm = c;
    // flag arrays: zeroed, then mapped everywhere (the exchange inside map_peers is the barrier between the two)
    int rc = ensure_err_word(ctx);
    if (rc == CLEDB_OK && cudaMalloc(&c->flags_local, kFlagWords * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); rc = CLEDB_ERR_NOMEM; ctx->err = "cledb_comm_init: out of device memory"; }
    if (rc == CLEDB_OK) {
        cudaMemset(c->flags_local, 0, kFlagWords * sizeof(unsigned));
        cudaDeviceSynchronize();
        const char* env = std::getenv("CLEDB_B200_PEER");
        void* fl[kMaxRanks];
        int ok = 0;
        rc = map_peers(ctx, c->flags_local, (kFlagWords - 1) * sizeof(unsigned), fl, c->flags_opened, &ok, &c->peer_why);
        if (rc == CLEDB_OK) {
            for (int r = 0; r < nranks; ++r) c->peers.flags[r] = reinterpret_cast<unsigned*>(fl[r]);
            c->peer_ok = ok != 0;
            c->transport = c->peer_ok && !(env && env[0] == '0') ? 1 : 0;
        }
    }
    if (rc != CLEDB_OK) { cledb_comm_destroy(ctx); return rc; }
    return CLEDB_OK;
}

// everything enqueued has run on every rank
static int comm_quiesce(cledb_ctx* ctx) {
    cudaDeviceSynchronize();
    int one = 1;
    std::vector<int> all((size_t)ctx->comm->nranks);
    return exchange_host(ctx, &one, all.data(), sizeof(int));
}

static void release_window(cledb_ctx* ctx) {
    cledb_comm* c = ctx->comm;
    unmap_peers(c, c->win_opened);
    for (int r = 0; r < c->nranks; ++r) c->peers.base[r] = nullptr;
    if (c->nranks > 1 && c->comm) comm_quiesce(ctx);      // nobody still holds a mapping of what is freed next
    if (c->win) cudaFree(c->win);
    c->win = nullptr;
    std::fill(c->win_bytes.begin(), c->win_bytes.end(), (size_t)0);
}

extern "C" int cledb_comm_destroy(cledb_ctx* ctx) {
    if (!ctx) return CLEDB_ERR_ARG;
    cledb_comm* c = ctx->comm;
    if (!c) return CLEDB_OK;
    cudaSetDevice(ctx->device);
    if (c->nranks > 1 && c->comm) comm_quiesce(ctx);
    else cudaDeviceSynchronize();
    release_window(ctx);
    unmap_peers(c, c->flags_opened);
    if (c->nranks > 1 && c->comm) comm_quiesce(ctx);
    if (c->flags_local) cudaFree(c->flags_local);
    if (c->side) {
        cudaStreamDestroy(c->side);
        for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
    }
    if (c->comm) g_nccl.comm_destroy(c->comm);
    for (Workspace* w : {&c->send, &c->recv, &c->order, &c->scratch})
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

extern "C" int cledb_comm_transport(cledb_ctx* ctx, int want, int32_t* active) {
    if (!ctx) return CLEDB_ERR_ARG;
    cledb_comm* c = ctx->comm;
    if (!c) { ctx->err = "cledb_comm_transport: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    if (want < -1 || want > 1) { ctx->err = "cledb_comm_transport: want must be -1 (query), 0 (NCCL) or 1 (peer memory)"; return CLEDB_ERR_ARG; }
    if (want == 1 && !c->peer_ok) {
        ctx->err = "cledb_comm_transport: peer memory is not available: " + (c->peer_why.empty() ? std::string("unknown reason") : c->peer_why);
        if (active) *active = c->transport;
        return CLEDB_ERR_COMM;
    }
    if (want == 0 && c->nranks > 1 && !c->comm) { ctx->err = "cledb_comm_transport: no NCCL communicator"; return CLEDB_ERR_COMM; }
    if (want >= 0) c->transport = want;
    if (active) *active = c->transport;
    return CLEDB_OK;
}

// The window of rank r must hold need[r] bytes.  Collective (the sizes are tracked identically on every rank, so either all
// ranks re-allocate or none does); the contents do not survive a re-allocation.
static int ensure_window(cledb_ctx* ctx, const size_t* need) {
    cledb_comm* c = ctx->comm;
    const int R = c->nranks, me = c->rank;
    bool grow = c->win == nullptr;
    for (int r = 0; r < R; ++r) grow = grow || c->win_bytes[(size_t)r] < need[r];
    if (!grow) return CLEDB_OK;
    int rc;
    if (R > 1 && (rc = comm_quiesce(ctx))) return rc;
    release_window(ctx);
    const size_t mine = std::max<size_t>(need[me], 256);
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, mine);
    int good = e == cudaSuccess;
    if (!good) cudaGetLastError();
    std::vector<int> oks((size_t)R);
    if ((rc = exchange_host(ctx, &good, oks.data(), sizeof(int)))) { if (p) cudaFree(p); return rc; }
    for (int r = 0; r < R; ++r) good &= oks[(size_t)r];
    if (!good) {
        if (p) cudaFree(p);
        ctx->err = "cledb_comm_window: out of device memory on some rank (" + std::to_string(mine >> 20) + " MiB asked here)";
        return CLEDB_ERR_NOMEM;
    }
    c->win = p;
    c->peers.base[me] = reinterpret_cast<unsigned char*>(p);
    if (c->peer_ok) {
        void* w[kMaxRanks];
        int ok = 0;
        std::string why;
        if ((rc = map_peers(ctx, p, 0, w, c->win_opened, &ok, &why))) return rc;
        if (ok)
            for (int r = 0; r < R; ++r) c->peers.base[r] = reinterpret_cast<unsigned char*>(w[r]);
        else {                                          // the same verdict on every rank: back to NCCL for good
            unmap_peers(c, c->flags_opened);
            for (int r = 0; r < R; ++r) if (r != me) c->peers.flags[r] = nullptr;
            c->peer_ok = false;
            c->transport = 0;
            c->peer_why = why;
        }
    }
    for (int r = 0; r < R; ++r) c->win_bytes[(size_t)r] = std::max<size_t>(need[r], 256);
    return CLEDB_OK;
}

extern "C" int cledb_comm_push_blocks(cledb_ctx* ctx, int blocks) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!ctx->comm) { ctx->err = "cledb_comm_push_blocks: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    if (blocks < 0) { ctx->err = "cledb_comm_push_blocks: blocks must be >= 0"; return CLEDB_ERR_ARG; }
    ctx->comm->push_blocks = blocks;
    return CLEDB_OK;
}

extern "C" int cledb_comm_window(cledb_ctx* ctx, int64_t bytes, void** ptr_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!ctx->comm) { ctx->err = "cledb_comm_window: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    if (bytes < 0 || !ptr_out) { ctx->err = "cledb_comm_window: bad argument"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t need[kMaxRanks];
    for (int r = 0; r < ctx->comm->nranks; ++r) need[r] = (size_t)bytes;
    const int rc = ensure_window(ctx, need);
    *ptr_out = rc ? nullptr : ctx->comm->win;
    return rc;
}

static inline bool in_window(const cledb_comm* c, const void* p, size_t bytes) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(c->win);
    const unsigned char* q = reinterpret_cast<const unsigned char*>(p);
    return b && q >= b && q + bytes <= b + c->win_bytes[(size_t)c->rank];
}

static int after_launch(cledb_ctx* ctx, int n = 1) {
    ctx->launches += n;
    CLEDB_CUDA(ctx, cudaGetLastError());
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
    const size_t total = (size_t)bytes * (size_t)c->nranks;
    if (c->transport == 1 && in_window(c, recv, total) && bytes % 16 == 0 && ((uintptr_t)send % 16) == 0 && ((uintptr_t)recv % 16) == 0) {
        // receive buffer inside the window: every rank stores its block into every rank's window
        const unsigned e = ++c->epoch;
        const size_t nvec = (size_t)bytes / 16;
        const size_t off = (size_t)(reinterpret_cast<unsigned char*>(recv) - reinterpret_cast<unsigned char*>(c->win)) + (size_t)c->rank * (size_t)bytes;
        if (c->push_blocks > 0) {              // a gather that runs beside a search: few SMs, each filled with one 1024-thread block
            const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((nvec + 1023) / 1024, (size_t)c->push_blocks));
            k_push_block<1024><<<grid, 1024, 0, st>>>(c->peers, reinterpret_cast<const uint4*>(send), nvec, off, e, comm_err_word(ctx));
        } else {
            const unsigned grid = (unsigned)std::min<size_t>((nvec + 511) / 512, (size_t)ctx->prop.multiProcessorCount * 2);
            k_push_block<512><<<grid, 512, 0, st>>>(c->peers, reinterpret_cast<const uint4*>(send), nvec, off, e, comm_err_word(ctx));
        }
        return after_launch(ctx);
    }
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

// The collective and the scatter in one pass: one warp per pixel of THIS rank's shard reads the pixel's packed record once
// and stores it at the pixel's place in the maps of every destination rank (the local one included) - whole rows as 16-byte
// vectors, so every store instruction is one contiguous 352 / 256-byte piece for NVLink.  Warps start at different
// destinations.  The kernel opens with the wait for the destinations' free flags and closes with this rank's done flag.
template <bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS) k_push_records(Peers P, unsigned dst_mask, const int32_t* __restrict__ my_order, int64_t nlocal, int k,
                                                      const unsigned char* __restrict__ send, size_t o_inv, size_t o_sf, size_t o_st, size_t o_iss,
                                                      size_t m_inv, size_t m_sf, size_t m_st, size_t m_iss, unsigned epoch,
                                                      int32_t* err_mapped) {
    block_open(P, epoch, err_mapped);
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int wi = k * 11, ws = k * 8;
    if (VEC && wi <= 128) {
        // nsearch <= 11: a row is at most 32 vectors, one per lane; four pixels per trip so that a lane has eight loads in flight
        constexpr int U = 4;
        const int ni = wi / 4, ns = ws / 4;
        for (int64_t j0 = warp * U; j0 < nlocal; j0 += nwarps * U) {
            int64_t p[U];
            float4 vi[U], vs[U];
#pragma unroll
            for (int u = 0; u < U; ++u) p[u] = j0 + u < nlocal ? (int64_t)my_order[j0 + u] : -1;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p[u] >= 0 && lane < ni) vi[u] = __ldcs(reinterpret_cast<const float4*>(send + o_inv) + (j0 + u) * ni + lane);
                if (p[u] >= 0 && lane < ns) vs[u] = __ldcs(reinterpret_cast<const float4*>(send + o_sf) + (j0 + u) * ns + lane);
            }
            uint8_t stv = 0;
            int32_t isv = 0;
            int64_t pl = -1;                                        // lane u carries the status / issue words of pixel u
#pragma unroll
            for (int u = 0; u < U; ++u) pl = lane == u ? p[u] : pl;
            if (pl >= 0) {
                stv = send[o_st + j0 + lane];
                isv = reinterpret_cast<const int32_t*>(send + o_iss)[j0 + lane];
            }
            const int first = (int)((j0 / U) % P.n);
            for (int q = 0; q < P.n; ++q) {
                const int r = (first + q) % P.n;
                if (!(dst_mask >> r & 1)) continue;
                unsigned char* b = P.base[r];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (p[u] < 0) continue;
                    if (lane < ni) (reinterpret_cast<float4*>(b + m_inv) + p[u] * ni)[lane] = vi[u];
                    if (lane < ns) (reinterpret_cast<float4*>(b + m_sf) + p[u] * ns)[lane] = vs[u];
                }
                if (pl >= 0) {
                    (b + m_st)[pl] = stv;
                    reinterpret_cast<int32_t*>(b + m_iss)[pl] = isv;
                }
            }
        }
    } else
    for (int64_t j = warp; j < nlocal; j += nwarps) {
        const int64_t p = my_order[j];
        const int first = (int)(j % P.n);
        if (VEC) {
            const float4* si = reinterpret_cast<const float4*>(send + o_inv) + j * (wi / 4);
            const float4* ss = reinterpret_cast<const float4*>(send + o_sf) + j * (ws / 4);
            for (int w = lane; w < wi / 4; w += 32) {
                const float4 v = __ldcs(si + w);
                for (int q = 0; q < P.n; ++q) {
                    const int r = (first + q) % P.n;
                    if (dst_mask >> r & 1) (reinterpret_cast<float4*>(P.base[r] + m_inv) + p * (wi / 4))[w] = v;
                }
            }
            for (int w = lane; w < ws / 4; w += 32) {
                const float4 v = __ldcs(ss + w);
                for (int q = 0; q < P.n; ++q) {
                    const int r = (first + q) % P.n;
                    if (dst_mask >> r & 1) (reinterpret_cast<float4*>(P.base[r] + m_sf) + p * (ws / 4))[w] = v;
                }
            }
        } else {
            const float* si = reinterpret_cast<const float*>(send + o_inv) + j * wi;
            const float* ss = reinterpret_cast<const float*>(send + o_sf) + j * ws;
            for (int q = 0; q < P.n; ++q) {
                const int r = (first + q) % P.n;
                if (!(dst_mask >> r & 1)) continue;
                float* di = reinterpret_cast<float*>(P.base[r] + m_inv) + p * wi;
                float* ds = reinterpret_cast<float*>(P.base[r] + m_sf) + p * ws;
                for (int w = lane; w < wi; w += 32) di[w] = si[w];
                for (int w = lane; w < ws; w += 32) ds[w] = ss[w];
            }
        }
        if (lane < P.n && (dst_mask >> lane & 1)) {              // lane r serves destination r
            (P.base[lane] + m_st)[p] = send[o_st + j];
            reinterpret_cast<int32_t*>(P.base[lane] + m_iss)[p] = reinterpret_cast<const int32_t*>(send + o_iss)[j];
        }
    }
    block_done(P, epoch, err_mapped);
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

namespace {
struct Maps { size_t o_inv, o_sf, o_st, o_iss, bytes; };
Maps map_layout(int64_t npix, int k) {
    Maps M;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    M.o_inv = off; off = up(off + (size_t)npix * k * 44);
    M.o_sf = off;  off = up(off + (size_t)npix * k * 32);
    M.o_st = off;  off = up(off + (size_t)npix);
    M.o_iss = off; off = up(off + (size_t)npix * 4);
    M.bytes = off;
    return M;
}
}  // namespace

extern "C" int cledb_map_layout(int64_t npix, int nsearch, int64_t layout[5]) {
    if (!layout || npix < 0 || nsearch < 1) return CLEDB_ERR_ARG;
    const Maps M = map_layout(npix, nsearch);
    layout[0] = (int64_t)M.o_inv; layout[1] = (int64_t)M.o_sf; layout[2] = (int64_t)M.o_st; layout[3] = (int64_t)M.o_iss; layout[4] = (int64_t)M.bytes;
    return CLEDB_OK;
}

// enqueue k_push_records; blocks > 0 bounds the grid (a push that runs beside a search leaves the other SMs to the search)
static int push_records(cledb_ctx* ctx, int64_t npix, int64_t nlocal, const int32_t* my_order_dev, int64_t maxcount, int nsearch,
                        const void* send, void* maps, int root, int blocks, cudaStream_t st, const char* who) {
    cledb_comm* c = ctx->comm;
    if (c->transport != 1) { ctx->err = std::string(who) + ": needs the peer-memory transport (cledb_comm_transport); use cledb_allgather_dev + cledb_unshard_dev"; return CLEDB_ERR_COMM; }
    if (npix < 0 || nlocal < 0 || nlocal > maxcount || nsearch < 1 || !maps || root >= c->nranks || (nlocal > 0 && (!my_order_dev || !send))) {
        ctx->err = std::string(who) + ": bad argument";
        return CLEDB_ERR_ARG;
    }
    const Maps M = map_layout(npix, nsearch);
    bool fits = in_window(c, maps, 0) && ((uintptr_t)maps % 256) == 0;
    if (fits) {
        const size_t woff0 = (size_t)(reinterpret_cast<unsigned char*>(maps) - reinterpret_cast<unsigned char*>(c->win));
        for (int r = 0; r < c->nranks; ++r)
            if (root < 0 || root == r) fits = fits && woff0 + M.bytes <= c->win_bytes[(size_t)r];
    }
    if (!fits) {
        ctx->err = std::string(who) + ": `maps` must be a 256-byte aligned block of cledb_map_layout bytes inside the window (cledb_comm_window)";
        return CLEDB_ERR_ARG;
    }
    const Packed L = packed_layout(maxcount, nsearch);
    const size_t woff = (size_t)(reinterpret_cast<unsigned char*>(maps) - reinterpret_cast<unsigned char*>(c->win));
    const unsigned mask = root < 0 ? (1u << c->nranks) - 1u : 1u << root;
    const unsigned e = ++c->epoch;
    const char* pg = std::getenv("CLEDB_B200_PUSH_GRID");                 // tuning: blocks per SM of the push kernel
    const int per_sm = pg && std::atoi(pg) > 0 ? std::atoi(pg) : 8;
    int64_t grid = std::min<int64_t>((nlocal * 32 + 255) / 256, (int64_t)ctx->prop.multiProcessorCount * per_sm);          // 256-thread blocks
    if (const char* pb = std::getenv("CLEDB_B200_PUSH_BLOCKS"))           // tuning: absolute bound of the grid
        if (blocks == 0 && std::atoi(pb) > 0) blocks = std::atoi(pb);
    if (blocks > 0) grid = std::min<int64_t>(grid, blocks);
    grid = std::max<int64_t>(grid, 1);
    const unsigned char* s8 = reinterpret_cast<const unsigned char*>(send);
    // a narrow push (one that runs beside a search) takes the whole SM it sits on - the search kernel cannot share an SM anyway:
    // 1024 threads per block; the wide one keeps 256-thread blocks at 72 registers
#define CLEDB_PUSH(VEC, T)                                                                                                              \
    k_push_records<VEC, T><<<(unsigned)grid, T, 0, st>>>(c->peers, mask, my_order_dev, nlocal, nsearch, s8, L.o_inv, L.o_sf, L.o_st, L.o_iss, \
                                                         woff + M.o_inv, woff + M.o_sf, woff + M.o_st, woff + M.o_iss, e, comm_err_word(ctx))
    const bool narrow = blocks > 0 && blocks <= 2 * ctx->prop.multiProcessorCount / 3;
    if (nsearch % 4 == 0) { if (narrow) CLEDB_PUSH(true, 1024); else CLEDB_PUSH(true, 256); }
    else                  { if (narrow) CLEDB_PUSH(false, 1024); else CLEDB_PUSH(false, 256); }
#undef CLEDB_PUSH
    return after_launch(ctx);
}

extern "C" int cledb_gather_maps_dev(cledb_ctx* ctx, int64_t npix, int64_t nlocal, const int32_t* my_order_dev, int64_t maxcount,
                                     int nsearch, const void* send, void* maps, int root, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!ctx->comm) { ctx->err = "cledb_gather_maps: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    return push_records(ctx, npix, nlocal, my_order_dev, maxcount, nsearch, send, maps, root, ctx->comm->push_blocks,
                        stream ? (cudaStream_t)stream : ctx->stream, "cledb_gather_maps");
}

// ------------------------------------------------------------------------------------------------------
// search + solutions + gather of a shard as a pipeline: the shard goes through in pieces, and the records of piece i
// cross NVLink (a narrow k_push_records on the communicator's second stream) while piece i+1 is being searched
// ------------------------------------------------------------------------------------------------------
extern "C" int cledb_invert_gather_dev(cledb_ctx* ctx, int mode, int precision, int64_t nlocal, const float* sobs, const float* snr,
                                       const int32_t* db_id, int nsearch, const float* yobs, const float* dobs, const float* rot_cos,
                                       const float* rot_sin, const void* dopp, int dopp_is_f64, const cledb_dbgrid* grid_dev,
                                       double maxchisq, int bcalc, int strict_topk, int64_t npix, const int32_t* my_order_dev, void* maps,
                                       int root, int pieces, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    cledb_comm* c = ctx->comm;
    if (!c) { ctx->err = "cledb_invert_gather: no communicator (cledb_comm_init)"; return CLEDB_ERR_COMM; }
    if (c->transport != 1) { ctx->err = "cledb_invert_gather: needs the peer-memory transport (cledb_comm_transport)"; return CLEDB_ERR_COMM; }
    if (nlocal < 0 || npix < nlocal || pieces < 0 || pieces > kMaxPieces) { ctx->err = "cledb_invert_gather: bad argument"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (pieces == 0) {
        // every rank must run the same number of pushes, so the choice depends on the map and the rank count only.  A search call
        // has a fixed cost of about 0.3 ms (its launches and the tail of its longest work item), a second piece hides half of the
        // NVLink time of the shard: measured on the 1024^2 map that is nothing at 2 and 4 ranks and 9 % at 8 (each rank then
        // receives 7/8 of the map next to a search of 1/8 of it), so two pieces from six ranks on, with pieces that still fill
        // the GPU (64 Ki pixels).  The brute-force searches are two orders of magnitude longer than the push: one piece
        const char* e = std::getenv("CLEDB_B200_PIECES");
        pieces = e && std::atoi(e) > 0 ? std::min(std::atoi(e), kMaxPieces) : (c->nranks >= 6 && npix / c->nranks >= 2 * 65536 - 8 && ctx->prune == 1 ? 2 : 1);
    }
    if (pieces > 1 && !c->side) {
        CLEDB_CUDA(ctx, cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxPieces + 1; ++i) CLEDB_CUDA(ctx, cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
    }
    // packed records of every piece, one block per piece
    const int64_t per = (nlocal + pieces - 1) / pieces;
    const Packed L = packed_layout(std::max<int64_t>(per, 1), nsearch);
    int rc;
    if ((rc = ensure_workspace(ctx, c->send, L.bytes * (size_t)pieces))) return rc;
    const char* sb = std::getenv("CLEDB_B200_PUSH_SIDE");                // tuning: blocks of a push that runs beside a search
    const int side_blocks = sb && std::atoi(sb) > 0 ? std::atoi(sb) : 24;       // 24 SMs at 1024 threads reach the NVLink rate (profiles/r02_tuning.txt)
    const size_t dsz = dopp ? (dopp_is_f64 ? 8 : 4) : 0;
    for (int i = 0; i < pieces; ++i) {                   // a rank with nothing (left) still pushes: the others wait for its flags
        const int64_t j0 = std::min(per * i, nlocal), n = std::min(per, nlocal - j0);
        unsigned char* send = reinterpret_cast<unsigned char*>(c->send.ptr) + L.bytes * (size_t)i;
        if (n > 0) {
            rc = cledb_invert_dev(ctx, mode, precision, n, sobs + j0 * 8, snr + j0 * 8, db_id + j0, nsearch, yobs + j0, dobs + j0, rot_cos + j0,
                                  rot_sin + j0, dopp ? reinterpret_cast<const unsigned char*>(dopp) + (size_t)j0 * dsz : nullptr, dopp_is_f64, 1,
                                  grid_dev, maxchisq, bcalc, strict_topk, reinterpret_cast<float*>(send + L.o_inv),
                                  reinterpret_cast<float*>(send + L.o_sf), send + L.o_st, reinterpret_cast<int32_t*>(send + L.o_iss), st);
            if (rc) return rc;
        }
        cudaStream_t ps = st;
        if (pieces > 1) {                                // every push on the second stream, in order: one operation at a time
            CLEDB_CUDA(ctx, cudaEventRecord(c->ev[i], st));
            CLEDB_CUDA(ctx, cudaStreamWaitEvent(c->side, c->ev[i], 0));
            ps = c->side;
        }
        rc = push_records(ctx, npix, n, my_order_dev + j0, std::max<int64_t>(per, 1), nsearch, send, maps, root,
                          i + 1 < pieces ? side_blocks : 0, ps, "cledb_invert_gather");
        if (rc) return rc;
    }
    if (pieces > 1) {
        CLEDB_CUDA(ctx, cudaEventRecord(c->ev[kMaxPieces], c->side));
        CLEDB_CUDA(ctx, cudaStreamWaitEvent(st, c->ev[kMaxPieces], 0));
    }
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

    // ---- device buffers: this rank's inputs (ctx->io), packed send (/ recv), pixel order (+ bounds), full maps in the window
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t nn = std::max<size_t>(n, 1);
    const size_t o_sobs = take(nn * 32), o_snr = take(nn * 32), o_id = take(nn * 4), o_y = take(nn * 4), o_d = take(nn * 4),
                 o_rc = take(nn * 4), o_rs = take(nn * 4), o_dopp = take(nn * dsz);
    int rc;
    const Maps M = map_layout(npix, nsearch);
    {
        size_t need[kMaxRanks];
        for (int r = 0; r < R; ++r) need[r] = (root < 0 || root == r) ? M.bytes : 256;
        if ((rc = ensure_window(ctx, need))) return rc;           // collective when it has to grow
    }
    const bool peer = c->transport == 1;
    if ((rc = ensure_workspace(ctx, ctx->io, off))) return rc;
    if (!peer && (rc = ensure_workspace(ctx, c->send, L.bytes))) return rc;
    if (!peer && (rc = ensure_workspace(ctx, c->recv, L.bytes * (size_t)R))) return rc;
    const size_t o_bounds = up256((size_t)npix * 4);
    if ((rc = ensure_workspace(ctx, c->order, o_bounds + ((size_t)R + 1) * 8))) return rc;
    const size_t f_inv = M.o_inv, f_sf = M.o_sf, f_st = M.o_st, f_iss = M.o_iss;
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
    cledb_dbgrid g;
    if ((rc = grid_to_device(ctx, grid, &g, st))) return rc;
    unsigned char* dord = reinterpret_cast<unsigned char*>(c->order.ptr);
    if (peer) {                                          // a rank pushes its own pixels: it needs their places, nothing else
        if (n > 0) CLEDB_CUDA(ctx, cudaMemcpyAsync(dord, order + bounds[me], n * 4, cudaMemcpyHostToDevice, st));
    } else if (want) {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(dord, order, (size_t)npix * 4, cudaMemcpyHostToDevice, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(dord + o_bounds, bounds, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, st));
    }

    unsigned char* f = reinterpret_cast<unsigned char*>(c->win);
    if (peer) {
        // ---- search + solutions of the shard in pieces, each piece's records stored at their pixels in the maps of the ranks
        //      that want them while the next piece is searched
        rc = cledb_invert_gather_dev(ctx, mode, precision, nlocal, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_snr),
                                     reinterpret_cast<int32_t*>(w + o_id), nsearch, reinterpret_cast<float*>(w + o_y), reinterpret_cast<float*>(w + o_d),
                                     reinterpret_cast<float*>(w + o_rc), reinterpret_cast<float*>(w + o_rs), dsz ? w + o_dopp : nullptr, dopp_is_f64,
                                     &g, maxchisq, bcalc, strict_topk, npix, reinterpret_cast<int32_t*>(dord), f, root, 0, st);
        if (rc) return rc;        // a rank that fails here leaves the others waiting in the collective: the launcher aborts the job
    } else {
        // ---- search + solutions of the shard, written straight into the packed send buffer; ONE all-gather; scatter
        unsigned char* send = reinterpret_cast<unsigned char*>(c->send.ptr);
        if (n > 0) {
            rc = cledb_invert_dev(ctx, mode, precision, nlocal, reinterpret_cast<float*>(w + o_sobs), reinterpret_cast<float*>(w + o_snr),
                                  reinterpret_cast<int32_t*>(w + o_id), nsearch, reinterpret_cast<float*>(w + o_y), reinterpret_cast<float*>(w + o_d),
                                  reinterpret_cast<float*>(w + o_rc), reinterpret_cast<float*>(w + o_rs), dsz ? w + o_dopp : nullptr, dopp_is_f64, 1,
                                  &g, maxchisq, bcalc, strict_topk, reinterpret_cast<float*>(send + L.o_inv), reinterpret_cast<float*>(send + L.o_sf),
                                  send + L.o_st, reinterpret_cast<int32_t*>(send + L.o_iss), st);
            if (rc) return rc;
        }
        if ((rc = cledb_allgather_dev(ctx, send, c->recv.ptr, (int64_t)L.bytes, st))) return rc;
        if (want) {
            rc = cledb_unshard_dev(ctx, R, reinterpret_cast<int64_t*>(dord + o_bounds), reinterpret_cast<int32_t*>(dord), npix, maxcount, nsearch,
                                   c->recv.ptr, reinterpret_cast<float*>(f + f_inv), reinterpret_cast<float*>(f + f_sf), f + f_st, reinterpret_cast<int32_t*>(f + f_iss), st);
            if (rc) return rc;
        }
    }
    if (want) {
        CLEDB_CUDA(ctx, cudaMemcpyAsync(invout, f + f_inv, (size_t)npix * k * 44, cudaMemcpyDeviceToHost, st));
        CLEDB_CUDA(ctx, cudaMemcpyAsync(sfound, f + f_sf, (size_t)npix * k * 32, cudaMemcpyDeviceToHost, st));
        if (status_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(status_out, f + f_st, (size_t)npix, cudaMemcpyDeviceToHost, st));
        if (issue_out) CLEDB_CUDA(ctx, cudaMemcpyAsync(issue_out, f + f_iss, (size_t)npix * 4, cudaMemcpyDeviceToHost, st));
    }
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    return check_async_error(ctx);
}
