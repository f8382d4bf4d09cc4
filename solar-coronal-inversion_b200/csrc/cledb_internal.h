// Internal declarations shared by the translation units of libcledb_b200.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/cledb_b200.h"

namespace cledb {

constexpr int kTile     = 256;   // database entries per tile = one warp step (32 lanes x 8 entries)
constexpr int kEPL      = 8;     // entries per lane
constexpr int kNC       = 5;     // components that enter chi^2: Q1 U1 I2 Q2 U2 (indices 1 2 4 5 6)
constexpr int kMaxP     = 40;    // pixels per work item
constexpr int kSuper    = 16;    // tiles per group of the two-level bounds (pruning layout)
constexpr int kParts    = 3;     // sign(V1) classes: 0 = positive, 1 = negative, 2 = zero

__host__ __device__ constexpr int chi_comp(int c) { return c == 0 ? 1 : c == 1 ? 2 : c == 2 ? 4 : c == 3 ? 5 : 6; }

// One resident database as the kernels see it (lives in a device array indexed by slot).
struct DbSlot {
    const void*    tiles;            // [ntiles_total][kNC][kTile] half bits (I16) or float (F32)
    const void*    vcomp;            // [npad][2]  V1, V2 in the same element type
    const int32_t* perm;             // [npad] original row index of each stored entry, -1 = padding
    const int32_t* row_pos;          // [n] stored position of an original row, -1 = never a candidate (NaN V1)
    const int32_t* row_kpos;         // [n] rank of a row among the rows of its own sign class (original order)
    const int32_t* kept_rank;        // [n] rank of a row among rows with non-NaN V1 (NULL if none is NaN)
    int64_t        n;                // rows of the original database
    int32_t        part_tile0[kParts];   // first tile of each sign class
    int32_t        part_ntiles[kParts];
    int32_t        part_cnt[kParts];     // entries in each class
    int32_t        part_zero[kParts];    // bit c set: some entry of the class has decoded chi-component c == 0
    int32_t        encoding;
    int32_t        live;
    int32_t        blocked;          // 1: pruning layout - the tiles of a sign class are k-d leaves, `boxes` / `sboxes` hold their bounds
    int32_t        pad_;
    const float*   boxes;            // [ntiles_total][16]: lo[5], hi[5] of the stored (code) values of a tile, one representative entry[5], pad
    const float*   sboxes;           // [sum over classes of ceil(ntiles/16)][12]: lo[5], hi[5], pad[2] of groups of 16 consecutive tiles of a class
    int32_t        part_sb0[kParts]; // first group of each sign class
    int32_t        pad2_;
    float          scale[8];         // 2^-shift per original component (1 for F32)
};

// What one CTA of the search kernel processes: <= kMaxP pixels of one bucket against a tile range.
struct Item {
    int32_t slot;        // database slot
    int32_t tile0;       // first tile (absolute index into DbSlot::tiles)
    int32_t ntiles;      // tiles to scan (may be 0)
    int32_t entry0;      // position of the first entry of the partition (for kpos)
    int32_t pix0;        // first pixel (index into the bucket-sorted order array)
    int32_t npx;         // pixels in this item (<= kMaxP)
    int32_t cnt_end;     // one past the last valid position (absolute) of the partition
    int32_t pad;
};

struct HostDb {
    DbSlot   slot;
    void*    d_tiles = nullptr;
    void*    d_vcomp = nullptr;
    int32_t* d_perm = nullptr;
    int32_t* d_invperm = nullptr;
    int32_t* d_kept_rank = nullptr;
    int32_t* d_row_kpos = nullptr;
    float*   d_rows8 = nullptr;
    float*   d_boxes = nullptr;          // tile bounds, then the bounds of the tile groups (pruning layout only)
    int64_t  npad = 0, nnan = 0, bytes = 0;
    int      shift[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct Workspace {
    void*  ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace cledb

struct cledb_comm;                                 // multi-GPU state (comm.cu)

struct cledb_ctx {
    int device = 0;
    cledb_comm* comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;            // device-to-host copies of finished chunks (big maps)
    cudaDeviceProp prop;
    std::string err;
    std::map<int, int> slot_of_id;                 // db_id -> slot
    std::vector<cledb::HostDb> dbs;                // by slot
    std::vector<int> free_slots;
    cledb::DbSlot* d_slots = nullptr;              // device mirror [slots_cap]
    int32_t* d_slot_of_id = nullptr;               // device table  [id_cap]
    int slots_cap = 0, id_cap = 0;
    bool tables_dirty = true;
    cledb::Workspace ws;                           // search workspace (grown on demand)
    cledb::Workspace ws2;                          // partial top-k lists of the search (sized once the work shape is known)
    cledb::Workspace red;                          // small tables of the reduced search
    cledb::Workspace raw;                          // raw search outputs of the fused search + solution call
    cledb::Workspace io;                           // staging for the host entry points
    float* d_tab = nullptr;                        // sin / cos tables of the grid angles as last uploaded (grid_to_device)
    std::vector<float> tab_host;                   // their host copy: an unchanged grid is not uploaded again
    int32_t* h_cnt = nullptr;                      // mapped pinned host copy of the bucket sizes
    int32_t* d_cnt_mapped = nullptr;               // its device alias
    size_t h_cnt_cap = 0;
    int32_t* h_err = nullptr;                      // mapped pinned word: a searched pixel named a database that is not resident
    int32_t* d_err_mapped = nullptr;               // its device alias
    int last_P = 0, last_split = 0;                // work shape of the last search
    int64_t launches = 0;
    int pixels_per_item = 0;                      // 0 = choose from the pixel count
    int split = 0;
    int dbg = 0;
    int prune = 1;                                 // 0: plain layout, brute force; 1: k-d tiles with bounds, exact pruning (default);
                                                   // 2: k-d tiles, seeded brute force (every pair evaluated).  1 and 2 share the layout
    unsigned long long* d_prune_stats = nullptr;   // [2] evaluated / skipped (pixel, tile) pairs of the last pruned search
    int occ = 0, regs = 0, smem = 0;               // search kernel occupancy facts
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;   // event pairs around the search kernel
    int prof_used = 0;
    bool prof_on = false;
};

#define CLEDB_CUDA(ctx, call)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return CLEDB_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#ifdef __CUDACC__
namespace cledb {
// ------------------------------------------------------------------------------------------------------
// chi^2 of one (pixel, entry) pair, shared by the full and the reduced search
// ------------------------------------------------------------------------------------------------------
// fp32: 10 FFMA.  acc = c0 + sum_c (d_c*a_c + nb_c)^2 = 2 * chi^2 (the halving is exact and done at the end).
__device__ __forceinline__ float eval_fp32(const float (&d)[kNC], const float (&a)[kNC], const float (&nb)[kNC], float c0) {
    float acc = c0;
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const float r = fmaf(d[c], a[c], nb[c]);
        acc = fmaf(r, r, acc);
    }
    return acc;
}
// fp64 validation: float32 subtract and divide, float64 weight/square, NumPy's pairwise order over the 8 terms
// ((t0+t1)+(t2+t3))+((t4+t5)+(t6+t7)) with t3 = t7 = +0, /2, round to 15 decimals as rint(x*1e15)/1e15
// (CLEDB_PROC.py:989-1005).  d must already be the decoded database value.
__device__ __forceinline__ double eval_fp64(const float (&d)[kNC], const float (&st)[kNC], const float (&sig)[kNC], double c0) {
    double q[kNC];
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const float r = __fdiv_rn(__fsub_rn(d[c], st[c]), sig[c]);
        const double t = (c == 2) ? __dmul_rn((double)r, 0.01) : (double)r;
        q[c] = __dmul_rn(t, t);
    }
    const double left = __dadd_rn(__dadd_rn(c0, q[0]), q[1]);
    const double right = __dadd_rn(__dadd_rn(q[2], q[3]), q[4]);
    const double half = __dmul_rn(__dadd_rn(left, right), 0.5);
    return __ddiv_rn(rint(__dmul_rn(half, 1e15)), 1e15);
}

}  // namespace cledb
#endif

namespace cledb {
int  ensure_workspace(cledb_ctx* ctx, Workspace& w, size_t bytes);
int  sync_tables(cledb_ctx* ctx, cudaStream_t st);
int  match_device(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                  const int32_t* db_id, int nsearch, int32_t* idx_out, float* chi_out, double* chi64_out,
                  int32_t* kpos_out, float* rows_out, uint8_t* status_out, int64_t* ncand_out, cudaStream_t st);
int  match_occupancy(cledb_ctx* ctx);
int  check_async_error(cledb_ctx* ctx);            // after a host synchronisation: CLEDB_ERR_NODB if a search met an unknown database id
int  launch_decode_db(cledb_ctx* ctx, int slot, float* d_out, cudaStream_t st);
int  launch_gather_rows(cledb_ctx* ctx, int slot, const int32_t* d_idx, int64_t n, float* d_rows, cudaStream_t st);
int  launch_blos(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac,
                 float* out, uint8_t* flags, cudaStream_t st);
int  measure_fp32_peak(cledb_ctx* ctx, double* tflops);
int  launch_qurotate(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref,
                     const float* hpxy, const float* in, float* out, float* aobs, float* yobs, cudaStream_t st);
}  // namespace cledb

namespace cledb {
// *g_dev = *grid with the four table pointers replaced by device copies (kept on the context; uploaded on `st` only when the
// caller's tables differ from the ones already there).  One in-flight call per context, as everywhere.
int grid_to_device(cledb_ctx* ctx, const cledb_dbgrid* grid, cledb_dbgrid* g_dev, cudaStream_t st);
int launch_build(cledb_ctx* ctx, int mode, int precision, int64_t npix, int nsearch, const float* sobs, const float* yobs,
                 const float* dobs, const float* rot_cos, const float* rot_sin, const void* dopp, int dopp_is_f64,
                 int64_t dopp_stride, const cledb_dbgrid& grid_dev, double maxchisq, int bcalc, int strict_topk,
                 const int32_t* idx, const float* chi, const double* chi64, const int32_t* kpos, const float* rows,
                 const uint8_t* status, const int32_t* nan_slots, float* invout, float* sfound, cudaStream_t st,
                 const float* snr = nullptr, int32_t* issue_out = nullptr);
}  // namespace cledb
