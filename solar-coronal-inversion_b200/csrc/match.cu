// The 2-line database search on sm_100a: classification/bucketing of pixels, the chi^2 + top-nsearch kernel
// and the merge/finalise kernel.  Replaces the per-pixel body of cledb_matchiquv / cledb_matchiqud
// (CLEDB_PROC.py:944-1013 and 1118-1188 of the reference) for a whole map at once.
//
// Work decomposition:
//   bucket  = all pixels that search the same candidate list: (database, sign(V_obs)) for IQUV, database for IQUD
//   item    = <= 40 pixels of one bucket  x  one contiguous tile range of one sign partition of the database
//   warp    = the unit of work: a warp fetches an item, owns its pixels exclusively and streams ALL tiles of the
//             item past them.  Nothing is shared between warps: no locks, no queues, no __syncthreads.
//             A step handles two tiles: a lane holds 2 x 8 database entries (5 chi^2 components each, decoded from
//             int16 in registers); per pixel 3 broadcast LDS.128 (10 coefficients + flag start value + threshold tau)
//             feed 80 packed FP32 FMAs (fma.rn.f32x2).
//   top-k   = per pixel a sorted list of nsearch (value, position) pairs in the warp's shared-memory slice, kept
//             distributed over the lanes while inserting (warp shuffles).  The k-th value tau is the pruning
//             threshold: the hot loop only sets a flag bit per (pixel, tile) - the sign of an accumulator started at
//             c0 - tau - and flagged pixels are re-evaluated exactly and inserted right after the pixel loop, so tau
//             is exact again before the pixel is visited for the next step.
//   TMA     = database tiles are contiguous 2.5 KB (int16) / 5 KB (fp32) blocks moved global->shared with
//             cp.async.bulk + mbarrier complete_tx, double buffered per warp.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "cledb_internal.h"

namespace cledb {

// ------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 128-bit shared-memory load from a 32-bit shared address (keeps the hot loop free of generic->shared conversions)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------------
// per-pixel records and per-list entries for the two arithmetic variants
// ------------------------------------------------------------------------------------------------------
template <int PREC> struct Rec;
// fp32 production: residual r_c = d_c * a_c + nb_c with a_c = w_c/sigma_c (x 2^-shift), nb_c = -a_c * s~_c
template <> struct __align__(16) Rec<CLEDB_PREC_FP32> {
    float a[kNC];
    float nb[kNC];
    float cm;    // in HBM: c0, the start value of the accumulation chain: 0 (the constant component-0 term
                 // ((1 - s~_0)/sigma_0)^2 is kept apart in Plan::c0 and added when chi^2 is written).  In shared memory:
                 // c0 - tau*(1 + 2^-19), the start value of the flagging accumulators (c0 itself moves to a side array)
    float tau;   // running threshold: k-th best value so far (same units as the accumulators)
};
// Flagging without a min tree: the hot loop accumulates v' = (c0 - tau_m) + sum r^2 with tau_m = tau*(1 + 2^-19), so
// that "some entry may beat the threshold" is the OR of the sign bits of the eight v'.  The 32-ulp margin covers the
// rounding of the two different summation chains (each at most 6 half-ulps of max(v, tau)): v <= tau implies v' < 0.
// Flagged pixels are re-evaluated exactly (accumulators started at c0) before anything is inserted.
__device__ __forceinline__ float flag_start(float c0, float tau) {
    return c0 - fmaxf(fmaf(tau, 1.9073486328125e-06f, tau), 1e-37f);
}
// fp64 validation: the reference's operation order, needs s~ and sigma themselves
template <> struct __align__(16) Rec<CLEDB_PREC_FP64> {
    float  st[kNC];
    float  sig[kNC];
    float  pad[2];
    double c0;
    double tau;
};
template <int PREC> struct Val { using type = float; };
template <> struct Val<CLEDB_PREC_FP64> { using type = double; };

// one top-k list element: value and position in the stored (stride-permuted) layout, -1 = empty slot
template <int PREC> struct __align__(8) Cand { float v; int32_t pos; };
template <> struct __align__(16) Cand<CLEDB_PREC_FP64> { double v; int32_t pos; int32_t pad; };

template <int ENC> struct EncTraits;
template <> struct EncTraits<CLEDB_ENC_I16> { static constexpr int kElem = 2; };
template <> struct EncTraits<CLEDB_ENC_F32> { static constexpr int kElem = 4; };
template <int ENC> __host__ __device__ constexpr int tile_bytes() { return kNC * kTile * EncTraits<ENC>::kElem; }

// One shared-memory stage per warp: a step's tiles are copied into registers at once, the refill is issued right after and
// is in flight during the step's compute (13 us for a 36-pixel item), so the registers are the second buffer.
constexpr int kStages = 1;

// position inside a tile of entry e of a lane.  Chosen so that the lane's shared-memory reads are 16-byte
// vectors at a 16-byte lane stride (conflict free): int16 -> 8 consecutive entries per lane and component,
// fp32 -> two groups of 4 consecutive entries, 128 entries apart.
template <int ENC> __device__ __forceinline__ int pos_in_tile(int lane, int e) {
    if (ENC == CLEDB_ENC_I16) return lane * 8 + e;
    return (e >> 2) * 128 + lane * 4 + (e & 3);
}

// decode one tile from shared memory into registers: d[e][c] for the lane's 8 entries
template <int ENC> __device__ __forceinline__ void load_tile_regs(const unsigned char* stage, int lane, float (&d)[kEPL][kNC]) {
    if (ENC == CLEDB_ENC_I16) {
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
            const uint4 w = *reinterpret_cast<const uint4*>(stage + (c * kTile + lane * 8) * 2);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 h = *reinterpret_cast<const __half2*>(&ww[j]);
                const float2 f = __half22float2(h);   // exact: the int16 code is a binary16 bit pattern
                d[2 * j][c] = f.x;
                d[2 * j + 1][c] = f.y;
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
            const float4 lo = *reinterpret_cast<const float4*>(stage + (c * kTile + lane * 4) * 4);
            const float4 hi = *reinterpret_cast<const float4*>(stage + (c * kTile + 128 + lane * 4) * 4);
            d[0][c] = lo.x; d[1][c] = lo.y; d[2][c] = lo.z; d[3][c] = lo.w;
            d[4][c] = hi.x; d[5][c] = hi.y; d[6][c] = hi.z; d[7][c] = hi.w;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// chi^2 of one (pixel, entry) pair
// ------------------------------------------------------------------------------------------------------
// The same arithmetic for two entries at once with the packed FP32 FMA of sm_100 (fma.rn.f32x2, FFMA2 in SASS): one
// issue slot per two FMAs.  The coefficients are scalars; ptxas encodes pack(a, a) as a broadcast .F32 operand.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void eval_fp32_pair(const float (&d0)[kNC], const float (&d1)[kNC], const float (&a)[kNC],
                                               const float (&nb)[kNC], float c0, float& v0, float& v1) {
    unsigned long long acc = pack2(c0, c0);
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const unsigned long long r = fma2(pack2(d0[c], d1[c]), pack2(a[c], a[c]), pack2(nb[c], nb[c]));
        acc = fma2(r, r, acc);
    }
    unpack2(acc, v0, v1);
}
// all 8 entries of a lane, bit-identical to eval_fp32 per entry
__device__ __forceinline__ void eval_fp32_all(const float (&d)[kEPL][kNC], const float (&a)[kNC], const float (&nb)[kNC], float c0,
                                              float (&v)[kEPL]) {
#pragma unroll
    for (int e = 0; e < kEPL; e += 2) eval_fp32_pair(d[e], d[e + 1], a, nb, c0, v[e], v[e + 1]);
}
// ------------------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {   // values are >= +0 or +inf: the bit patterns order like uints
    return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float  inf_of(float)  { return CUDART_INF_F; }
__device__ __forceinline__ double inf_of(double) { return CUDART_INF; }

// Insert candidate (v, pos) into the sorted list of one pixel.  Called by a whole warp with warp-uniform
// arguments; the list belongs to this warp alone.  Lane j < k holds list element j.  Order: ascending value; equal
// values by ascending ORIGINAL row (NumPy argmin = first occurrence).  Lists store positions in the stride-permuted
// layout, so the original rows are looked up (perm[]) only in the rare case that values are exactly equal.
template <int PREC> __device__ __forceinline__ void set_tau(Rec<PREC>* rec, float c0, typename Val<PREC>::type tau);
template <> __device__ __forceinline__ void set_tau<CLEDB_PREC_FP32>(Rec<CLEDB_PREC_FP32>* rec, float c0, float tau) {
    rec->tau = tau;
    rec->cm = flag_start(c0, tau);
}
template <> __device__ __forceinline__ void set_tau<CLEDB_PREC_FP64>(Rec<CLEDB_PREC_FP64>* rec, float, double tau) { rec->tau = tau; }

template <int PREC>
__device__ __forceinline__ void list_insert(Cand<PREC>* list, Rec<PREC>* rec, float c0, int k,
                                            typename Val<PREC>::type v, int pos, const int32_t* __restrict__ perm, int lane) {
    using V = typename Val<PREC>::type;
    V lv = inf_of(V());
    int lp = -1;
    if (lane < k) {
        lv = list[lane].v;
        lp = list[lane].pos;
    }
    bool before = (lane < k) && (lv < v);
    const bool equal = (lane < k) && (lv == v) && lp >= 0;
    if (__any_sync(0xffffffffu, equal)) {
        const int cand_row = perm[pos];
        if (equal) before = perm[lp] < cand_row;
    }
    const int q = __popc(__ballot_sync(0xffffffffu, before));   // elements that stay in front (a prefix: list is sorted)
    if (q < k) {
        const V uv = __shfl_up_sync(0xffffffffu, lv, 1);
        const int up = __shfl_up_sync(0xffffffffu, lp, 1);
        if (lane == q) {
            lv = v;
            lp = pos;
        } else if (lane > q) {
            lv = uv;
            lp = up;
        }
        if (lane >= q && lane < k) {
            list[lane].v = lv;
            list[lane].pos = lp;
        }
        if (lane == k - 1) set_tau<PREC>(rec, c0, lv);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// device-side bookkeeping shared by the kernels
// ------------------------------------------------------------------------------------------------------
struct Plan {                 // all pointers are device pointers into the workspace
    int32_t* key;             // [npix]  bucket of a pixel, -1 = not searched
    int32_t* rank;            // [npix]  position of the pixel in `order`
    int32_t* order;           // [npix]  pixels sorted by bucket
    int32_t* cnt;             // [nkeys] pixels per bucket
    int32_t* cursor;          // [nkeys]
    int32_t* off;             // [nkeys] first position of the bucket in `order`
    int32_t* item0;           // [nkeys+1] first item of the bucket; item0[nkeys] = total items
    int32_t* nsub;            // [nkeys] candidate-list pieces (sign partitions x chunks) per pixel tile
    int32_t* counters;        // [0] next item to hand out, [1] error flag, [2] total items
    int32_t* err_mapped;      // mapped host word: set when a searched pixel names a database that is not resident
    void*    recs;            // [npix] Rec<PREC>
    double*  c0;              // [npix] fp32 search: the per-pixel constant term of chi^2*2 (I1 component), added when the result is written
    void*    plist;           // [maxitems][P][k] Cand<PREC>
    uint8_t* status;          // [npix]
    int32_t  nkeys;
    int32_t  P;               // pixels per item
    int32_t  split;           // chunks per sign partition
    int32_t  mode;
    int32_t  k;
    int32_t  dbg;             // timing experiments only: 1 = no list lock, 2 = skip the rare pass (results invalid)
    int64_t  npix;
    // exact tile pruning only (NULL otherwise)
    int32_t* seed;            // [npix][kParts] tile (relative to its sign class) whose representative entry fits the pixel best, -1 = none
    int32_t* fcnt;            // [slots * ftiles] pixels per (database, seed tile): the fine sort key inside a bucket
    int32_t* fcursor;         // [slots * ftiles]
    int32_t* foff;            // [slots * ftiles] first position of the fine bucket in `order`
    int32_t  ftiles;          // fine keys per database slot (>= tiles of the largest resident database)
    int32_t  fused;           // pruned IQUD search: one item per pixel group walks every sign class with one list per pixel
    // longest-items-first queue of the pruned search (NULL: items are taken in index order)
    float*   wpix;            // [npix] predicted weight of a pixel: chi^2 of the best tile representative (k_seed)
    int32_t* qbin;            // [max items] weight bin of an item
    int32_t* qorder;          // [qocap] queue entries, heaviest bin first: item | piece << 24 (piece 0 = the whole item), -1 = hole
    int32_t* qhist;           // [64] items per bin, [64] cursors, [2] queue length and split bin
    int64_t  qcap;            // capacity of qbin
    int64_t  qocap;           // capacity of qorder
    int32_t  qsplit;          // heavy items are handed out in pieces of qpiece pixels when their bin is >= median bin + qsplit (0: never)
    int32_t  qpiece;
};

// the sign class with the most entries (ties: the lowest class)
__device__ __forceinline__ int largest_class(const DbSlot& s) {
    int q = 0;
    if (s.part_cnt[1] > s.part_cnt[q]) q = 1;
    if (s.part_cnt[2] > s.part_cnt[q]) q = 2;
    return q;
}
__device__ __forceinline__ int nonempty_parts(const DbSlot& s) {
    return (s.part_cnt[0] > 0) + (s.part_cnt[1] > 0) + (s.part_cnt[2] > 0);
}

// ------------------------------------------------------------------------------------------------------
// kernel 1: classify pixels, build their records, count bucket sizes
// ------------------------------------------------------------------------------------------------------
template <int PREC>
__global__ void __launch_bounds__(256) k_classify(Plan pl, const float* __restrict__ sobs, const float* __restrict__ snr,
                                                   const int32_t* __restrict__ db_id, const DbSlot* __restrict__ slots,
                                                   const int32_t* __restrict__ slot_of_id, int id_cap) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pl.npix) return;
    float s[8], n[8];
    {
        const float4 a = reinterpret_cast<const float4*>(sobs)[2 * p], b = reinterpret_cast<const float4*>(sobs)[2 * p + 1];
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
        const float4 c = reinterpret_cast<const float4*>(snr)[2 * p], d = reinterpret_cast<const float4*>(snr)[2 * p + 1];
        n[0] = c.x; n[1] = c.y; n[2] = c.z; n[3] = c.w; n[4] = d.x; n[5] = d.y; n[6] = d.z; n[7] = d.w;
    }
    pl.key[p] = -1;
    const int id = db_id[p];
    int slot = (id >= 0 && id < id_cap) ? slot_of_id[id] : -1;
    int nnan = 0, nzero = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        nnan += isnan(s[c]) ? 1 : 0;
        nzero += (s[c] == 0.0f) ? 1 : 0;
    }
    // null pixel: CLEDB_PROC.py:944 (.any()) for IQUV, :1118 (.all()) for IQUD
    const bool null_pix = (pl.mode == CLEDB_MODE_IQUV) ? (nnan > 0 || nzero == 8) : (nnan == 8 || nzero == 8);
    if (null_pix) {
        pl.status[p] = CLEDB_PIX_NULL;
        return;
    }
    if (slot < 0 || !slots[slot].live) {     // only searched pixels need their database
        atomicExch(&pl.counters[1], 1);
        *reinterpret_cast<volatile int32_t*>(pl.err_mapped) = 1;      // the host reads it after its next synchronisation
        __threadfence_system();
        pl.status[p] = CLEDB_PIX_NULL;
        return;
    }
    const DbSlot& db = slots[slot];
    if (pl.mode == CLEDB_MODE_IQUD && nnan > 0) {   // np.max is NaN -> argwhere()[0,0] raises (CLEDB_PROC.py:1155)
        pl.status[p] = CLEDB_PIX_REFRAISES;
        return;
    }
    int cls = 0;
    float nf;
    if (pl.mode == CLEDB_MODE_IQUV) {
        cls = s[3] > 0.0f ? 0 : (s[3] < 0.0f ? 1 : 2);          // np.sign(sobs[3]) against np.sign(D[:,3]), :974
        nf = s[0];                                               // :980
        if (db.part_cnt[cls] == 0) {                             // empty candidate list: argmin of empty raises
            pl.status[p] = CLEDB_PIX_REFRAISES;
            return;
        }
    } else {
        nf = s[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) nf = fmaxf(nf, s[c]);       // :1155 (no NaN here)
        if (nonempty_parts(db) == 0) {
            pl.status[p] = CLEDB_PIX_REFRAISES;
            return;
        }
    }
    float st[8], sig[8];
    bool allnan = false, allinf = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        st[c] = __fdiv_rn(s[c], nf);                             // sobs/norm_fact, float32
        sig[c] = __fdiv_rn(st[c], n[c]);                         // sobs/norm_fact/snr, float32
        if (!isfinite(st[c]) || isnan(sig[c])) allnan = true;
    }
    if (sig[3] == 0.0f || sig[7] == 0.0f) allnan = true;         // x/0 = inf or NaN, times weight 0 = NaN
    if (sig[0] == 0.0f) allinf = true;                           // (1 - 0)/0 = inf with weight 1
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const int oc = chi_comp(c);
        if (sig[oc] == 0.0f) {
            // (D - s~)/0 is +-inf, or NaN for an entry with D == s~ (== 0): argmin then lands on the NaN
            bool has_equal = false;
            if (st[oc] == 0.0f) {
                if (pl.mode == CLEDB_MODE_IQUV) has_equal = (db.part_zero[cls] >> c) & 1;
                else has_equal = ((db.part_zero[0] | db.part_zero[1] | db.part_zero[2]) >> c) & 1;
            }
            if (has_equal) allnan = true; else allinf = true;
        }
    }
    if (allnan) { pl.status[p] = CLEDB_PIX_ALLNAN; return; }
    if (allinf) { pl.status[p] = CLEDB_PIX_ALLINF; return; }

    const float r0 = __fdiv_rn(__fsub_rn(1.0f, st[0]), sig[0]);  // component 0 of the database is exactly 1
    const double c0 = (double)r0 * (double)r0;
    Rec<PREC>* rec = reinterpret_cast<Rec<PREC>*>(pl.recs) + p;
    if (PREC == CLEDB_PREC_FP32) {
        Rec<CLEDB_PREC_FP32> r;
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
            const int oc = chi_comp(c);
            const double w = (oc == 4) ? 0.01 : 1.0;             // wtg_fact, CLEDB_PROC.py:981
            const float a = (float)(w / (double)sig[oc]);
            r.nb[c] = (float)(-(double)a * (double)st[oc]);
            r.a[c] = a * db.scale[oc];                           // fold the int16 block exponent (exact)
        }
        // The I1 term c0 is the same for every candidate of the pixel, so the fp32 search ranks the candidates by the sum of the
        // other five terms and c0 is added (in double) when chi^2 is written: in IQUD mode c0 can be 10^7 times those terms
        // (I1 far below max(sobs) in a bad pixel), and accumulating on top of it would round thousands of candidates to the
        // same float - wrong order against the reference's float64, and every one of those ties resolved by row lookups.
        r.cm = 0.f;                                            // start value of the accumulation chain, see Rec
        pl.c0[p] = c0;
        r.tau = CUDART_INF_F;
        *reinterpret_cast<Rec<CLEDB_PREC_FP32>*>(rec) = r;
    } else {
        Rec<CLEDB_PREC_FP64> r;
#pragma unroll
        for (int c = 0; c < kNC; ++c) {
            r.st[c] = st[chi_comp(c)];
            r.sig[c] = sig[chi_comp(c)];
        }
        r.pad[0] = r.pad[1] = 0.f;
        r.c0 = c0;
        r.tau = CUDART_INF;
        *reinterpret_cast<Rec<CLEDB_PREC_FP64>*>(rec) = r;
    }
    const int key = slot * kParts + cls;
    pl.key[p] = key;
    pl.status[p] = CLEDB_PIX_SEARCHED;
    // warp-aggregated count: the lanes that reach this point with the same bucket send one atomic
    const uint32_t peers = __match_any_sync(__activemask(), key);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&pl.cnt[key], __popc(peers));
}

// ------------------------------------------------------------------------------------------------------
// kernel 2: one block scans the bucket sizes into pixel offsets and item offsets
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan(Plan pl, const DbSlot* __restrict__ slots) {
    __shared__ int s_pix[1024], s_item[1024];
    const int tid = threadIdx.x;
    const int per = (pl.nkeys + 1023) / 1024;
    const int k0 = tid * per, k1 = min(pl.nkeys, k0 + per);
    int pix = 0, items = 0;
    for (int key = k0; key < k1; ++key) {
        const int c = pl.cnt[key];
        int ns = 0;
        if (c > 0) {
            const DbSlot& db = slots[key / kParts];
            ns = (pl.mode == CLEDB_MODE_IQUV) ? pl.split : (pl.fused ? 1 : nonempty_parts(db) * pl.split);
        }
        pl.nsub[key] = ns;
        pix += c;
        items += ((c + pl.P - 1) / pl.P) * ns;
    }
    s_pix[tid] = pix;
    s_item[tid] = items;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {        // Hillis-Steele inclusive scan
        int a = 0, b = 0;
        if (tid >= o) { a = s_pix[tid - o]; b = s_item[tid - o]; }
        __syncthreads();
        s_pix[tid] += a;
        s_item[tid] += b;
        __syncthreads();
    }
    int po = s_pix[tid] - pix, io = s_item[tid] - items;
    for (int key = k0; key < k1; ++key) {
        const int c = pl.cnt[key];
        pl.off[key] = po;
        pl.item0[key] = io;
        po += c;
        io += ((c + pl.P - 1) / pl.P) * pl.nsub[key];
    }
    if (tid == 1023) {
        pl.item0[pl.nkeys] = s_item[1023];
        pl.counters[2] = s_item[1023];
    }
}

// kernel 3: scatter pixels into bucket order
__global__ void __launch_bounds__(256) k_scatter(Plan pl) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pl.npix) return;
    const int key = pl.key[p];
    if (key < 0) { pl.rank[p] = -1; return; }
    // warp-aggregated slot allocation inside the bucket
    const int lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1;
    int first = 0;
    if (lane == leader) first = atomicAdd(&pl.cursor[key], __popc(peers));
    first = __shfl_sync(peers, first, leader);
    const int r = pl.off[key] + first + __popc(peers & ((1u << lane) - 1u));
    pl.order[r] = (int32_t)p;
    pl.rank[p] = r;
}

// item index -> descriptor (binary search over the per-bucket item offsets)
__device__ __forceinline__ Item make_item(const Plan& pl, const DbSlot* slots, int it) {
    int lo = 0, hi = pl.nkeys;                  // largest key with item0[key] <= it (and items in it)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pl.item0[mid] <= it) lo = mid; else hi = mid;
    }
    const int key = lo;
    const int ns = pl.nsub[key];
    const int local = it - pl.item0[key];
    // Items of a bucket: pixel tiles x pieces.  The brute-force search numbers them tile-major; the pruned search piece-major,
    // so that warps working side by side are in the same sign class (IQUD: the tiny V = 0 class takes other code paths than
    // the two big ones, and the kernel is sensitive to its instruction-cache footprint).
    const int ntg = (pl.cnt[key] + pl.P - 1) / pl.P;
    int tile = local / ns, sub = local - tile * ns;
    if (pl.seed) { sub = local / ntg; tile = local - sub * ntg; }
    const DbSlot& db = slots[key / kParts];
    int part = key % kParts, chunk = sub;
    if (pl.mode == CLEDB_MODE_IQUD) {
        int which = sub / pl.split;
        chunk = sub - which * pl.split;
        part = 0;
        for (int q = 0; q < kParts; ++q) {
            if (db.part_cnt[q] > 0) {
                if (which == 0) { part = q; break; }
                --which;
            }
        }
    }
    const int T = db.part_ntiles[part];
    const int t0 = (int)(((int64_t)T * chunk) / pl.split), t1 = (int)(((int64_t)T * (chunk + 1)) / pl.split);
    Item im;
    im.slot = key / kParts;
    im.tile0 = db.part_tile0[part] + t0;
    im.ntiles = t1 - t0;
    im.entry0 = db.part_tile0[part] * kTile;
    im.pix0 = pl.off[key] + tile * pl.P;
    im.npx = min(pl.P, pl.cnt[key] - tile * pl.P);
    im.cnt_end = im.entry0 + db.part_cnt[part];
    im.pad = 0;
    return im;
}

// ------------------------------------------------------------------------------------------------------
// kernel 4: the search
// ------------------------------------------------------------------------------------------------------
constexpr int kListCap = 320;     // list entries per warp: pixels per item x nsearch <= 320 (40 pixels at nsearch 8)

template <int ENC, int PREC> struct KCfg {
    static constexpr int kWpc = PREC == CLEDB_PREC_FP32 ? 12 : 8;       // independent warps per CTA (shared memory budget)
    static constexpr int kTB = tile_bytes<ENC>();
    // tiles per step: the fp32 kernel holds two tiles (16 entries per lane) in registers, so that a pixel record loaded
    // from shared memory feeds 80 packed FMAs instead of 40 (hot-loop ceiling 82 % -> 87 %, profiles/r01_ubench_hot.txt)
    static constexpr int kTPS = PREC == CLEDB_PREC_FP32 ? 2 : 1;
    static constexpr int kStageBytes = kTPS * kTB;
    static constexpr int kRecBytes = (kMaxP + 1) * (int)sizeof(Rec<PREC>);   // one spare record: prefetch overrun
    static constexpr int kListBytes = kListCap * (int)sizeof(Cand<PREC>);
    static constexpr int kC0Bytes = ((kMaxP + 3) / 4) * 16;                    // c0 of the item's pixels (fp32 kernel)
    // seeded brute force (k-d layout, fp32): seed tile of every pixel of the item + one selection's candidates
    static constexpr int kSeedBytes = PREC == CLEDB_PREC_FP32 ? ((kMaxP * 4 + 15) / 16 * 16 + 32 * 8) : 0;
    // every warp slice starts 128-byte aligned (bulk-copy destination, 16-byte vector loads)
    static constexpr int kWarpBytes = (kStages * kStageBytes + kRecBytes + kListBytes + kStages * 8 + 16 + kC0Bytes + kSeedBytes + 127) / 128 * 128;
};

// evaluate the lane's 8 entries against the pixel record in shared memory; returns the 8 values
template <int ENC, int PREC>
__device__ __forceinline__ void eval_pixel(const Rec<PREC>* rec, float c0f, const float (&d)[kEPL][kNC], const float (&scale)[kNC],
                                           typename Val<PREC>::type (&v)[kEPL], typename Val<PREC>::type& tau) {
    if (PREC == CLEDB_PREC_FP32) {
        const float4* rp = reinterpret_cast<const float4*>(rec);
        const float4 x0 = rp[0], x1 = rp[1], x2 = rp[2];          // three broadcast LDS.128
        const float a[kNC] = {x0.x, x0.y, x0.z, x0.w, x1.x};
        const float nb[kNC] = {x1.y, x1.z, x1.w, x2.x, x2.y};
        float vf[kEPL];
        eval_fp32_all(d, a, nb, c0f, vf);
#pragma unroll
        for (int e = 0; e < kEPL; ++e) v[e] = vf[e];
        tau = x2.w;
    } else {
        const Rec<CLEDB_PREC_FP64>* r = reinterpret_cast<const Rec<CLEDB_PREC_FP64>*>(rec);
        float st[kNC], sig[kNC];
#pragma unroll
        for (int c = 0; c < kNC; ++c) { st[c] = r->st[c]; sig[c] = r->sig[c]; }
        const double c0 = r->c0;
#pragma unroll
        for (int e = 0; e < kEPL; ++e) {
            float dd[kNC];
#pragma unroll
            for (int c = 0; c < kNC; ++c) dd[c] = d[e][c] * scale[c];   // decoded database value (exact)
            v[e] = eval_fp64(dd, st, sig, c0);
        }
        tau = r->tau;
    }
}

// Entries whose value EQUALS the threshold enter the list only if their original row is lower than the row of the list's
// last element (first-occurrence ties).  A pixel whose candidates all tie - every chi^2 the same float, e.g. an IQUD pixel
// with max(sobs) ~ 0, or +inf everywhere - would otherwise send every entry of every tile through list_insert; this filter
// looks the rows up once and drops the ties that cannot win (they are set to NaN: never a hit).
template <int ENC, typename V>
__device__ __forceinline__ void drop_losing_ties(V (&v)[kEPL], V tau, const Cand<(sizeof(V) == 8 ? CLEDB_PREC_FP64 : CLEDB_PREC_FP32)>* list,
                                                 int base, int cnt_end, const int32_t* __restrict__ perm, int k, int lane) {
    uint32_t tied = 0;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) tied |= (v[e] == tau ? 1u : 0u) << e;
    if (!__any_sync(0xffffffffu, tied != 0u)) return;
    const int last = list[k - 1].pos;
    const int kth_row = last >= 0 ? perm[last] : 0x7fffffff;
#pragma unroll
    for (int e = 0; e < kEPL; ++e)
        if ((tied >> e) & 1u) {
            const int pos = base + pos_in_tile<ENC>(lane, e);
            if (pos >= cnt_end || perm[pos] >= kth_row) v[e] = (V)CUDART_NAN;
        }
}

// exact, unhurried handling of one (tile, pixel): evaluate, insert every entry that beats the live threshold.
// Used by the fp64 validation kernel for every step and by the fp32 kernel when a candidate buffer overflows.
template <int ENC, int PREC, bool TIE_FILTER = true>
__device__ __forceinline__ void slow_pixel(Rec<PREC>* rec, float c0f, Cand<PREC>* list, const float (&d)[kEPL][kNC],
                                           const float (&scale)[kNC], int base, int cnt_end, const int32_t* __restrict__ perm,
                                           int k, int lane) {
    using V = typename Val<PREC>::type;
    V v[kEPL], tau;
    eval_pixel<ENC, PREC>(rec, c0f, d, scale, v, tau);
    uint32_t em = 0;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) em |= (v[e] <= tau) ? (1u << e) : 0u;        // padding evaluates to NaN: never true
    uint32_t lanes = __ballot_sync(0xffffffffu, em != 0);
    if (TIE_FILTER && __popc(lanes) > 4) {        // many hits at once: usually ties with the threshold that cannot win
        drop_losing_ties<ENC, V>(v, tau, list, base, cnt_end, perm, k, lane);
        em = 0;
#pragma unroll
        for (int e = 0; e < kEPL; ++e) em |= (v[e] <= tau) ? (1u << e) : 0u;
        lanes = __ballot_sync(0xffffffffu, em != 0);
    }
    while (lanes) {
        const int src = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        uint32_t m = __shfl_sync(0xffffffffu, em, src);
        while (m) {
            const int e = __ffs(m) - 1;
            m &= m - 1;
            // v[e] for a warp-uniform runtime e, kept in registers: a three-level select tree on the bits of e
            const bool b0 = e & 1, b1 = e & 2, b2 = e & 4;
            const V m01 = b0 ? v[1] : v[0], m23 = b0 ? v[3] : v[2], m45 = b0 ? v[5] : v[4], m67 = b0 ? v[7] : v[6];
            const V m03 = b1 ? m23 : m01, m47 = b1 ? m67 : m45;
            const V mine = b2 ? m47 : m03;
            const V cv = __shfl_sync(0xffffffffu, mine, src);
            const int cp = base + pos_in_tile<ENC>(src, e);
            if (cp < cnt_end) list_insert<PREC>(list, rec, c0f, k, cv, cp, perm, lane);   // re-checks against the list
        }
    }
}

// ---- selection for the pruned search ------------------------------------------------------------------------------------
// rank_select: new list = the k smallest of (current list) + (the lane's entries with v[e] < inf), when they fit into one warp
// (<= 32 candidates) and no two candidates have exactly the same value: every candidate goes to one lane, a lane's rank is
// the number of smaller candidates, rank r < k is list element r.  Returns false (nothing changed) otherwise; the callers
// then take the paths that know how to break exact ties by original row.
template <int ENC>
__device__ __forceinline__ bool rank_select(Rec<CLEDB_PREC_FP32>* rec, float c0f, Cand<CLEDB_PREC_FP32>* list, const float (&v)[kEPL],
                                            int base, int k, int lane, Cand<CLEDB_PREC_FP32>* scratch) {
    uint32_t em = 0;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) em |= (v[e] < CUDART_INF_F ? 1u : 0u) << e;
    const int cnt = __popc(em);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int H = __shfl_sync(0xffffffffu, incl, 31);
    float cv = CUDART_INF_F;
    int cp = -1;
    if (lane < k) { cv = list[lane].v; cp = list[lane].pos; }
    const int L = __popc(__ballot_sync(0xffffffffu, cp >= 0));          // the filled slots are a prefix of the sorted list
    const int n = L + H;
    if (n > 32) return false;
    int o = incl - cnt;
#pragma unroll
    for (int e = 0; e < kEPL; ++e)
        if ((em >> e) & 1u) {
            scratch[o].v = v[e];
            scratch[o].pos = base + pos_in_tile<ENC>(lane, e);
            ++o;
        }
    __syncwarp();
    if (lane >= L) {
        cv = CUDART_INF_F;
        cp = -1;
        if (lane < n) { cv = scratch[lane - L].v; cp = scratch[lane - L].pos; }
    }
    int rank = 0;
    bool tie = false;
    for (int i = 0; i < n; ++i) {
        const float vi = __shfl_sync(0xffffffffu, cv, i);
        rank += (vi < cv || (vi == cv && i < lane)) ? 1 : 0;
        tie |= vi == cv && i != lane;
    }
    if (__any_sync(0xffffffffu, tie && lane < n)) return false;
    __syncwarp();
    if (lane < n && rank < k) {
        list[rank].v = cv;
        list[rank].pos = cp;
        if (rank == k - 1) set_tau<CLEDB_PREC_FP32>(rec, c0f, cv);
    }
    __syncwarp();
    return true;
}

// slow_pixel for tiles that may hold MANY entries under the threshold (the neighbours of a pixel's seed tile in the pruned
// search): one rank_select when it applies; otherwise candidates go in best first, so the threshold drops to the tile's k-th
// best after at most k insertions and the remaining hits fall out by comparison.  Same list, same tie rule as slow_pixel.
template <int ENC>
__device__ __forceinline__ void merge_pixel(Rec<CLEDB_PREC_FP32>* rec, float c0f, Cand<CLEDB_PREC_FP32>* list, const float (&d)[kEPL][kNC],
                                            const float (&scale)[kNC], int base, int cnt_end, const int32_t* __restrict__ perm,
                                            int k, int lane, Cand<CLEDB_PREC_FP32>* scratch) {
    float v[kEPL], tau;
    eval_pixel<ENC, CLEDB_PREC_FP32>(rec, c0f, d, scale, v, tau);
#pragma unroll
    for (int e = 0; e < kEPL; ++e)
        if (!(v[e] <= tau) || base + pos_in_tile<ENC>(lane, e) >= cnt_end) v[e] = CUDART_INF_F;      // not a hit (NaN, padding)
    if (rank_select<ENC>(rec, c0f, list, v, base, k, lane, scratch)) return;
    drop_losing_ties<ENC, float>(v, tau, list, base, cnt_end, perm, k, lane);
#pragma unroll
    for (int e = 0; e < kEPL; ++e)
        if (!(v[e] <= tau)) v[e] = CUDART_INF_F;                                                      // the dropped ties are NaN
    for (;;) {
        float lm = v[0];
        int le = 0;
#pragma unroll
        for (int e = 1; e < kEPL; ++e)
            if (v[e] < lm) { lm = v[e]; le = e; }
        const float gm = warp_min(lm);
        if (!(gm < CUDART_INF_F)) break;
        const int src = __ffs(__ballot_sync(0xffffffffu, lm == gm)) - 1;
        const int es = __shfl_sync(0xffffffffu, le, src);
        list_insert<CLEDB_PREC_FP32>(list, rec, c0f, k, gm, base + pos_in_tile<ENC>(src, es), perm, lane);
        tau = rec->tau;
#pragma unroll
        for (int e = 0; e < kEPL; ++e)
            if ((lane == src && e == es) || !(v[e] <= tau)) v[e] = CUDART_INF_F;
    }
}

// a pixel's first tile in the pruned search (empty list): a threshold T with at least k entries under it - the k-th smallest
// of the 32 lane minima - then one rank_select over the entries <= T.  Returns false if that does not apply (fewer than k
// lanes hold a finite value, too many candidates, exact ties): the caller runs bootstrap_pixel.
template <int ENC>
__device__ __forceinline__ bool seed_pixel(Rec<CLEDB_PREC_FP32>* rec, float c0f, Cand<CLEDB_PREC_FP32>* list, const float (&d)[kEPL][kNC],
                                           const float (&scale)[kNC], int base, int cnt_end, int k, int lane,
                                           Cand<CLEDB_PREC_FP32>* scratch) {
    float v[kEPL], tau_unused;
    eval_pixel<ENC, CLEDB_PREC_FP32>(rec, c0f, d, scale, v, tau_unused);
    float lm = CUDART_INF_F;
#pragma unroll
    for (int e = 0; e < kEPL; ++e) {
        if (!(v[e] < CUDART_INF_F) || base + pos_in_tile<ENC>(lane, e) >= cnt_end) v[e] = CUDART_INF_F;   // NaN, inf, padding
        lm = fminf(lm, v[e]);
    }
    float T = CUDART_INF_F;
    for (int r = 0; r < k; ++r) {
        T = warp_min(lm);
        if (lm == T) lm = CUDART_INF_F;
    }
    if (!(T < CUDART_INF_F)) return false;
#pragma unroll
    for (int e = 0; e < kEPL; ++e)
        if (!(v[e] <= T)) v[e] = CUDART_INF_F;
    return rank_select<ENC>(rec, c0f, list, v, base, k, lane, scratch);
}

// first tile of an item: exact top-k of its 256 entries by k rounds of warp-min selection
template <int ENC, int PREC>
__device__ __forceinline__ void bootstrap_pixel(Rec<PREC>* rec, float c0f, Cand<PREC>* list, const float (&d)[kEPL][kNC],
                                                const float (&scale)[kNC], int base, int cnt_end, const int (&rows)[kEPL],
                                                int k, int lane) {
    using V = typename Val<PREC>::type;
    V v[kEPL], tau_unused;
    eval_pixel<ENC, PREC>(rec, c0f, d, scale, v, tau_unused);
    int ps[kEPL];
#pragma unroll
    for (int e = 0; e < kEPL; ++e) {
        const int pos = base + pos_in_tile<ENC>(lane, e);
        const bool ok = !isnan(v[e]) && pos < cnt_end;    // padding decodes to NaN
        ps[e] = ok ? rows[e] : 0x7fffffff;                // ties go to the lowest original row
        if (!ok) v[e] = inf_of(V());
    }
    uint32_t used = 0;
    V keep_v = inf_of(V());
    int keep_p = -1;
    for (int r = 0; r < k; ++r) {
        V lm = inf_of(V());
        int lp = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < kEPL; ++e) {
            const bool free_e = !((used >> e) & 1u);
            if (free_e && (v[e] < lm || (v[e] == lm && ps[e] < lp))) { lm = v[e]; lp = ps[e]; }
        }
        const V gm = warp_min(lm);
        const int gp = (int)__reduce_min_sync(0xffffffffu, (uint32_t)((lm == gm) ? lp : 0x7fffffff));
        if (gp != 0x7fffffff) {
#pragma unroll
            for (int e = 0; e < kEPL; ++e)
                if (ps[e] == gp) used |= 1u << e;
        }
        int won_pos = -1;                                   // stored position of the winner (the owner lane knows it)
#pragma unroll
        for (int e = 0; e < kEPL; ++e)
            if (ps[e] == gp && gp != 0x7fffffff) won_pos = base + pos_in_tile<ENC>(lane, e);
        won_pos = (int)__reduce_max_sync(0xffffffffu, won_pos);
        if (lane == r) { keep_v = gm; keep_p = won_pos; }
    }
    if (lane < k) {
        list[lane].v = keep_v;
        list[lane].pos = keep_p;
    }
    if (lane == k - 1) set_tau<PREC>(rec, c0f, keep_v);
    __syncwarp();
}

template <int ENC, int PREC, bool SEEDED = false>
__global__ void __launch_bounds__(KCfg<ENC, PREC>::kWpc * 32, (PREC == CLEDB_PREC_FP32 && KCfg<ENC, PREC>::kTPS == 1) ? 2 : 1)
k_match(Plan pl, const DbSlot* __restrict__ slots) {
    using C = KCfg<ENC, PREC>;
    using V = typename Val<PREC>::type;
    using CandT = Cand<PREC>;
    using RecT = Rec<PREC>;
    constexpr int TB = C::kTB, TPS = C::kTPS, SB = C::kStageBytes;

    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* stages = smem + (size_t)warp * C::kWarpBytes;                        // [kStages][TB]
    RecT* recs = reinterpret_cast<RecT*>(stages + kStages * SB);                        // [kMaxP + 1]
    CandT* lists = reinterpret_cast<CandT*>(reinterpret_cast<unsigned char*>(recs) + C::kRecBytes);   // [npx][k]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(lists) + C::kListBytes);   // [kStages]
    float* c0s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + (kStages * 8 + 15) / 16 * 16);   // [kMaxP]
    int32_t* seeds = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(c0s) + C::kC0Bytes);                // [kMaxP] (seeded)
    Cand<CLEDB_PREC_FP32>* seed_scratch = reinterpret_cast<Cand<CLEDB_PREC_FP32>*>(reinterpret_cast<unsigned char*>(seeds) + (kMaxP * 4 + 15) / 16 * 16);
    const int k = pl.k;
    uint32_t phase_bits = 0;
    // Seeded brute force: the database is in the k-d tile layout (pl.seed != NULL) and every pixel starts from the exact
    // top-k of ITS OWN seed tile - the tile whose representative entry fits it best (k_seed) - so its threshold is close to
    // final before the scan begins, and the scan over ALL tiles flags only the few tiles that really hold contenders
    // (about 10 per pixel instead of the k ln(N/256) = 54 of a random-order scan).  Every (pixel, entry) pair is still evaluated.
    constexpr bool seeded = SEEDED && PREC == CLEDB_PREC_FP32;      // a separate instantiation: the plain kernel carries none of it

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();

    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&pl.counters[0], 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= pl.counters[2]) break;
        const Item im = make_item(pl, slots, it);
        const DbSlot& db = slots[im.slot];
        const unsigned char* gt = reinterpret_cast<const unsigned char*>(db.tiles) + (size_t)im.tile0 * TB;
        CandT* gout = reinterpret_cast<CandT*>(pl.plist) + (size_t)it * pl.P * k;

        // ---- stage the item: first steps (TPS tiles each) in flight, pixel records, empty lists
        const int nsteps = (im.ntiles + TPS - 1) / TPS;
        if (lane == 0 && !seeded) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (s < nsteps) {
                    const uint32_t bytes = (uint32_t)min(TPS, im.ntiles - s * TPS) * TB;
                    mbar_expect_tx(&bars[s], bytes);
                    tma_load_1d(stages + s * SB, gt + (size_t)s * SB, bytes, &bars[s]);
                }
            }
        }
        for (int i = lane; i < im.npx * (int)(sizeof(RecT) / 16); i += 32) {
            const int j = i / (int)(sizeof(RecT) / 16), w = i % (int)(sizeof(RecT) / 16);
            const int p = pl.order[im.pix0 + j];
            reinterpret_cast<uint4*>(recs + j)[w] = reinterpret_cast<const uint4*>(reinterpret_cast<const RecT*>(pl.recs) + p)[w];
        }
        for (int i = lane; i < im.npx * k; i += 32) {
            lists[i].v = inf_of(V());
            lists[i].pos = -1;
        }
        __syncwarp();
        if (PREC == CLEDB_PREC_FP32) {          // c0 moves to the side array; the record slot holds the flagging start value
            for (int j = lane; j < im.npx; j += 32) {
                Rec<CLEDB_PREC_FP32>* r = reinterpret_cast<Rec<CLEDB_PREC_FP32>*>(recs) + j;
                c0s[j] = r->cm;
                r->cm = -CUDART_INF_F;
            }
            __syncwarp();
        }

        float scale[kNC];
#pragma unroll
        for (int c = 0; c < kNC; ++c) scale[c] = db.scale[chi_comp(c)];
        const uint32_t rec_base = smem_u32(recs);

        float d[kEPL][kNC];                       // tile A of the step
        float d2[TPS == 2 ? kEPL : 1][kNC];       // tile B of the step (fp32 kernel)
        bool exact_only = false;
        if constexpr (PREC == CLEDB_PREC_FP32) {
            if (seeded && im.ntiles > 0) {
                // ---- seed pass: the distinct seed tiles of the item's pixels, one after the other; each pixel takes the exact
                //      top-k of its own.  Seed tiles are counted relative to the item's first tile; a pixel without seed, or
                //      whose seed lies outside this chunk of the class, starts from the chunk's first tile.
                Cand<CLEDB_PREC_FP32>* scratch = seed_scratch;
                const int part_tile0 = im.entry0 / kTile;
                int part = 0;
                for (int q = 0; q < kParts; ++q) if (db.part_tile0[q] == part_tile0 && db.part_cnt[q] > 0) part = q;
                for (int j = lane; j < im.npx; j += 32) {
                    int sd = pl.seed[(size_t)pl.order[im.pix0 + j] * kParts + part];
                    sd = sd < 0 ? 0 : sd - (im.tile0 - part_tile0);
                    seeds[j] = (sd < 0 || sd >= im.ntiles) ? 0 : sd;
                }
                __syncwarp();
                for (int j0 = 0; j0 < im.npx; j0 += 32) {
                    const int mine = j0 + lane < im.npx ? seeds[j0 + lane] : -1;
                    uint32_t active = __ballot_sync(0xffffffffu, mine >= 0);
                    while (active) {
                        const int ts = __shfl_sync(0xffffffffu, mine, __ffs(active) - 1);
                        uint32_t own = __ballot_sync(0xffffffffu, mine == ts);
                        active &= ~own;
                        if (lane == 0) {
                            mbar_expect_tx(&bars[0], TB);
                            tma_load_1d(stages, gt + (size_t)ts * TB, TB, &bars[0]);
                        }
                        mbar_wait(&bars[0], phase_bits & 1u);
                        phase_bits ^= 1u;
                        load_tile_regs<ENC>(stages, lane, d);
                        __syncwarp();
                        const int base = (im.tile0 + ts) * kTile;
                        bool have_rows = false;
                        int rows[kEPL];
                        while (own) {
                            const int j = j0 + __ffs(own) - 1;
                            own &= own - 1;
                            RecT* rj = recs + j;
                            if (!seed_pixel<ENC>(rj, c0s[j], lists + j * k, d, scale, base, im.cnt_end, k, lane, scratch)) {
                                if (!have_rows) {             // original rows of my 8 entries: the tie-break of the exact selection
#pragma unroll
                                    for (int e = 0; e < kEPL; ++e) {
                                        const int pos = base + pos_in_tile<ENC>(lane, e);
                                        rows[e] = pos < im.cnt_end ? db.perm[pos] : 0x7fffffff;
                                    }
                                    have_rows = true;
                                }
                                bootstrap_pixel<ENC, PREC>(rj, c0s[j], lists + j * k, d, scale, base, im.cnt_end, rows, k, lane);
                            }
                        }
                    }
                }
                __syncwarp();
                for (int j = lane; j < im.npx; j += 32) exact_only |= !(recs[j].tau < CUDART_INF_F);
                exact_only = __any_sync(0xffffffffu, exact_only);
                // the scan starts: first steps in flight
                if (lane == 0) {
#pragma unroll
                    for (int s = 0; s < kStages; ++s) {
                        if (s < nsteps) {
                            const uint32_t bytes = (uint32_t)min(TPS, im.ntiles - s * TPS) * TB;
                            mbar_expect_tx(&bars[s], bytes);
                            tma_load_1d(stages + s * SB, gt + (size_t)s * SB, bytes, &bars[s]);
                        }
                    }
                }
            }
        }
        for (int step = 0; step < nsteps; ++step) {
            const int s = step % kStages;
            const int t = step * TPS, nt = min(TPS, im.ntiles - t);
            mbar_wait(&bars[s], (phase_bits >> s) & 1u);
            phase_bits ^= 1u << s;
            load_tile_regs<ENC>(stages + s * SB, lane, d);
            if (TPS == 2) {
                if (nt == 2) load_tile_regs<ENC>(stages + s * SB + TB, lane, reinterpret_cast<float(&)[kEPL][kNC]>(d2));
                else {                            // odd tile count: the missing tile is NaN, never flagged, never a hit
#pragma unroll
                    for (int e = 0; e < kEPL; ++e)
#pragma unroll
                        for (int c = 0; c < kNC; ++c) d2[TPS == 2 ? e : 0][c] = CUDART_NAN_F;
                }
            }
            __syncwarp();                         // every lane has its registers: the stage can be refilled
            if (lane == 0 && step + kStages < nsteps) {
                const uint32_t bytes = (uint32_t)min(TPS, im.ntiles - (step + kStages) * TPS) * TB;
                mbar_expect_tx(&bars[s], bytes);
                tma_load_1d(stages + s * SB, gt + (size_t)(step + kStages) * SB, bytes, &bars[s]);
            }
            const int base = (im.tile0 + t) * kTile;
            if (step == 0 && !seeded) {
                int rows[kEPL];                  // original rows of my 8 entries of the first tile (tie-break only)
#pragma unroll
                for (int e = 0; e < kEPL; ++e) {
                    const int pos = base + pos_in_tile<ENC>(lane, e);
                    rows[e] = pos < im.cnt_end ? db.perm[pos] : 0x7fffffff;
                }
                bool unfilled = false;            // a list that the first tile could not fill keeps tau = inf
                for (int j = 0; j < im.npx; ++j) {
                    bootstrap_pixel<ENC, PREC>(recs + j, PREC == CLEDB_PREC_FP32 ? c0s[j] : 0.f, lists + j * k, d, scale, base,
                                               im.cnt_end, rows, k, lane);
                    if (TPS == 2 && nt == 2)
                        slow_pixel<ENC, PREC>(recs + j, c0s[j], lists + j * k, reinterpret_cast<float(&)[kEPL][kNC]>(d2), scale,
                                              base + kTile, im.cnt_end, db.perm, k, lane);
                    unfilled |= !(recs[j].tau < inf_of(V()));
                }
                exact_only = unfilled;
                continue;
            }
            if (PREC == CLEDB_PREC_FP32 && pl.dbg != 3 && !exact_only) {
                // fast pass: 40 packed FMAs per pixel and tile, no control flow, no min tree.  The accumulators start at
                // c0 - tau*(1 + 2^-19), so the OR of the sign bits says "one of my 8 entries of this tile may beat the
                // threshold of pixel j"; one funnel shift moves that bit into bit (npx-1-j) of the tile's mask.
                // Items can hold more than 32 pixels: the pixel loop runs in segments of <= 32 (one mask word per tile each).
                for (int j0 = 0; j0 < im.npx; j0 += 32) {
                    const int nseg = min(32, im.npx - j0);
                    uint32_t hmask = 0, hmask2 = 0;
                    uint32_t ra = rec_base + (uint32_t)j0 * (uint32_t)sizeof(RecT);
                    float4 n0 = lds128(ra), n1 = lds128(ra + 16), n2 = lds128(ra + 32);
#pragma unroll 2
                    for (int j = 0; j < nseg; ++j) {
                        const float4 x0 = n0, x1 = n1, x2 = n2;
                        ra += (uint32_t)sizeof(RecT);
                        n0 = lds128(ra); n1 = lds128(ra + 16); n2 = lds128(ra + 32);     // prefetch (recs has a spare slot)
                        const float a[kNC] = {x0.x, x0.y, x0.z, x0.w, x1.x};
                        const float nb[kNC] = {x1.y, x1.z, x1.w, x2.x, x2.y};
                        float v[kEPL];
                        eval_fp32_all(d, a, nb, x2.z, v);
                        const uint32_t sg = ((__float_as_uint(v[0]) | __float_as_uint(v[1]) | __float_as_uint(v[2])) |
                                             (__float_as_uint(v[3]) | __float_as_uint(v[4]) | __float_as_uint(v[5]))) |
                                            (__float_as_uint(v[6]) | __float_as_uint(v[7]));
                        hmask = __funnelshift_l(sg, hmask, 1);
                        if (TPS == 2) {
                            float w[kEPL];
                            eval_fp32_all(reinterpret_cast<float(&)[kEPL][kNC]>(d2), a, nb, x2.z, w);
                            const uint32_t sg2 = ((__float_as_uint(w[0]) | __float_as_uint(w[1]) | __float_as_uint(w[2])) |
                                                  (__float_as_uint(w[3]) | __float_as_uint(w[4]) | __float_as_uint(w[5]))) |
                                                 (__float_as_uint(w[6]) | __float_as_uint(w[7]));
                            hmask2 = __funnelshift_l(sg2, hmask2, 1);
                        }
                    }
                    // rare pass: pixels flagged by any lane are evaluated again and their candidates inserted at once, so
                    // the threshold is exact before the pixel is visited for the next step (tile A before tile B)
                    uint32_t any = pl.dbg == 2 ? 0u : __reduce_or_sync(0xffffffffu, hmask);
                    while (any) {
                        const int b = 31 - __clz(any);
                        any &= ~(1u << b);
                        const int j = j0 + nseg - 1 - b;
                        if constexpr (seeded) {
                            if (seeds[j] == t) continue;                       // its own seed tile: those entries are in the list already
                            // a neighbour of the seed tile can hold dozens of entries under the threshold: best first
                            merge_pixel<ENC>(recs + j, c0s[j], lists + j * k, d, scale, base, im.cnt_end, db.perm, k, lane, seed_scratch);
                        } else
                        slow_pixel<ENC, PREC>(recs + j, c0s[j], lists + j * k, d, scale, base, im.cnt_end, db.perm, k, lane);
                    }
                    if (TPS == 2) {
                        any = pl.dbg == 2 ? 0u : __reduce_or_sync(0xffffffffu, hmask2);
                        while (any) {
                            const int b = 31 - __clz(any);
                            any &= ~(1u << b);
                            const int j = j0 + nseg - 1 - b;
                            if constexpr (seeded) {
                                if (seeds[j] == t + 1) continue;
                                merge_pixel<ENC>(recs + j, c0s[j], lists + j * k, reinterpret_cast<float(&)[kEPL][kNC]>(d2), scale,
                                                 base + kTile, im.cnt_end, db.perm, k, lane, seed_scratch);
                            } else
                            slow_pixel<ENC, PREC>(recs + j, c0s[j], lists + j * k, reinterpret_cast<float(&)[kEPL][kNC]>(d2), scale,
                                                  base + kTile, im.cnt_end, db.perm, k, lane);
                        }
                    }
                }
            } else {
                for (int j = 0; j < im.npx; ++j) {
                    if (!(seeded && seeds[j] == t))
                        slow_pixel<ENC, PREC>(recs + j, PREC == CLEDB_PREC_FP32 ? c0s[j] : 0.f, lists + j * k, d, scale, base,
                                              im.cnt_end, db.perm, k, lane);
                    if (TPS == 2 && nt == 2 && !(seeded && seeds[j] == t + 1))
                        slow_pixel<ENC, PREC>(recs + j, c0s[j], lists + j * k, reinterpret_cast<float(&)[kEPL][kNC]>(d2), scale,
                                              base + kTile, im.cnt_end, db.perm, k, lane);
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < im.npx * k; i += 32) gout[i] = lists[i];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------
// kernels 4b: the search with exact tile pruning (the default: databases put while cledb_set_pruning is on)
// ------------------------------------------------------------------------------------------------------
// The tiles of such a database are the leaves of a k-d partition (db.cu), each with the bounds lo_c <= d_c <= hi_c of its
// entries and one representative entry; groups of 16 consecutive tiles carry their joint bounds.  With r_c(d) =
// fma(d, a_c, nb_c) monotone in d, |r_c| >= m_c = distance of the interval [r_c(lo), r_c(hi)] from 0, and the accumulation
// chain acc = fma(r, r, acc) is monotone in |r| and in acc: the bound computed with the SAME chain is <= the computed chi^2 of
// every entry of the tile, exactly, in floating point.  A tile whose bound exceeds a pixel's threshold cannot change that
// pixel's list, so the result is bit-identical to k_match (tests/test_bound_property.py, tests/test_gpu_parity.py).
//
//   k_seed        four threads per pixel, pixels taken in bucket order (one database per block, its representatives staged in
//                 shared memory): the tile whose representative entry fits the pixel best - a coarse search over one entry
//                 per tile.  k_fscan + k_scatter_fine then order the pixels by (bucket, seed tile), so that the pixels of an
//                 item are neighbours in the space of the chi^2 components and need nearly the same tiles.
//   k_match_pruned  one warp per item of <= 32 pixels: (1) every pixel takes the exact top-k of its OWN seed tile: a
//                 threshold close to the final one; (2) the lanes bound the tile groups, then the tiles of surviving groups,
//                 for their pixel (32/P lanes share a pixel): a bit mask per pixel, whose union is the item's tile list;
//                 (3) a listed tile is loaded (TMA) and runs through the fast pass of k_match for the pixels that need it,
//                 flagged pairs are merged exactly.  Tiles nobody needs are not even loaded.
__device__ __forceinline__ float tile_lower_bound(const float4 b0, const float4 b1, const float4 b2, const float (&a)[kNC],
                                                  const float (&nb)[kNC], float c0) {
    const float lo[kNC] = {b0.x, b0.y, b0.z, b0.w, b1.x};
    const float hi[kNC] = {b1.y, b1.z, b1.w, b2.x, b2.y};
    float lb = c0;
#pragma unroll
    for (int c = 0; c < kNC; ++c) {
        const float rl = fmaf(lo[c], a[c], nb[c]), rh = fmaf(hi[c], a[c], nb[c]);
        float mn = fmaxf(fmaxf(fminf(rl, rh), -fmaxf(rl, rh)), 0.f);      // both > 0: the smaller; both < 0: the smaller |.|; else 0
        const float t = rl + rh;
        mn = (t == t) ? mn : 0.f;                                         // NaN anywhere: no bound from this component
        lb = fmaf(mn, mn, lb);
    }
    return lb;
}

// Fine sort key of a pixel inside its database slot: the buckets of k_scan (slot, sign class) must stay contiguous and in
// order, so every class owns the keys [tile0 + q, tile0 + q + ntiles] - one per tile plus one for pixels without seed.
__device__ __forceinline__ int fine_key(const Plan& pl, const DbSlot& db, int key, const int32_t* seed_of_pixel) {
    int q = key % kParts;
    if (pl.mode != CLEDB_MODE_IQUV) {
        q = 0;
        for (int i = kParts - 1; i >= 0; --i) if (db.part_cnt[i] > 0) q = i;      // the first class that has entries
        if (pl.fused) q = largest_class(db);                                      // the class a fused item searches first
    }
    int t = seed_of_pixel[q];
    if (t < 0 || t >= db.part_ntiles[q]) t = db.part_ntiles[q];
    return (key / kParts) * pl.ftiles + db.part_tile0[q] + q + t;
}

constexpr int kSeedCap = 1528;          // representative entries a block of k_seed stages in shared memory (8 floats each: 48 KB)
constexpr int kSeedLanes = 4;           // threads that share a pixel in k_seed (each takes every 4th tile)
constexpr int kSeedRounds = 2;          // pixels per thread quadruple: a block of 256 threads seeds 128 pixels with one staging

__global__ void __launch_bounds__(256) k_seed(Plan pl, const DbSlot* __restrict__ slots) {
    __shared__ __align__(16) float4 s_rep[kSeedCap * 2];      // (r0 r1 r2 r3), (r4 - - -) per tile: two 16-byte reads per evaluation
    __shared__ int s_slot, s_mixed;
    // pixels are taken in bucket order (k_scatter ran before): a block works on one database, mostly on one sign class
    const int sub = threadIdx.x % kSeedLanes;
    const int64_t nsearched = (int64_t)pl.off[pl.nkeys - 1] + pl.cnt[pl.nkeys - 1];
    const int64_t ro0 = (int64_t)blockIdx.x * (blockDim.x / kSeedLanes) * kSeedRounds + threadIdx.x / kSeedLanes;
    // the representative entries of the block's database into shared memory, if the whole block searches one database
    // (the usual case) and they fit; otherwise every thread reads them from global memory
    if (threadIdx.x == 0) { s_slot = -1; s_mixed = 0; }
    __syncthreads();
    int64_t pr[kSeedRounds];
    int keyr[kSeedRounds];
#pragma unroll
    for (int rd = 0; rd < kSeedRounds; ++rd) {
        const int64_t ro = ro0 + (int64_t)rd * (blockDim.x / kSeedLanes);
        pr[rd] = ro < nsearched ? pl.order[ro] : -1;
        keyr[rd] = pr[rd] >= 0 ? pl.key[pr[rd]] : -1;
        if (keyr[rd] >= 0) atomicMax(&s_slot, keyr[rd] / kParts);
    }
    __syncthreads();
    const int bslot = s_slot;
#pragma unroll
    for (int rd = 0; rd < kSeedRounds; ++rd)
        if (keyr[rd] >= 0 && keyr[rd] / kParts != bslot) s_mixed = 1;
    __syncthreads();
    bool staged = false;
    if (!s_mixed && bslot >= 0) {
        const DbSlot& db = slots[bslot];
        const int nt = db.part_tile0[kParts - 1] + db.part_ntiles[kParts - 1];
        if (nt <= kSeedCap) {
            const float4* bx = reinterpret_cast<const float4*>(db.boxes);
            for (int i = threadIdx.x; i < nt; i += blockDim.x) {
                const float4 b2 = __ldg(bx + (size_t)i * 4 + 2), b3 = __ldg(bx + (size_t)i * 4 + 3);
                s_rep[2 * i] = make_float4(b2.z, b2.w, b3.x, b3.y);
                s_rep[2 * i + 1] = make_float4(b3.z, 0.f, 0.f, 0.f);
            }
            staged = true;
        }
    }
    __syncthreads();
    const uint32_t quad = 0xfu << ((threadIdx.x & 31) & ~(kSeedLanes - 1));
    for (int rd = 0; rd < kSeedRounds; ++rd) {
        // the kSeedLanes threads of a pixel sit in one warp and skip together (shuffles below)
        const int64_t p = pr[rd];
        const int key = keyr[rd];
        if (key < 0) continue;
        const DbSlot& db = slots[key / kParts];
        const Rec<CLEDB_PREC_FP32> r = reinterpret_cast<const Rec<CLEDB_PREC_FP32>*>(pl.recs)[p];
        int32_t seeds[kParts];
        for (int q = 0; q < kParts; ++q) {
            int best_t = 0x7fffffff;
            float best = CUDART_INF_F;
            const bool mine = pl.mode == CLEDB_MODE_IQUV ? q == key % kParts : (pl.fused ? q == largest_class(db) : db.part_cnt[q] > 0);
            if (mine) {
                const int nt = db.part_ntiles[q], t0 = db.part_tile0[q];
                if (staged) {
#pragma unroll 4
                    for (int t = sub; t < nt; t += kSeedLanes) {
                        const float4 x = s_rep[2 * (t0 + t)], y = s_rep[2 * (t0 + t) + 1];
                        const float rep[kNC] = {x.x, x.y, x.z, x.w, y.x};
                        const float v = eval_fp32(rep, r.a, r.nb, r.cm);
                        if (v < best) { best = v; best_t = t; }
                    }
                } else {
                    const float4* bx = reinterpret_cast<const float4*>(db.boxes) + (size_t)t0 * 4;
#pragma unroll 4
                    for (int t = sub; t < nt; t += kSeedLanes) {
                        const float4 b2 = __ldg(bx + (size_t)t * 4 + 2), b3 = __ldg(bx + (size_t)t * 4 + 3);
                        const float rep[kNC] = {b2.z, b2.w, b3.x, b3.y, b3.z};
                        const float v = eval_fp32(rep, r.a, r.nb, r.cm);
                        if (v < best) { best = v; best_t = t; }
                    }
                }
#pragma unroll
                for (int o = 1; o < kSeedLanes; o <<= 1) {                  // best of the pixel's threads, lowest tile on ties
                    const float ov = __shfl_xor_sync(quad, best, o);
                    const int ot = __shfl_xor_sync(quad, best_t, o);
                    if (ov < best || (ov == best && ot < best_t)) { best = ov; best_t = ot; }
                }
            }
            if (best_t == 0x7fffffff) best_t = -1;
            seeds[q] = best_t;
            if (sub == 0) pl.seed[p * kParts + q] = best_t;
            // the search cost of a pixel grows with its threshold, and the chi^2 of the best representative is a monotone proxy
            // of it; the class a pixel searches first (its own for IQUV, the largest for fused IQUD) gives the weight
            if (sub == 0 && pl.wpix && mine && (pl.mode == CLEDB_MODE_IQUV || q == largest_class(db))) pl.wpix[p] = best_t < 0 ? CUDART_INF_F : best;
        }
        if (sub == 0) atomicAdd(&pl.fcnt[fine_key(pl, db, key, seeds)], 1);
    }
}

// exclusive scan of the fine counts: one block per database slot, starting from the slot's first position in `order`
// (k_scan's offset of the slot's first bucket; the fine keys of a slot are class-major like the buckets)
__global__ void __launch_bounds__(256) k_fscan(Plan pl) {
    __shared__ int s_sum[256];
    const int tid = threadIdx.x, slot = blockIdx.x;
    const int n = pl.ftiles;
    int32_t* cnt = pl.fcnt + (size_t)slot * n;
    int32_t* off = pl.foff + (size_t)slot * n;
    const int per = (n + 255) / 256;
    const int k0 = tid * per, k1 = min(n, k0 + per);
    int sum = 0;
    for (int i = k0; i < k1; ++i) sum += cnt[i];
    s_sum[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int a = 0;
        if (tid >= o) a = s_sum[tid - o];
        __syncthreads();
        s_sum[tid] += a;
        __syncthreads();
    }
    int run = pl.off[slot * kParts] + s_sum[tid] - sum;
    for (int i = k0; i < k1; ++i) { off[i] = run; run += cnt[i]; }
}

// pixels into (bucket, seed tile) order; the buckets of k_scan are unions of consecutive fine buckets
__global__ void __launch_bounds__(256) k_scatter_fine(Plan pl, const DbSlot* __restrict__ slots) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pl.npix) return;
    const int key = pl.key[p];
    if (key < 0) { pl.rank[p] = -1; return; }
    const int f = fine_key(pl, slots[key / kParts], key, pl.seed + p * kParts);
    const int r = pl.foff[f] + atomicAdd(&pl.fcursor[f], 1);
    pl.order[r] = (int32_t)p;
    pl.rank[p] = r;
}

#ifndef CLEDB_PRUNED_WPC
#define CLEDB_PRUNED_WPC 16
#endif
// Longest items first.  The items of the pruned search differ a lot in length - a pixel whose observation fits no entry well
// keeps a loose threshold and has to look at a hundred tiles instead of ten - and the kernel ends when its longest item
// ends: taken in index order, a long item that comes up late leaves the other warps idle (256^2 map: half the warps were
// idle at 60 % of the kernel's duration).  So the queue hands the items out heaviest first: weight of an item = sum of the
// seed chi^2 of its pixels (k_seed), quantised to 64 logarithmic bins, counting sort by bin.
__global__ void __launch_bounds__(256) k_item_weight(Plan pl, const DbSlot* __restrict__ slots) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= pl.counters[2]) return;
    const Item im = make_item(pl, slots, it);
    float w = 0.f;
    for (int j = 0; j < im.npx; ++j) w += pl.wpix[pl.order[im.pix0 + j]];
    int bin = 63;                                                   // inf / NaN: a pixel without seed - the exact path over every tile
    if (w < 1e30f) bin = min(62, max(0, (int)(2.f * log2f(fmaxf(w, 1e-6f))) + 24));
    pl.qbin[it] = bin;
    atomicAdd(&pl.qhist[bin], 1);
}
// Queue layout: bins from heavy to light; an item of a bin at least qsplit bins above the median bin (weight >= 2^(qsplit/2)
// times the median item) is handed out in pieces of qpiece pixels - its pixels are independent, a piece is just a smaller item
// with the same storage - so that the longest work unit of a small map is one heavy pixel, not four of them in a row.  Every
// heavy item reserves P / qpiece entries; pieces beyond the item's pixel count stay holes (-1).
__global__ void __launch_bounds__(256) k_item_order(Plan pl) {
    __shared__ int s_hist[64], s_before[64], s_split;
    const int total = pl.counters[2];
    if (threadIdx.x < 64) s_hist[threadIdx.x] = pl.qhist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        int sb = 64;
        if (pl.qsplit > 0) {
            int acc = 0, med = 0;
            for (int b = 0; b < 64; ++b) { acc += s_hist[b]; if (2 * acc >= total) { med = b; break; } }
            sb = min(64, med + pl.qsplit);
        }
        const int per = (pl.P + pl.qpiece - 1) / pl.qpiece;
        s_split = sb;
        int run = 0;
        for (int b = 63; b >= 0; --b) { s_before[b] = run; run += s_hist[b] * (b >= sb ? per : 1); }      // heavier bins come first
        if (blockIdx.x == 0) { pl.qhist[128] = run; pl.qhist[129] = sb; }
    }
    __syncthreads();
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const int bin = pl.qbin[it];
    if (bin >= s_split) {
        const int per = (pl.P + pl.qpiece - 1) / pl.qpiece;
        const int pos = s_before[bin] + atomicAdd(&pl.qhist[64 + bin], per);
        for (int c = 0; c < per; ++c) pl.qorder[pos + c] = it | ((c + 1) << 24);
    } else {
        pl.qorder[s_before[bin] + atomicAdd(&pl.qhist[64 + bin], 1)] = it;
    }
}

template <int ENC> struct PCfg {        // shared memory of one warp of the pruned search: one tile in flight, <= 32 pixels
    static constexpr int kWpc = CLEDB_PRUNED_WPC;
    static constexpr int kTB = tile_bytes<ENC>();
    // stage buffers of the tile walk (tile s + kNst is requested as soon as tile s sits in registers)
    static constexpr int kNst = 1;       // measured (profiles/r02_tuning.txt): 2 or 3 stages are 5 - 8 % slower - the extra shared memory
                                        // takes the L1 the bound tables live in, and the walk is not bound by the load latency
    static constexpr int kRecBytes = (32 + 1) * (int)sizeof(Rec<CLEDB_PREC_FP32>);
    static constexpr int kListBytes = kListCap * (int)sizeof(Cand<CLEDB_PREC_FP32>);
};
constexpr int kTileListCap = 96;        // tile list of a warp: evaluated at >= 64 entries, one group adds <= 16

template <int ENC> __host__ __device__ constexpr int pruned_warp_bytes() {
    return (PCfg<ENC>::kNst * PCfg<ENC>::kTB + PCfg<ENC>::kRecBytes + PCfg<ENC>::kListBytes + 32 + 128 + (2 * kTileListCap + 32) * 4 + 32 * 8 + 127) / 128 * 128;
}

template <int ENC>
__global__ void __launch_bounds__(PCfg<ENC>::kWpc * 32, 1)
k_match_pruned(Plan pl, const DbSlot* __restrict__ slots, unsigned long long* __restrict__ stats) {
    constexpr int PREC = CLEDB_PREC_FP32;
    using C = PCfg<ENC>;
    using CandT = Cand<PREC>;
    using RecT = Rec<PREC>;
    constexpr int TB = C::kTB;

    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NST = C::kNst;
    unsigned char* stages = smem + (size_t)warp * pruned_warp_bytes<ENC>();                 // NST tiles
    RecT* recs = reinterpret_cast<RecT*>(stages + NST * TB);                                 // [32 + 1]
    CandT* lists = reinterpret_cast<CandT*>(reinterpret_cast<unsigned char*>(recs) + C::kRecBytes);   // [npx][k]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(lists) + C::kListBytes);
    float* c0s = reinterpret_cast<float*>(bars + 4);                                         // [32]
    int32_t* tlist = reinterpret_cast<int32_t*>(c0s + 32);                                   // [kTileListCap] tiles to evaluate
    uint32_t* tmask = reinterpret_cast<uint32_t*>(tlist + kTileListCap);                     // [kTileListCap] the pixels that need each of them
    int32_t* slist = tlist + 2 * kTileListCap;                                               // [32] the item's seed tiles
    CandT* scratch = reinterpret_cast<CandT*>(slist + 32);                                   // [32] candidates of one selection
    const int k = pl.k;
    uint32_t phase = 0;
    unsigned long long n_eval = 0, n_all = 0;

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NST; ++q) mbar_init(&bars[q], 1);
        fence_barrier_init();
        if (pl.dbg == 7 && stats && blockIdx.x == 0 && warp == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            stats[2] = t;
        }
    }
    __syncwarp();

    for (;;) {
        int it = 0;
        if (lane == 0) {
            it = atomicAdd(&pl.counters[0], 1);
            if (pl.qorder) it = it < pl.qhist[128] ? pl.qorder[it] : 0x7fffffff;   // the queue hands the items out heaviest first
            else if (it >= pl.counters[2]) it = 0x7fffffff;
        }
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it == 0x7fffffff) break;
        if (it < 0) continue;                                                 // a hole: piece of a heavy item beyond its pixels
        const int piece = it >> 24;                                           // 0: the whole item, c: its c-th piece of qpiece pixels
        it &= 0xffffff;
        Item im = make_item(pl, slots, it);                                   // items hold <= 32 pixels and whole sign classes
        int piece_j0 = 0;
        if (piece > 0) {
            piece_j0 = (piece - 1) * pl.qpiece;
            if (piece_j0 >= im.npx) continue;
            im.pix0 += piece_j0;
            im.npx = min(pl.qpiece, im.npx - piece_j0);
        }
        const DbSlot& db = slots[im.slot];
        unsigned long long dbg_t0 = 0, dbg_e0 = n_eval;
        int dbg_exact = 0;
        if (pl.dbg == 7) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_t0));
        CandT* gout = reinterpret_cast<CandT*>(pl.plist) + ((size_t)it * pl.P + piece_j0) * k;
        // The sign classes this item searches: its own (IQUV), or - IQUD (pl.fused) - every non-empty class of the database, one
        // after the other with ONE list per pixel: the largest class first (seed tiles, thresholds close to final), so the
        // others start with tight thresholds, skip the seed pass and lose most of their tiles to the bounds at once.
        int cls[kParts], ncls = 0;
        if (pl.fused) {
            const int first = largest_class(db);
            cls[ncls++] = first;
            for (int q = 0; q < kParts; ++q) if (q != first && db.part_cnt[q] > 0) cls[ncls++] = q;
        } else {
            int part = 0;
            for (int q = 0; q < kParts; ++q) if (db.part_tile0[q] == im.tile0 && db.part_ntiles[q] == im.ntiles && db.part_cnt[q] > 0) part = q;
            cls[ncls++] = part;
        }

        for (int i = lane; i < im.npx * (int)(sizeof(RecT) / 16); i += 32) {
            const int j = i / (int)(sizeof(RecT) / 16), w = i % (int)(sizeof(RecT) / 16);
            const int p = pl.order[im.pix0 + j];
            reinterpret_cast<uint4*>(recs + j)[w] = reinterpret_cast<const uint4*>(reinterpret_cast<const RecT*>(pl.recs) + p)[w];
        }
        for (int i = lane; i < im.npx * k; i += 32) {
            lists[i].v = CUDART_INF_F;
            lists[i].pos = -1;
        }
        __syncwarp();
        // Lane L owns pixel L for the seeds.  For the bounds the warp is cut into 32/Pb groups of Pb lanes (Pb = 4, 8, 16 or 32,
        // the smallest that holds the item): lane L bounds pixel L % Pb and takes every (32/Pb)-th tile of a tile group.
        const int Pb = im.npx <= 4 ? 4 : (im.npx <= 8 ? 8 : (im.npx <= 16 ? 16 : 32)), lanes_per_pixel = 32 / Pb;
        const int bj = lane % Pb, bg = lane / Pb;
        float pa[kNC], pnb[kNC], pc0 = 0.f;
        if (bj < im.npx) {
            const float4* rp = reinterpret_cast<const float4*>(recs + bj);
            const float4 x0 = rp[0], x1 = rp[1], x2 = rp[2];
            pa[0] = x0.x; pa[1] = x0.y; pa[2] = x0.z; pa[3] = x0.w; pa[4] = x1.x;
            pnb[0] = x1.y; pnb[1] = x1.z; pnb[2] = x1.w; pnb[3] = x2.x; pnb[4] = x2.y;
            pc0 = x2.z;
        } else {
#pragma unroll
            for (int c = 0; c < kNC; ++c) { pa[c] = 0.f; pnb[c] = 0.f; }
        }
        __syncwarp();
        if (lane < im.npx) {                      // c0 to the side array, the flagging start value into its slot
            c0s[lane] = pc0;
            recs[lane].cm = -CUDART_INF_F;
        }
        __syncwarp();
        float scale[kNC];
#pragma unroll
        for (int c = 0; c < kNC; ++c) scale[c] = db.scale[chi_comp(c)];
        const uint32_t rec_base = smem_u32(recs);
        float d[kEPL][kNC];
        for (int ci = 0; ci < ncls; ++ci) {
        const int part = cls[ci];
        const int c_tile0 = db.part_tile0[part], c_ntiles = db.part_ntiles[part], c_cnt_end = c_tile0 * kTile + db.part_cnt[part];
        if (c_ntiles == 0) continue;
        const unsigned char* gt = reinterpret_cast<const unsigned char*>(db.tiles) + (size_t)c_tile0 * TB;
        const float4* boxes = reinterpret_cast<const float4*>(db.boxes) + (size_t)c_tile0 * 4;
        const float4* sboxes = reinterpret_cast<const float4*>(db.sboxes) + (size_t)db.part_sb0[part] * 3;
        n_all += (unsigned long long)im.npx * c_ntiles;
        int my_seed = -1;                         // seed tiles only for the first class of the item (the lists are empty then)
        if (ci == 0 && lane < im.npx) {
            my_seed = pl.seed[(size_t)pl.order[im.pix0 + lane] * kParts + part];
            if (my_seed >= c_ntiles) my_seed = -1;
        }
        bool exact_only = false;
        if (ci == 0) {
        // ---- the distinct seed tiles of the item's pixels
        int ns = 0;
        {
            uint32_t active = __ballot_sync(0xffffffffu, my_seed >= 0);
            while (active) {
                const int leader = __ffs(active) - 1;
                const int ts = __shfl_sync(0xffffffffu, my_seed, leader);
                active &= ~__ballot_sync(0xffffffffu, my_seed == ts);
                if (lane == 0) slist[ns] = ts;
                ++ns;
            }
            if (ns == 0) { if (lane == 0) slist[0] = 0; ns = 1; }
            __syncwarp();
        }

        // ---- every pixel starts with the exact top-k of its own seed tile: a threshold close to the final one
        {
            if (lane == 0) {
                mbar_expect_tx(&bars[0], TB);
                tma_load_1d(stages, gt + (size_t)slist[0] * TB, TB, &bars[0]);
            }
            for (int i = 0; i < ns; ++i) {
                const int ts = slist[i];
                mbar_wait(&bars[0], phase & 1u);
                phase ^= 1u;
                load_tile_regs<ENC>(stages, lane, d);
                __syncwarp();
                if (lane == 0 && i + 1 < ns) {
                    mbar_expect_tx(&bars[0], TB);
                    tma_load_1d(stages, gt + (size_t)slist[i + 1] * TB, TB, &bars[0]);
                }
                const int base = (c_tile0 + ts) * kTile;
                uint32_t own = __ballot_sync(0xffffffffu, my_seed == ts);
                n_eval += __popc(own);
                while (own) {
                    const int j = __ffs(own) - 1;
                    own &= own - 1;
                    if (!seed_pixel<ENC>(recs + j, c0s[j], lists + j * k, d, scale, base, c_cnt_end, k, lane, scratch)) {
                        int rows[kEPL];                   // original rows of my 8 entries: the tie-break of the exact selection
#pragma unroll
                        for (int e = 0; e < kEPL; ++e) {
                            const int pos = base + pos_in_tile<ENC>(lane, e);
                            rows[e] = pos < c_cnt_end ? db.perm[pos] : 0x7fffffff;
                        }
                        bootstrap_pixel<ENC, PREC>(recs + j, c0s[j], lists + j * k, d, scale, base, c_cnt_end, rows, k, lane);
                    }
                }
            }
            __syncwarp();
        }
        }
        // a list its seed tile could not fill (or a pixel without seed) keeps tau = inf: exact path for the whole item
        exact_only = __any_sync(0xffffffffu, lane < im.npx && !(recs[lane].tau < CUDART_INF_F));
        dbg_exact |= exact_only ? 1 : 0;
        // one tile against the pixels in `set` (bit j = pixel j): the fast pass of k_match over those pixels, then the exact
        // re-evaluation of the flagged ones
        auto tile_pixels = [&](uint32_t set, const float (&dd)[kEPL][kNC], int base) {
            n_eval += __popc(set);
            if (!exact_only) {
                uint32_t hm = 0, rem = set;
                int j = __ffs(rem) - 1;
                rem &= rem - 1;
                uint32_t ra = rec_base + (uint32_t)max(j, 0) * (uint32_t)sizeof(RecT);
                float4 n0 = lds128(ra), n1 = lds128(ra + 16), n2 = lds128(ra + 32);
                while (j >= 0) {
                    const float4 x0 = n0, x1 = n1, x2 = n2;
                    const int jn = __ffs(rem) - 1;                                  // -1 when none is left
                    rem &= rem - 1;
                    ra = rec_base + (uint32_t)max(jn, 0) * (uint32_t)sizeof(RecT);
                    n0 = lds128(ra); n1 = lds128(ra + 16); n2 = lds128(ra + 32);    // next record in flight
                    const float a[kNC] = {x0.x, x0.y, x0.z, x0.w, x1.x};
                    const float nb[kNC] = {x1.y, x1.z, x1.w, x2.x, x2.y};
                    float v[kEPL];
                    eval_fp32_all(dd, a, nb, x2.z, v);
                    const uint32_t sg = ((__float_as_uint(v[0]) | __float_as_uint(v[1]) | __float_as_uint(v[2])) |
                                         (__float_as_uint(v[3]) | __float_as_uint(v[4]) | __float_as_uint(v[5]))) |
                                        (__float_as_uint(v[6]) | __float_as_uint(v[7]));
                    hm |= (sg >> 31) << j;
                    j = jn;
                }
                uint32_t any = __reduce_or_sync(0xffffffffu, hm);
                while (any) {
                    const int jj = __ffs(any) - 1;
                    any &= any - 1;
                    merge_pixel<ENC>(recs + jj, c0s[jj], lists + jj * k, dd, scale, base, c_cnt_end, db.perm, k, lane, scratch);
                }
            } else {
                while (set) {
                    const int jj = __ffs(set) - 1;
                    set &= set - 1;
                    slow_pixel<ENC, PREC, false>(recs + jj, c0s[jj], lists + jj * k, dd, scale, base, c_cnt_end, db.perm, k, lane);
                }
            }
        };
        // the listed tiles one after the other, the next NST in flight
        auto run_tiles = [&](int n) {
            auto issue = [&](int s) {
                if (lane == 0) {
                    const int q = s % NST;
                    mbar_expect_tx(&bars[q], TB);
                    tma_load_1d(stages + q * TB, gt + (size_t)tlist[s] * TB, TB, &bars[q]);
                }
            };
            for (int s = 0; s < min(n, NST); ++s) issue(s);
            for (int s = 0; s < n; ++s) {
                const int ta = tlist[s];
                const uint32_t set_a = tmask[s];
                const int q = s % NST;
                mbar_wait(&bars[q], (phase >> q) & 1u);
                phase ^= 1u << q;
                load_tile_regs<ENC>(stages + q * TB, lane, d);
                __syncwarp();
                if (s + NST < n) issue(s + NST);
                tile_pixels(set_a, d, (c_tile0 + ta) * kTile);
            }
            __syncwarp();
        };

        // ---- bounds against the thresholds as they are by then (they only tighten: a superset of the pairs that matter).
        //      Two levels: groups of 16 tiles first, 64 groups at a time and the lanes of a pixel taking them in turns; then
        //      the tiles of the groups some pixel keeps, again in turns.  Listed tiles are evaluated whenever 64 have piled
        //      up, which tightens the thresholds for the groups still to come.
        const uint32_t low = Pb == 32 ? 0xffffffffu : ((1u << Pb) - 1u);
        const int ngroups = (c_ntiles + kSuper - 1) / kSuper;
        int nu = 0, g0 = -64;
        uint32_t gm0 = 0u, gm1 = 0u, gu0 = 0u, gu1 = 0u;             // surviving groups of the 64: my pixel's, any pixel's
        float tau = bj < im.npx ? recs[bj].tau : -CUDART_INF_F;      // a lane without pixel needs nothing
        for (;;) {
            while (nu < 64) {
                if ((gu0 | gu1) == 0u) {                             // next 64 groups
                    g0 += 64;
                    if (g0 >= ngroups) break;
                    const int ng = min(64, ngroups - g0);
                    const float4* sb = sboxes + (size_t)g0 * 3;
                    gm0 = gm1 = 0u;
                    for (int gi = bg; gi < ng; gi += lanes_per_pixel) {
                        const float lbg = tile_lower_bound(__ldg(sb + gi * 3), __ldg(sb + gi * 3 + 1), __ldg(sb + gi * 3 + 2), pa, pnb, pc0);
                        if (!(lbg > tau)) {                                           // a NaN bound keeps the group
                            if (gi < 32) gm0 |= 1u << gi; else gm1 |= 1u << (gi - 32);
                        }
                    }
                    if (Pb <= 16) { gm0 |= __shfl_xor_sync(0xffffffffu, gm0, 16); gm1 |= __shfl_xor_sync(0xffffffffu, gm1, 16); }
                    if (Pb <= 8) { gm0 |= __shfl_xor_sync(0xffffffffu, gm0, 8); gm1 |= __shfl_xor_sync(0xffffffffu, gm1, 8); }
                    if (Pb <= 4) { gm0 |= __shfl_xor_sync(0xffffffffu, gm0, 4); gm1 |= __shfl_xor_sync(0xffffffffu, gm1, 4); }
                    gu0 = __reduce_or_sync(0xffffffffu, gm0);
                    gu1 = __reduce_or_sync(0xffffffffu, gm1);
                    continue;
                }
                const bool hi = gu0 == 0u;
                uint32_t cur = hi ? gu1 : gu0;
                const int gb = __ffs(cur) - 1;
                cur &= cur - 1;
                if (hi) gu1 = cur; else gu0 = cur;
                const int t0 = (g0 + (hi ? 32 : 0) + gb) * kSuper, nt = min(kSuper, c_ntiles - t0);
                uint32_t mg = 0;
                if (((hi ? gm1 : gm0) >> gb) & 1u) {
                    const float4* bx = boxes + (size_t)t0 * 4;
                    for (int i = bg; i < nt; i += lanes_per_pixel) {
                        const float lb = tile_lower_bound(__ldg(bx + i * 4), __ldg(bx + i * 4 + 1), __ldg(bx + i * 4 + 2), pa, pnb, pc0);
                        mg |= (!(lb > tau) ? 1u : 0u) << i;                          // a NaN bound keeps the tile
                    }
                }
                if (Pb <= 16) mg |= __shfl_xor_sync(0xffffffffu, mg, 16);          // the lanes of a pixel put their shares together
                if (Pb <= 8) mg |= __shfl_xor_sync(0xffffffffu, mg, 8);
                if (Pb <= 4) mg |= __shfl_xor_sync(0xffffffffu, mg, 4);
                if (my_seed >= t0 && my_seed < t0 + kSuper) mg &= ~(1u << (my_seed - t0));   // the own seed tile is in the list already
                uint32_t w = __reduce_or_sync(0xffffffffu, lane < Pb ? mg : 0u);
                while (w) {
                    const int b = __ffs(w) - 1;
                    w &= w - 1;
                    const uint32_t set = __ballot_sync(0xffffffffu, (mg >> b) & 1u) & low;
                    if (lane == 0) { tlist[nu] = t0 + b; tmask[nu] = set; }
                    ++nu;
                }
            }
            if (nu == 0) break;
            __syncwarp();
            run_tiles(nu);
            nu = 0;
            if (g0 >= ngroups) break;
            tau = bj < im.npx ? recs[bj].tau : -CUDART_INF_F;
        }
        __syncwarp();
        }       // sign classes
        for (int i = lane; i < im.npx * k; i += 32) gout[i] = lists[i];
        __syncwarp();
        if (pl.dbg == 7 && lane == 0 && stats && it < 30000) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            stats[4096 + it * 4 + 0] = dbg_t0;
            stats[4096 + it * 4 + 1] = t1;
            stats[4096 + it * 4 + 2] = n_eval - dbg_e0;
            stats[4096 + it * 4 + 3] = (unsigned long long)dbg_exact | ((unsigned long long)im.npx << 8);
        }
    }
    if (lane == 0 && stats) {
        atomicAdd(&stats[0], n_eval);
        atomicAdd(&stats[1], n_all - n_eval);
        if (pl.dbg == 7) {                                 // timing experiment: when did this warp run out of items (globaltimer, ns)
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            stats[4 + blockIdx.x * C::kWpc + warp] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// kernel 5: merge the partial lists of a pixel, map positions to original rows, decode winner rows
// ------------------------------------------------------------------------------------------------------
template <int ENC> __device__ __forceinline__ float decode_chi_comp(const DbSlot& db, int pos, int c) {
    const int tile = pos / kTile, in = pos - tile * kTile;
    const size_t e = ((size_t)tile * kNC + c) * kTile + in;
    const float sc = db.scale[chi_comp(c)];
    if (ENC == CLEDB_ENC_I16) return __half2float(__ushort_as_half(reinterpret_cast<const unsigned short*>(db.tiles)[e])) * sc;
    return reinterpret_cast<const float*>(db.tiles)[e];
}
template <int ENC> __device__ __forceinline__ float decode_v(const DbSlot& db, int pos, int which) {
    const size_t e = (size_t)pos * 2 + which;
    if (ENC == CLEDB_ENC_I16)
        return __half2float(__ushort_as_half(reinterpret_cast<const unsigned short*>(db.vcomp)[e])) * db.scale[which ? 7 : 3];
    return reinterpret_cast<const float*>(db.vcomp)[e];
}
__device__ __forceinline__ void decode_row(const DbSlot& db, int pos, float* row) {
    row[0] = 1.0f;
    if (db.encoding == CLEDB_ENC_I16) {
        row[1] = decode_chi_comp<CLEDB_ENC_I16>(db, pos, 0); row[2] = decode_chi_comp<CLEDB_ENC_I16>(db, pos, 1);
        row[3] = decode_v<CLEDB_ENC_I16>(db, pos, 0);        row[4] = decode_chi_comp<CLEDB_ENC_I16>(db, pos, 2);
        row[5] = decode_chi_comp<CLEDB_ENC_I16>(db, pos, 3); row[6] = decode_chi_comp<CLEDB_ENC_I16>(db, pos, 4);
        row[7] = decode_v<CLEDB_ENC_I16>(db, pos, 1);
    } else {
        row[1] = decode_chi_comp<CLEDB_ENC_F32>(db, pos, 0); row[2] = decode_chi_comp<CLEDB_ENC_F32>(db, pos, 1);
        row[3] = decode_v<CLEDB_ENC_F32>(db, pos, 0);        row[4] = decode_chi_comp<CLEDB_ENC_F32>(db, pos, 2);
        row[5] = decode_chi_comp<CLEDB_ENC_F32>(db, pos, 3); row[6] = decode_chi_comp<CLEDB_ENC_F32>(db, pos, 4);
        row[7] = decode_v<CLEDB_ENC_F32>(db, pos, 1);
    }
}

constexpr int kMaxSub = 24;   // 3 sign partitions x up to 8 chunks
constexpr int kFinPix = 64;   // pixels per block of the finalise kernel

// Two phases per block of 64 pixels.  (1) one thread per pixel merges the pixel's partial lists (chunks / sign classes)
// into shared memory: value, stored position, original row of its nsearch winners.  (2) one thread per (pixel, slot)
// decodes the winner's row and writes idx / chi / kpos / rows: consecutive threads write consecutive elements.
template <int PREC>
__global__ void __launch_bounds__(256) k_finalize(Plan pl, const DbSlot* __restrict__ slots, int32_t* __restrict__ idx_out,
                                                   float* __restrict__ chi_out, double* __restrict__ chi64_out,
                                                   int32_t* __restrict__ kpos_out, float* __restrict__ rows_out,
                                                   int64_t* __restrict__ ncand_out) {
    using V = typename Val<PREC>::type;
    using CandT = Cand<PREC>;
    extern __shared__ __align__(16) unsigned char fsm[];
    const int k = pl.k;
    V* s_val = reinterpret_cast<V*>(fsm);                                        // [kFinPix][k]
    int32_t* s_pos = reinterpret_cast<int32_t*>(s_val + kFinPix * k);            // [kFinPix][k] stored position, -1 = empty slot
    int32_t* s_idx = s_pos + kFinPix * k;                                        // [kFinPix][k] original row
    int32_t* s_key = s_idx + kFinPix * k;                                        // [kFinPix]
    const int64_t p0 = (int64_t)blockIdx.x * kFinPix;

    // phase 1, one thread per (pixel, slot): a pixel with ONE list (the pruned search, unsplit classes) has nothing to merge -
    // slot r is entry r of the list; a pixel with several lists is merged by the thread of its slot 0
    for (int t = threadIdx.x; t < kFinPix * k; t += blockDim.x) {
        const int lp = t / k, r0 = t - lp * k;
        const int64_t p = p0 + lp;
        int key = -1;
        if (p < pl.npix) key = pl.key[p];
        if (r0 == 0) s_key[lp] = key;
        if (key < 0) {
            s_pos[t] = -1;
            if (r0 == 0 && ncand_out && p < pl.npix) ncand_out[p] = 0;
            continue;
        }
        const DbSlot& db = slots[key / kParts];
        const int local = pl.rank[p] - pl.off[key];
        const int tile = local / pl.P, j = local - tile * pl.P;
        const int ns = pl.nsub[key];
        const int ntg = (pl.cnt[key] + pl.P - 1) / pl.P;                    // item numbering: see make_item
        const CandT* base = reinterpret_cast<const CandT*>(pl.plist) + ((size_t)(pl.item0[key] + (pl.seed ? tile : tile * ns)) * pl.P + j) * k;
        if (ns == 1) {
            const CandT c = base[r0];
            s_val[t] = c.v;
            s_pos[t] = c.pos;                                  // a list is filled from the front: empty slots (-1) come last
            s_idx[t] = c.pos >= 0 ? db.perm[c.pos] : 0x7fffffff;
        } else if (r0 == 0) {
            const size_t sub_stride = (size_t)(pl.seed ? ntg : 1) * pl.P * k;
            unsigned char head[kMaxSub];
            for (int s = 0; s < ns; ++s) head[s] = 0;
            for (int r = 0; r < k; ++r) {
                int best = -1, bpos = -1, bidx = 0x7fffffff;
                V bv = inf_of(V());
                for (int s = 0; s < ns; ++s) {
                    if (head[s] >= k) continue;
                    const CandT c = base[s * sub_stride + head[s]];
                    if (c.pos < 0) { head[s] = (unsigned char)k; continue; }
                    const int oi = db.perm[c.pos];             // lists carry stored positions; ties go to the lowest original row
                    if (best < 0 || c.v < bv || (c.v == bv && oi < bidx)) { best = s; bv = c.v; bpos = c.pos; bidx = oi; }
                }
                if (best >= 0) head[best]++;
                s_val[lp * k + r] = bv;
                s_pos[lp * k + r] = best >= 0 ? bpos : -1;
                s_idx[lp * k + r] = bidx;
            }
        }
        if (r0 == 0 && ncand_out) {
            const int cls = key % kParts;
            ncand_out[p] = pl.mode == CLEDB_MODE_IQUV ? db.part_cnt[cls] : (int64_t)db.part_cnt[0] + db.part_cnt[1] + db.part_cnt[2];
        }
    }
    __syncthreads();
    const int64_t nout = min((int64_t)kFinPix, pl.npix - p0) * k;
    for (int64_t t = threadIdx.x; t < nout; t += blockDim.x) {
        const int lp = (int)(t / k);
        const int64_t o = p0 * k + t;
        const int bpos = s_pos[t];
        float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (bpos < 0) {
            idx_out[o] = -1;
            chi_out[o] = 0.f;
            if (chi64_out) chi64_out[o] = 0.0;
            if (kpos_out) kpos_out[o] = -1;
        } else {
            const DbSlot& db = slots[s_key[lp] / kParts];
            const int bidx = s_idx[t];
            const double chi = (PREC == CLEDB_PREC_FP32) ? 0.5 * ((double)s_val[t] + pl.c0[p0 + lp]) : (double)s_val[t];
            idx_out[o] = bidx;
            chi_out[o] = (float)chi;
            if (chi64_out) chi64_out[o] = chi;
            if (kpos_out) {
                if (pl.mode == CLEDB_MODE_IQUV) kpos_out[o] = db.row_kpos[bidx];
                else kpos_out[o] = db.kept_rank ? db.kept_rank[bidx] : bidx;
            }
            if (rows_out) decode_row(db, bpos, row);
        }
        if (rows_out) {
            reinterpret_cast<float4*>(rows_out)[2 * o] = make_float4(row[0], row[1], row[2], row[3]);
            reinterpret_cast<float4*>(rows_out)[2 * o + 1] = make_float4(row[4], row[5], row[6], row[7]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
template <int ENC, int PREC> static size_t match_smem_bytes(int) {
    return (size_t)KCfg<ENC, PREC>::kWpc * KCfg<ENC, PREC>::kWarpBytes + 128;
}

template <int ENC, int PREC> static int configure_match(cledb_ctx* ctx, int k, int* occ_out, size_t* smem_out) {
    const size_t smem = match_smem_bytes<ENC, PREC>(k);
    CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_match<ENC, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (PREC == CLEDB_PREC_FP32)
        CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_match<ENC, PREC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CLEDB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_match<ENC, PREC>, KCfg<ENC, PREC>::kWpc * 32, smem));
    if (occ < 1) { ctx->err = "search kernel does not fit on an SM"; return CLEDB_ERR_CUDA; }
    *occ_out = occ;
    *smem_out = smem;
    return CLEDB_OK;
}

int match_occupancy(cledb_ctx* ctx) {
    int occ = 0;
    size_t smem = 0;
    int rc = configure_match<CLEDB_ENC_I16, CLEDB_PREC_FP32>(ctx, 8, &occ, &smem);
    if (rc) return rc;
    cudaFuncAttributes fa;
    CLEDB_CUDA(ctx, cudaFuncGetAttributes(&fa, k_match<CLEDB_ENC_I16, CLEDB_PREC_FP32>));
    ctx->occ = occ;
    ctx->regs = fa.numRegs;
    ctx->smem = (int)smem;
    return CLEDB_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Work-shape choice once the bucket sizes are known (host side, after the classify kernel): pixels per warp item P and
// chunks per candidate list.  Model, in units of one (pixel, tile) step of the hot loop: an item of P pixels over T tiles
// costs T*(P + c_tile) + P*c_boot + P*k*ln(T+1)*c_event; with no more items than resident warps the kernel lasts as long
// as its longest item, otherwise the warps pull items from a queue: the work spreads evenly, plus a penalty for a partly
// filled last round (constants fitted to B200 timings of the 256x256 map at P = 9..32, profiles/r01_tuning.txt).
static void choose_shape(cledb_ctx* ctx, const int32_t* cnt, int nkeys, int mode, int k, int64_t wslots, int pmax, int* P_out,
                         int* split_out) {
    const double c_tile = 2.5, c_boot = 5.0, c_event = 2.5;
    double best = 1e300;
    int bp = pmax, bs = 1;
    const int p_lo = ctx->pixels_per_item > 0 ? std::min(ctx->pixels_per_item, pmax) : 1, p_hi = ctx->pixels_per_item > 0 ? p_lo : pmax;
    const int s_lo = ctx->split > 0 ? ctx->split : 1, s_hi = ctx->split > 0 ? ctx->split : 8;
    for (int P = p_hi; P >= p_lo; --P) {
        for (int split = s_lo; split <= s_hi; ++split) {
            double work = 0.0, longest = 0.0;
            int64_t items = 0;
            for (int key = 0; key < nkeys; ++key) {
                const int c = cnt[key];
                if (c <= 0) continue;
                const DbSlot& db = ctx->dbs[key / kParts].slot;
                const int64_t ptiles = (c + P - 1) / P;
                for (int q = 0; q < kParts; ++q) {
                    if (mode == CLEDB_MODE_IQUV ? q != key % kParts : db.part_cnt[q] == 0) continue;
                    const double T = std::max(1.0, (double)db.part_ntiles[q] / split);
                    const double cost = T * (P + c_tile) + P * c_boot + P * k * std::log(T + 1.0) * c_event;
                    work += cost * ptiles * split;
                    items += ptiles * split;
                    longest = std::max(longest, cost);
                }
            }
            if (items == 0) continue;
            const double rounds = (double)items / (double)wslots, f = rounds - std::floor(rounds);
            const double span = items <= wslots ? longest : work / (double)wslots + 0.8 * f * (1.0 - f) * work / (double)items;
            if (span < best * 0.999) { best = span; bp = P; bs = split; }
        }
    }
    *P_out = bp;
    *split_out = bs;
}

// bucket counts and the three counters into mapped host memory
__global__ void k_export_counts(const int32_t* __restrict__ cnt, int nkeys, const int32_t* __restrict__ counters, int32_t* __restrict__ host) {
    for (int i = threadIdx.x; i < nkeys; i += blockDim.x) host[i] = cnt[i];
    if (threadIdx.x < 3) host[nkeys + threadIdx.x] = counters[threadIdx.x];
    __threadfence_system();
}

// A searched pixel that names a database which is not resident sets a word in mapped host memory (k_classify).  The host
// looks at it whenever it has synchronised with the device anyway - never in the middle of an enqueue.
int check_async_error(cledb_ctx* ctx) {
    if (ctx->h_err && *reinterpret_cast<volatile int32_t*>(ctx->h_err) != 0) {
        *reinterpret_cast<volatile int32_t*>(ctx->h_err) = 0;
        ctx->err = "cledb_match: a pixel refers to a database id that is not resident (cledb_db_put it first)";
        return CLEDB_ERR_NODB;
    }
    if (ctx->h_err && reinterpret_cast<volatile int32_t*>(ctx->h_err)[1] != 0) {      // comm.cu: a flag wait gave up
        reinterpret_cast<volatile int32_t*>(ctx->h_err)[1] = 0;
        ctx->err = "multi-GPU: a peer rank did not arrive at a collective within 60 s (peer-memory transport); the gathered maps are incomplete";
        return CLEDB_ERR_COMM;
    }
    return CLEDB_OK;
}

template <int PREC>
static int run_match(cledb_ctx* ctx, int enc, Plan& pl, int64_t wslots, int pmax, const float* sobs, const float* snr,
                     const int32_t* db_id, int32_t* idx_out, float* chi_out, double* chi64_out, int32_t* kpos_out,
                     float* rows_out, int64_t* ncand_out, cudaStream_t st) {
    const int64_t npix = pl.npix;
    const int nb = (int)((npix + 255) / 256);
    k_classify<PREC><<<nb, 256, 0, st>>>(pl, sobs, snr, db_id, ctx->d_slots, ctx->d_slot_of_id, ctx->id_cap);
    ctx->launches += 1;
    // k-d tile layout resident (fp32 search only): exact tile pruning (cledb_set_pruning 1, the default) or the seeded brute
    // force (cledb_set_pruning 2: the same layout and seed tiles, no bounds - every (pixel, entry) pair is evaluated)
    const bool blocked = PREC == CLEDB_PREC_FP32 && pl.seed != nullptr;
    const bool pruned = blocked && ctx->prune != 2;
    // IQUD: one item per pixel group walks all sign classes with one list per pixel (17 % fewer tiles evaluated; 1024^2 map
    // 10.2 -> 7.8 ms, 256^2 map 1.02 -> 0.95 ms)
    pl.fused = (pruned && pl.mode == CLEDB_MODE_IQUD) ? 1 : 0;
    // heaviest-first queue from 32 Ki pixels on; smaller maps last as long as their longest item whatever the order, and index
    // order keeps neighbouring items - which share tiles - on one SM (128^2: 10 % faster).  dbg 8: index order (timing experiment)
    if (!pruned || pl.dbg == 8 || npix < 32768) { pl.wpix = nullptr; pl.qorder = nullptr; }
    int64_t tiles_of_pixels = 0;
    if (pruned) {
        // The pruned search needs nothing from the host in mid-call: its work shape follows from the pixel count alone - 4, 8
        // or 32 pixels per item (measured on 128^2 .. 1024^2 maps: small maps need many small items to keep every warp busy
        // to the end, big maps amortise the per-item work over more pixels) - and the number of items is bounded by
        // npix/P + one partly filled item per bucket.  The call is a pure enqueue: no synchronisation, capturable in a graph.
        pl.P = npix >= 512 * wslots ? 32 : (npix >= 64 * wslots ? 8 : 4);
        // fused IQUD items are twice as long: fewer, larger items balance better (profiles/exp_ppi.py: 4 below 42 Ki pixels,
        // 8 below 100 Ki, 16 up to 400 Ki, 32 above)
        if (pl.fused) pl.P = npix >= 225 * wslots ? 32 : (npix >= 56 * wslots ? 16 : (npix >= 24 * wslots ? 8 : 4));
        pl.P = std::min(pl.P, pmax);
        if (ctx->pixels_per_item > 0) pl.P = std::min(std::min(32, pmax), ctx->pixels_per_item);
        pl.split = 1;
        tiles_of_pixels = npix / pl.P + pl.nkeys;
    } else {
        // ---- brute force: the bucket sizes come back to the host (a few hundred bytes, one stream synchronisation) and fix
        // the work shape.  The counts travel through mapped pinned memory written by a tiny kernel, not through a copy-engine
        // transfer: a device-to-host copy would queue behind the result maps of the previous chunk of a big map.
        if ((int)ctx->h_cnt_cap < pl.nkeys + 4) {
            if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
            ctx->h_cnt = nullptr;
            CLEDB_CUDA(ctx, cudaHostAlloc(&ctx->h_cnt, sizeof(int32_t) * (pl.nkeys + 4), cudaHostAllocMapped));
            CLEDB_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_cnt_mapped, ctx->h_cnt, 0));
            ctx->h_cnt_cap = pl.nkeys + 4;
        }
        k_export_counts<<<1, 256, 0, st>>>(pl.cnt, pl.nkeys, pl.counters, ctx->d_cnt_mapped);
        ctx->launches += 1;
        CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
        int rc0 = check_async_error(ctx);
        if (rc0) return rc0;
        choose_shape(ctx, ctx->h_cnt, std::min<int>(pl.nkeys, (int)ctx->dbs.size() * kParts), pl.mode, pl.k, wslots, pmax, &pl.P, &pl.split);
        for (int key = 0; key < pl.nkeys; ++key) tiles_of_pixels += (ctx->h_cnt[key] + pl.P - 1) / pl.P;
    }
    ctx->last_P = pl.P;
    ctx->last_split = pl.split;
    const size_t cand_sz = sizeof(Cand<PREC>);
    const int64_t max_items = tiles_of_pixels * (int64_t)((pl.mode == CLEDB_MODE_IQUV || pl.fused) ? 1 : kParts) * pl.split;
    int rc = ensure_workspace(ctx, ctx->ws2, std::max<size_t>(256, (size_t)max_items * pl.P * pl.k * cand_sz));
    if (rc) return rc;
    pl.plist = ctx->ws2.ptr;

    k_scan<<<1, 1024, 0, st>>>(pl, ctx->d_slots);
    if (blocked) {
        k_scatter<<<nb, 256, 0, st>>>(pl);                 // bucket order first: k_seed then stages one database per block
        k_seed<<<(int)((npix + 64 * kSeedRounds - 1) / (64 * kSeedRounds)), 256, 0, st>>>(pl, ctx->d_slots);
        k_fscan<<<ctx->slots_cap, 256, 0, st>>>(pl);
        k_scatter_fine<<<nb, 256, 0, st>>>(pl, ctx->d_slots);
        ctx->launches += 3;
        pl.qpiece = pl.P >= 16 ? 4 : 1;
        // pieces only on small maps and only for the items 16 x the median weight and more: single pixels lose the tile sharing of
        // their item, so the total work grows (256^2 IQUV kernel 425 -> 387 us; from 512^2 on the queue order alone balances
        // the load and pieces cost 5 - 10 %: profiles/r02_tuning.txt)
        pl.qsplit = npix < 100000 ? 8 : 0;
        if (const char* e = std::getenv("CLEDB_B200_QSPLIT")) pl.qsplit = std::atoi(e);
        if (pruned && pl.qorder && (max_items > pl.qcap || max_items * ((pl.P + pl.qpiece - 1) / pl.qpiece) > pl.qocap || max_items >= (1 << 24)))
            pl.qorder = nullptr;                                                  // tiny items asked for by cledb_set_tuning, huge maps
        if (pruned && pl.qorder) {
            const int nbi = (int)((max_items + 255) / 256);
            CLEDB_CUDA(ctx, cudaMemsetAsync(pl.qorder, 0xff, (size_t)max_items * ((pl.P + pl.qpiece - 1) / pl.qpiece) * 4, st));
            k_item_weight<<<nbi, 256, 0, st>>>(pl, ctx->d_slots);
            k_item_order<<<nbi, 256, 0, st>>>(pl);
            ctx->launches += 2;
        }
    } else k_scatter<<<nb, 256, 0, st>>>(pl);
    ctx->launches += 2;
    int occ = 0;
    size_t smem = 0;
    if (enc == CLEDB_ENC_I16) rc = configure_match<CLEDB_ENC_I16, PREC>(ctx, pl.k, &occ, &smem);
    else rc = configure_match<CLEDB_ENC_F32, PREC>(ctx, pl.k, &occ, &smem);
    if (rc) return rc;
    const int wpc = enc == CLEDB_ENC_I16 ? KCfg<CLEDB_ENC_I16, PREC>::kWpc : KCfg<CLEDB_ENC_F32, PREC>::kWpc;
    const int grid = (int)std::min<int64_t>((int64_t)ctx->prop.multiProcessorCount * occ, std::max<int64_t>(1, (max_items + wpc - 1) / wpc));
    int pgrid = 1;
    if (pruned) {
        if (!ctx->d_prune_stats) CLEDB_CUDA(ctx, cudaMalloc(&ctx->d_prune_stats, 1 << 20));
        CLEDB_CUDA(ctx, cudaMemsetAsync(ctx->d_prune_stats, 0, ctx->dbg == 7 ? (1 << 20) : 16, st));
        const int pw = enc == CLEDB_ENC_I16 ? PCfg<CLEDB_ENC_I16>::kWpc : PCfg<CLEDB_ENC_F32>::kWpc;
        smem = (size_t)pw * (enc == CLEDB_ENC_I16 ? pruned_warp_bytes<CLEDB_ENC_I16>() : pruned_warp_bytes<CLEDB_ENC_F32>());
        if (enc == CLEDB_ENC_I16) CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_match_pruned<CLEDB_ENC_I16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CLEDB_CUDA(ctx, cudaFuncSetAttribute(k_match_pruned<CLEDB_ENC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pgrid = (int)std::min<int64_t>((int64_t)ctx->prop.multiProcessorCount, std::max<int64_t>(1, (max_items + pw - 1) / pw));
    }
    const bool prof = ctx->prof_on && ctx->prof_used < (int)ctx->prof.size();
    if (prof) CLEDB_CUDA(ctx, cudaEventRecord(ctx->prof[ctx->prof_used].first, st));
    if (pruned) {
        if (enc == CLEDB_ENC_I16) k_match_pruned<CLEDB_ENC_I16><<<pgrid, PCfg<CLEDB_ENC_I16>::kWpc * 32, smem, st>>>(pl, ctx->d_slots, ctx->d_prune_stats);
        else k_match_pruned<CLEDB_ENC_F32><<<pgrid, PCfg<CLEDB_ENC_F32>::kWpc * 32, smem, st>>>(pl, ctx->d_slots, ctx->d_prune_stats);
    } else if (blocked) {
        if (enc == CLEDB_ENC_I16) k_match<CLEDB_ENC_I16, PREC, true><<<grid, KCfg<CLEDB_ENC_I16, PREC>::kWpc * 32, smem, st>>>(pl, ctx->d_slots);
        else k_match<CLEDB_ENC_F32, PREC, true><<<grid, KCfg<CLEDB_ENC_F32, PREC>::kWpc * 32, smem, st>>>(pl, ctx->d_slots);
    } else if (enc == CLEDB_ENC_I16) k_match<CLEDB_ENC_I16, PREC><<<grid, KCfg<CLEDB_ENC_I16, PREC>::kWpc * 32, smem, st>>>(pl, ctx->d_slots);
    else k_match<CLEDB_ENC_F32, PREC><<<grid, KCfg<CLEDB_ENC_F32, PREC>::kWpc * 32, smem, st>>>(pl, ctx->d_slots);
    if (prof) {
        CLEDB_CUDA(ctx, cudaEventRecord(ctx->prof[ctx->prof_used].second, st));
        ctx->prof_used++;
    }
    const size_t fin_smem = (size_t)kFinPix * pl.k * (sizeof(typename Val<PREC>::type) + 8) + kFinPix * 4;
    k_finalize<PREC><<<(int)((npix + kFinPix - 1) / kFinPix), 256, fin_smem, st>>>(pl, ctx->d_slots, idx_out, chi_out, chi64_out,
                                                                                  kpos_out, rows_out, ncand_out);
    ctx->launches += 2;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

int match_device(cledb_ctx* ctx, int mode, int precision, int64_t npix, const float* sobs, const float* snr,
                 const int32_t* db_id, int nsearch, int32_t* idx_out, float* chi_out, double* chi64_out,
                 int32_t* kpos_out, float* rows_out, uint8_t* status_out, int64_t* ncand_out, cudaStream_t st) {
    if (npix < 0 || !sobs || !snr || !db_id || !idx_out || !chi_out) { ctx->err = "cledb_match: null argument"; return CLEDB_ERR_ARG; }
    if (nsearch < 1 || nsearch > CLEDB_MAX_NSEARCH) { ctx->err = "cledb_match: nsearch must be in 1..32"; return CLEDB_ERR_ARG; }
    if (mode != CLEDB_MODE_IQUV && mode != CLEDB_MODE_IQUD) { ctx->err = "cledb_match: unknown mode"; return CLEDB_ERR_ARG; }
    if (precision != CLEDB_PREC_FP32 && precision != CLEDB_PREC_FP64) { ctx->err = "cledb_match: unknown precision"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    if (npix > (int64_t)1 << 30) { ctx->err = "cledb_match: more than 2^30 pixels in one call"; return CLEDB_ERR_ARG; }
    int enc = -1;
    for (auto& h : ctx->dbs) {
        if (!h.slot.live) continue;
        if (enc < 0) enc = h.slot.encoding;
        else if (enc != h.slot.encoding) { ctx->err = "cledb_match: resident databases mix encodings"; return CLEDB_ERR_ARG; }
    }
    if (enc < 0) { ctx->err = "cledb_match: no database resident"; return CLEDB_ERR_NODB; }
    int rc = sync_tables(ctx, st);
    if (rc) return rc;
    if (ctx->occ == 0 && (rc = match_occupancy(ctx))) return rc;
    if (!ctx->h_err) {
        CLEDB_CUDA(ctx, cudaHostAlloc(&ctx->h_err, 64, cudaHostAllocMapped));
        std::memset(ctx->h_err, 0, 64);
        CLEDB_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_err_mapped, ctx->h_err, 0));
    }

    Plan pl;
    std::memset(&pl, 0, sizeof(pl));
    pl.err_mapped = ctx->d_err_mapped;
    pl.npix = npix;
    pl.mode = mode;
    pl.k = nsearch;
    pl.dbg = ctx->dbg;
    pl.nkeys = ctx->slots_cap * kParts;
    // Every warp streams the whole candidate list of its item past the item's pixels: more pixels per item amortise the
    // tile loads better, while the number of items should fill the resident warps in whole rounds (choose_shape).
    const int wpc_est = precision == CLEDB_PREC_FP32
                            ? (enc == CLEDB_ENC_I16 ? KCfg<CLEDB_ENC_I16, CLEDB_PREC_FP32>::kWpc : KCfg<CLEDB_ENC_F32, CLEDB_PREC_FP32>::kWpc)
                            : (enc == CLEDB_ENC_I16 ? KCfg<CLEDB_ENC_I16, CLEDB_PREC_FP64>::kWpc : KCfg<CLEDB_ENC_F32, CLEDB_PREC_FP64>::kWpc);
    int occ = ctx->occ;
    if (precision == CLEDB_PREC_FP64) occ = 1;
    const int64_t wslots = (int64_t)ctx->prop.multiProcessorCount * std::max(1, occ) * wpc_est;
    const int pmax = std::max(1, std::min(kMaxP, kListCap / nsearch));
    const size_t rec_sz = precision == CLEDB_PREC_FP32 ? sizeof(Rec<CLEDB_PREC_FP32>) : sizeof(Rec<CLEDB_PREC_FP64>);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_key = take(npix * 4), o_rank = take(npix * 4), o_order = take(npix * 4);
    const size_t o_small = take((size_t)(5 * pl.nkeys + 1 + 4) * 4);
    const size_t o_recs = take(npix * rec_sz), o_status = take(npix), o_c0 = take(precision == CLEDB_PREC_FP32 ? npix * 8 : 0);
    // exact tile pruning: only with the fp32 kernel, and only if every resident database was put in the pruning layout
    int nblocked = 0, nlive = 0, ftiles = 1;
    for (auto& h : ctx->dbs) {
        if (!h.slot.live) continue;
        ++nlive;
        nblocked += h.slot.blocked ? 1 : 0;
        ftiles = std::max(ftiles, (int)(h.npad / kTile) + kParts);
    }
    if (nblocked != 0 && nblocked != nlive) { ctx->err = "cledb_match: resident databases mix the pruning and the plain layout"; return CLEDB_ERR_ARG; }
    const bool pruned = precision == CLEDB_PREC_FP32 && nblocked > 0;
    const size_t nfine = pruned ? (size_t)ctx->slots_cap * ftiles : 0;
    const size_t o_seed = take(pruned ? npix * kParts * 4 : 0), o_fine = take(3 * nfine * 4);
    const size_t qcap = pruned ? (size_t)npix / 4 + pl.nkeys + 64 : 0;          // items at the smallest P, one partly filled item per bucket
    const size_t qocap = pruned ? (size_t)npix + (size_t)pl.nkeys * 32 + 64 : 0;  // every item in single-pixel pieces
    const size_t o_wpix = take(pruned ? npix * 4 : 0), o_qbin = take(qcap * 4), o_qorder = take(qocap * 4), o_qhist = take(pruned ? 132 * 4 : 0);
    if ((rc = ensure_workspace(ctx, ctx->ws, off))) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->ws.ptr);
    pl.key = reinterpret_cast<int32_t*>(w + o_key);
    pl.rank = reinterpret_cast<int32_t*>(w + o_rank);
    pl.order = reinterpret_cast<int32_t*>(w + o_order);
    int32_t* small = reinterpret_cast<int32_t*>(w + o_small);
    pl.cnt = small;
    pl.cursor = small + pl.nkeys;
    pl.off = small + 2 * pl.nkeys;
    pl.nsub = small + 3 * pl.nkeys;
    pl.item0 = small + 4 * pl.nkeys;                 // nkeys + 1 entries
    pl.counters = small + 5 * pl.nkeys + 1;
    pl.recs = w + o_recs;
    pl.c0 = reinterpret_cast<double*>(w + o_c0);
    pl.status = status_out ? status_out : (w + o_status);
    CLEDB_CUDA(ctx, cudaMemsetAsync(small, 0, (size_t)(5 * pl.nkeys + 1 + 4) * 4, st));
    if (pruned) {
        pl.seed = reinterpret_cast<int32_t*>(w + o_seed);
        pl.fcnt = reinterpret_cast<int32_t*>(w + o_fine);
        pl.fcursor = pl.fcnt + nfine;
        pl.foff = pl.fcursor + nfine;
        pl.ftiles = ftiles;                          // tiles of the largest database + one key per sign class
        CLEDB_CUDA(ctx, cudaMemsetAsync(pl.fcnt, 0, 2 * nfine * 4, st));
        pl.wpix = reinterpret_cast<float*>(w + o_wpix);
        pl.qbin = reinterpret_cast<int32_t*>(w + o_qbin);
        pl.qorder = reinterpret_cast<int32_t*>(w + o_qorder);
        pl.qhist = reinterpret_cast<int32_t*>(w + o_qhist);
        pl.qcap = (int64_t)qcap;
        pl.qocap = (int64_t)qocap;
        CLEDB_CUDA(ctx, cudaMemsetAsync(pl.qhist, 0, 132 * 4, st));
    }
    if (precision == CLEDB_PREC_FP32)
        rc = run_match<CLEDB_PREC_FP32>(ctx, enc, pl, wslots, pmax, sobs, snr, db_id, idx_out, chi_out, chi64_out, kpos_out, rows_out, ncand_out, st);
    else
        rc = run_match<CLEDB_PREC_FP64>(ctx, enc, pl, wslots, pmax, sobs, snr, db_id, idx_out, chi_out, chi64_out, kpos_out, rows_out, ncand_out, st);
    return rc;
}

// ------------------------------------------------------------------------------------------------------
// small database kernels: decoded read-back and row gather
// ------------------------------------------------------------------------------------------------------
__global__ void k_decode_db(const DbSlot* slots, int slot, const int32_t* invperm, float* out) {
    const DbSlot& db = slots[slot];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= db.n) return;
    const int pos = invperm[i];
    float row[8] = {1.f, 0.f, 0.f, CUDART_NAN_F, 0.f, 0.f, 0.f, 0.f};
    if (pos >= 0) decode_row(db, pos, row);
    for (int c = 0; c < 8; ++c) out[i * 8 + c] = row[c];
}
__global__ void k_gather_rows(const DbSlot* slots, int slot, const int32_t* invperm, const int32_t* idx, int64_t n, float* rows) {
    const DbSlot& db = slots[slot];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ix = idx[i];
    float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ix >= 0 && ix < db.n) {
        const int pos = invperm[ix];
        if (pos >= 0) decode_row(db, pos, row);
        else { row[0] = 1.f; row[3] = CUDART_NAN_F; }
    }
    for (int c = 0; c < 8; ++c) rows[i * 8 + c] = row[c];
}
int launch_decode_db(cledb_ctx* ctx, int slot, float* d_out, cudaStream_t st) {
    const int64_t n = ctx->dbs[slot].slot.n;
    k_decode_db<<<(int)((n + 255) / 256), 256, 0, st>>>(ctx->d_slots, slot, ctx->dbs[slot].d_invperm, d_out);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}
int launch_gather_rows(cledb_ctx* ctx, int slot, const int32_t* d_idx, int64_t n, float* d_rows, cudaStream_t st) {
    if (n == 0) return CLEDB_OK;
    k_gather_rows<<<(int)((n + 255) / 256), 256, 0, st>>>(ctx->d_slots, slot, ctx->dbs[slot].d_invperm, d_idx, n, d_rows);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

}  // namespace cledb
