// Database residency: sign partitioning, int16 encoding, upload.  Replaces the per-task pickling of
// database[db_enc[xx,yy]] (CLEDB_PROC.py:304-307) and the per-pixel argwhere/gather on sign(V1)
// (CLEDB_PROC.py:974-976) of the reference: both are done once per database here.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "cledb_internal.h"

using namespace cledb;

// ------------------------------------------------------------------------------------------------------
// codec: int16 code = binary16 bit pattern of x * 2^shift
// ------------------------------------------------------------------------------------------------------
static inline uint16_t f32_to_f16_bits_rn(float f) {
    // round-to-nearest-even float32 -> binary16, written out so that host results do not depend on F16C
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (uint16_t)(sign | (absx > 0x7f800000u ? 0x7e00u : 0x7c00u));   // NaN / inf
    if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);            // rounds to >= 65520 -> inf
    if (absx < 0x33000001u) return (uint16_t)sign;                         // <= 2^-25 -> +-0
    int32_t exp = (int32_t)(absx >> 23) - 127;
    uint32_t man = (absx & 0x7fffffu) | 0x800000u;                         // 24-bit significand
    int shift;                                                             // bits to drop
    uint32_t hexp;
    if (exp < -14) { shift = 13 + (-14 - exp); hexp = 0; }                 // subnormal half
    else { shift = 13; hexp = (uint32_t)(exp + 15); }
    const uint32_t keep = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u);
    const uint32_t halfway = 1u << (shift - 1);
    uint32_t h = (hexp << 10) + (hexp ? (keep & 0x3ffu) : keep);
    if (rem > halfway || (rem == halfway && (keep & 1u))) h += 1;          // carries propagate into the exponent
    return (uint16_t)(sign | h);
}
static inline float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            float v = std::ldexp((float)man, -24);
            std::memcpy(&x, &v, 4);
            x |= sign;
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

extern "C" int cledb_codec_shift(const float* x, int64_t n, int64_t stride) {
    float mx = 0.f;
    for (int64_t i = 0; i < n; ++i) {
        const float a = std::fabs(x[i * stride]);
        if (std::isfinite(a) && a > mx) mx = a;
    }
    if (mx == 0.f) return 0;
    int e;
    std::frexp(mx, &e);                 // mx = m * 2^e, m in [0.5, 1)  ->  mx * 2^(15-e) in [2^14, 2^15)
    int s = 15 - e;
    if (s > 100) s = 100;
    if (s < -100) s = -100;
    return s;
}
extern "C" void cledb_codec_encode(const float* x, int64_t n, int shift, int16_t* codes) {
    const float up = std::ldexp(1.0f, shift);      // a power of two: the product is exact (|shift| <= 100)
    for (int64_t i = 0; i < n; ++i) codes[i] = (int16_t)f32_to_f16_bits_rn(x[i] * up);
}
extern "C" void cledb_codec_decode(const int16_t* codes, int64_t n, int shift, float* x) {
    const float down = std::ldexp(1.0f, -shift);
    for (int64_t i = 0; i < n; ++i) x[i] = f16_bits_to_f32((uint16_t)codes[i]) * down;
}

// ------------------------------------------------------------------------------------------------------
namespace cledb {

int ensure_workspace(cledb_ctx* ctx, Workspace& w, size_t bytes) {
    if (bytes <= w.bytes) return CLEDB_OK;
    if (w.ptr) {
        CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CLEDB_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(w.ptr);
        w.ptr = nullptr;
        w.bytes = 0;
    }
    const size_t want = bytes + bytes / 4;
    if (cudaMalloc(&w.ptr, want) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "out of device memory for the search workspace";
        return CLEDB_ERR_NOMEM;
    }
    w.bytes = want;
    return CLEDB_OK;
}

int sync_tables(cledb_ctx* ctx, cudaStream_t st) {
    if (!ctx->tables_dirty) return CLEDB_OK;
    const int nslots = (int)ctx->dbs.size();
    int max_id = -1;
    for (auto& kv : ctx->slot_of_id) max_id = std::max(max_id, kv.first);
    if (nslots > ctx->slots_cap || !ctx->d_slots) {
        CLEDB_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->d_slots) cudaFree(ctx->d_slots);
        ctx->slots_cap = std::max(16, nslots * 2);
        CLEDB_CUDA(ctx, cudaMalloc(&ctx->d_slots, sizeof(DbSlot) * ctx->slots_cap));
    }
    if (max_id + 1 > ctx->id_cap || !ctx->d_slot_of_id) {
        CLEDB_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->d_slot_of_id) cudaFree(ctx->d_slot_of_id);
        ctx->id_cap = std::max(64, (max_id + 1) * 2);
        CLEDB_CUDA(ctx, cudaMalloc(&ctx->d_slot_of_id, sizeof(int32_t) * ctx->id_cap));
    }
    std::vector<DbSlot> hs(ctx->slots_cap);
    std::memset(hs.data(), 0, sizeof(DbSlot) * hs.size());
    for (int i = 0; i < nslots; ++i) hs[i] = ctx->dbs[i].slot;
    std::vector<int32_t> hid(ctx->id_cap, -1);
    for (auto& kv : ctx->slot_of_id) hid[kv.first] = kv.second;
    // synchronous copies from pageable memory: the tables are tiny and change only when databases change
    CLEDB_CUDA(ctx, cudaStreamSynchronize(st));
    CLEDB_CUDA(ctx, cudaMemcpy(ctx->d_slots, hs.data(), sizeof(DbSlot) * hs.size(), cudaMemcpyHostToDevice));
    CLEDB_CUDA(ctx, cudaMemcpy(ctx->d_slot_of_id, hid.data(), sizeof(int32_t) * hid.size(), cudaMemcpyHostToDevice));
    ctx->tables_dirty = false;
    return CLEDB_OK;
}

static void free_db(HostDb& h) {
    if (h.d_tiles) cudaFree(h.d_tiles);
    if (h.d_vcomp) cudaFree(h.d_vcomp);
    if (h.d_perm) cudaFree(h.d_perm);
    if (h.d_invperm) cudaFree(h.d_invperm);
    if (h.d_kept_rank) cudaFree(h.d_kept_rank);
    if (h.d_row_kpos) cudaFree(h.d_row_kpos);
    if (h.d_rows8) cudaFree(h.d_rows8);
    if (h.d_boxes) cudaFree(h.d_boxes);
    h = HostDb();
    std::memset(&h.slot, 0, sizeof(DbSlot));
}

}  // namespace cledb

extern "C" int cledb_db_drop(cledb_ctx* ctx, int db_id) {
    if (!ctx) return CLEDB_ERR_ARG;
    auto it = ctx->slot_of_id.find(db_id);
    if (it == ctx->slot_of_id.end()) { ctx->err = "cledb_db_drop: unknown database id"; return CLEDB_ERR_NODB; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    CLEDB_CUDA(ctx, cudaDeviceSynchronize());
    free_db(ctx->dbs[it->second]);
    ctx->free_slots.push_back(it->second);
    ctx->slot_of_id.erase(it);
    ctx->tables_dirty = true;
    return CLEDB_OK;
}

// k-d partition of ids[lo, hi) into leaves of kTile entries (the last one may be short): median split at a multiple of the
// tile size along the component with the widest weighted extent; the upper levels run their halves on separate threads.
static void kd_split(const float* db, int32_t* ids, int64_t lo, int64_t hi, int depth) {
    const int64_t nt = (hi - lo + kTile - 1) / kTile;
    if (nt <= 1) return;
    const float wdim[kNC] = {1.f, 1.f, 0.1f, 1.f, 1.f};                   // sqrt(wtg_fact) of the components, CLEDB_PROC.py:981
    int dim = 0;
    float widest = -1.f;
    for (int c = 0; c < kNC; ++c) {
        const int oc = chi_comp(c);
        float mn = INFINITY, mx = -INFINITY;
        for (int64_t j = lo; j < hi; ++j) { const float v = db[(int64_t)ids[j] * 8 + oc]; mn = std::min(mn, v); mx = std::max(mx, v); }
        const float ext = (mx - mn) * wdim[c];
        if (ext > widest) { widest = ext; dim = c; }
    }
    const int oc = chi_comp(dim);
    const int64_t mid = lo + (nt / 2) * kTile;
    std::nth_element(ids + lo, ids + mid, ids + hi, [&](int32_t x, int32_t y) {
        float vx = db[(int64_t)x * 8 + oc], vy = db[(int64_t)y * 8 + oc];
        if (std::isnan(vx)) vx = INFINITY;                                 // keep the order strict and weak
        if (std::isnan(vy)) vy = INFINITY;
        return vx < vy || (vx == vy && x < y);
    });
    if (depth < 3 && hi - lo > 32768) {
        std::thread left(kd_split, db, ids, lo, mid, depth + 1);
        kd_split(db, ids, mid, hi, depth + 1);
        left.join();
    } else {
        kd_split(db, ids, lo, mid, depth + 1);
        kd_split(db, ids, mid, hi, depth + 1);
    }
}

extern "C" int cledb_db_put(cledb_ctx* ctx, int db_id, const float* db, int64_t n, int encoding) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (!db || n <= 0 || db_id < 0 || db_id > (1 << 24)) { ctx->err = "cledb_db_put: bad argument"; return CLEDB_ERR_ARG; }
    if (n > (int64_t)1 << 29) { ctx->err = "cledb_db_put: more than 2^29 rows"; return CLEDB_ERR_ARG; }
    if (encoding != CLEDB_ENC_F32 && encoding != CLEDB_ENC_I16) { ctx->err = "cledb_db_put: unknown encoding"; return CLEDB_ERR_ARG; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));

    // ---- validate the layout contract; find the rows that can ever be candidates (V1 not NaN)
    std::vector<int8_t> cls(n);
    int64_t cnt[kParts] = {0, 0, 0}, nnan = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* r = db + i * 8;
        // A row whose V1 is NaN is never a candidate in the reference (np.sign(NaN) == ... is False, CLEDB_PROC.py:974; the
        // "crutch" of :1148 drops it in IQUD), whatever its other components hold - e.g. the all-NaN row that the division
        // by I1 == 0 leaves at a zero-intensity grid point.  Such rows are left out before anything else is checked.
        if (std::isnan(r[3])) { cls[i] = -1; ++nnan; continue; }
        if (r[0] != 1.0f) {
            ctx->err = "cledb_db_put: component 0 of candidate row " + std::to_string(i) + " is not exactly 1 (database must be normalised by I1, CLEDB_PREPINV.py:306-317)";
            return CLEDB_ERR_DBFORMAT;
        }
        cls[i] = 0;
        if (!(std::isfinite(r[1]) && std::isfinite(r[2]) && std::isfinite(r[4]) && std::isfinite(r[5]) && std::isfinite(r[6]))) {
            ctx->err = "cledb_db_put: non-finite Stokes I/Q/U in candidate row " + std::to_string(i);
            return CLEDB_ERR_DBFORMAT;
        }
    }
    HostDb h;
    std::memset(&h.slot, 0, sizeof(DbSlot));
    // ---- block exponents per component (int16 encoding); scale = 2^-shift
    for (int c = 0; c < 8; ++c) { h.shift[c] = 0; h.slot.scale[c] = 1.0f; }
    if (encoding == CLEDB_ENC_I16) {
        for (int c = 1; c < 8; ++c) {
            float mx = 0.f;
            for (int64_t i = 0; i < n; ++i) {
                if (cls[i] < 0) continue;
                const float a = std::fabs(db[i * 8 + c]);
                if (std::isfinite(a) && a > mx) mx = a;
            }
            int s = 0;
            if (mx > 0.f) { int e; std::frexp(mx, &e); s = std::max(-100, std::min(100, 15 - e)); }
            h.shift[c] = s;
            h.slot.scale[c] = std::ldexp(1.0f, -s);
        }
    }
    float up2[8];                                  // 2^shift per component: scaling by a power of two is exact
    for (int c = 0; c < 8; ++c) up2[c] = std::ldexp(1.0f, h.shift[c]);
    // ---- sign classes of the DECODED V1 (the decoded array is the database: a V1 that underflows the int16
    //      code is a zero, exactly as the oracle sees it), CLEDB_PROC.py:974
    for (int64_t i = 0; i < n; ++i) {
        if (cls[i] < 0) continue;
        float v = db[i * 8 + 3];
        if (encoding == CLEDB_ENC_I16) v = f16_bits_to_f32(f32_to_f16_bits_rn(v * up2[3]));
        const int c = v > 0.f ? 0 : (v < 0.f ? 1 : 2);
        cls[i] = (int8_t)c;
        ++cnt[c];
    }
    int64_t tile = 0;
    int64_t pos0[kParts];
    for (int q = 0; q < kParts; ++q) {
        h.slot.part_tile0[q] = (int32_t)tile;
        h.slot.part_ntiles[q] = (int32_t)((cnt[q] + kTile - 1) / kTile);
        h.slot.part_cnt[q] = (int32_t)cnt[q];
        pos0[q] = tile * kTile;
        tile += h.slot.part_ntiles[q];
    }
    const int64_t ntiles = tile, npad = ntiles * kTile;
    h.npad = npad;
    h.nnan = nnan;

    // ---- fill the partitioned arrays
    const int elem = encoding == CLEDB_ENC_I16 ? 2 : 4;
    std::vector<unsigned char> tiles((size_t)std::max<int64_t>(1, npad) * kNC * elem);
    std::vector<unsigned char> vcomp((size_t)std::max<int64_t>(1, npad) * 2 * elem);
    std::vector<int32_t> perm(std::max<int64_t>(1, npad), -1), invperm(n, -1), kept_rank;
    {   // NaN padding: a padded entry can never win (x <= tau is false for NaN)
        if (encoding == CLEDB_ENC_I16) {
            uint16_t* t = reinterpret_cast<uint16_t*>(tiles.data());
            for (size_t i = 0; i < tiles.size() / 2; ++i) t[i] = 0x7e00u;
            uint16_t* v = reinterpret_cast<uint16_t*>(vcomp.data());
            for (size_t i = 0; i < vcomp.size() / 2; ++i) v[i] = 0x7e00u;
        } else {
            const float qnan = std::nanf("");
            float* t = reinterpret_cast<float*>(tiles.data());
            for (size_t i = 0; i < tiles.size() / 4; ++i) t[i] = qnan;
            float* v = reinterpret_cast<float*>(vcomp.data());
            for (size_t i = 0; i < vcomp.size() / 4; ++i) v[i] = qnan;
        }
    }
    if (nnan > 0) kept_rank.assign(n, -1);
    std::vector<int32_t> row_kpos(n, -1);
    // Stored order inside a sign class: a golden-ratio stride permutation of the original order.  The original
    // order walks (gx, phi, theta), so chi^2 is smooth along it and improvements of the running threshold would
    // arrive in bursts; with the stride order every tile samples the whole class evenly, the threshold converges
    // as for a random order (about nsearch*ln(N/nsearch) insertions per pixel) and insertions spread out in time.
    // Ties still resolve to the lowest ORIGINAL row: the top-k lists carry original indices.
    int64_t stride[kParts];
    for (int q = 0; q < kParts; ++q) {
        const int64_t c = cnt[q];
        int64_t st = 1;
        if (c > 2) {
            st = (int64_t)(0.6180339887498949 * (double)c);
            if (st < 1) st = 1;
            auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
            while (gcd(st, c) != 1) ++st;
        }
        stride[q] = st;
    }
    const bool blocked = ctx->prune != 0;
    // Pruning layout: the tiles of a sign class are the leaves of a k-d partition of its entries in the space of the five
    // chi^2 components (median splits at multiples of the tile size, along the component with the widest weighted extent),
    // so a tile is a small box in that space and the exact bound of k_match_pruned rejects most tiles of a pixel.
    std::vector<int32_t> kdpos;
    if (blocked) {
        kdpos.assign(n, -1);
        std::vector<int32_t> ids[kParts];
        std::vector<std::thread> workers;
        for (int q = 0; q < kParts; ++q) {
            ids[q].reserve(cnt[q]);
            for (int64_t i = 0; i < n; ++i) if (cls[i] == q) ids[q].push_back((int32_t)i);
            if (!ids[q].empty()) workers.emplace_back(kd_split, db, ids[q].data(), (int64_t)0, (int64_t)ids[q].size(), 0);
        }
        for (auto& w : workers) w.join();
        for (int q = 0; q < kParts; ++q)
            for (size_t j = 0; j < ids[q].size(); ++j) kdpos[ids[q][j]] = (int32_t)j;
    }
    std::vector<float> boxes;
    if (blocked) {
        boxes.assign((size_t)std::max<int64_t>(1, ntiles) * 16, 0.f);
        for (int64_t t = 0; t < ntiles; ++t)
            for (int c = 0; c < kNC; ++c) { boxes[t * 16 + c] = INFINITY; boxes[t * 16 + 5 + c] = -INFINITY; boxes[t * 16 + 10 + c] = NAN; }
    }
    int64_t seen[kParts] = {0, 0, 0};
    int32_t zero_bits[kParts] = {0, 0, 0};
    int64_t rank = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int q = cls[i];
        if (q < 0) continue;
        if (nnan > 0) kept_rank[i] = (int32_t)rank;
        ++rank;
        const int64_t r = seen[q]++;                                  // rank inside the class, original order
        row_kpos[i] = (int32_t)r;
        int64_t pos;
        if (!blocked) pos = pos0[q] + (int64_t)(((__int128)r * stride[q]) % cnt[q]);
        else pos = pos0[q] + kdpos[i];
        perm[pos] = (int32_t)i;
        invperm[i] = (int32_t)pos;
        const int64_t tl = pos / kTile, in = pos % kTile;
        const float* rr = db + i * 8;
        for (int c = 0; c < kNC; ++c) {
            const int oc = chi_comp(c);
            const size_t e = ((size_t)tl * kNC + c) * kTile + in;
            float stored;                         // the value the kernel multiplies: the half code as a float, or the float
            if (encoding == CLEDB_ENC_I16) {
                const uint16_t code = f32_to_f16_bits_rn(rr[oc] * up2[oc]);
                reinterpret_cast<uint16_t*>(tiles.data())[e] = code;
                if ((code & 0x7fffu) == 0) zero_bits[q] |= 1 << c;
                stored = f16_bits_to_f32(code);
            } else {
                reinterpret_cast<float*>(tiles.data())[e] = rr[oc];
                if (rr[oc] == 0.f) zero_bits[q] |= 1 << c;
                stored = rr[oc];
            }
            if (blocked) {
                boxes[tl * 16 + c] = std::min(boxes[tl * 16 + c], stored);
                boxes[tl * 16 + 5 + c] = std::max(boxes[tl * 16 + 5 + c], stored);
                if (in == 0) boxes[tl * 16 + 10 + c] = stored;          // representative entry, replaced below
            }
        }
        for (int w = 0; w < 2; ++w) {
            const int oc = w ? 7 : 3;
            if (encoding == CLEDB_ENC_I16)
                reinterpret_cast<uint16_t*>(vcomp.data())[pos * 2 + w] = f32_to_f16_bits_rn(rr[oc] * up2[oc]);
            else
                reinterpret_cast<float*>(vcomp.data())[pos * 2 + w] = rr[oc];
        }
    }
    for (int q = 0; q < kParts; ++q) h.slot.part_zero[q] = zero_bits[q];
    if (blocked) {
        // representative entry of a tile: the stored entry nearest to the middle of its box
        const float wdim[kNC] = {1.f, 1.f, 0.1f, 1.f, 1.f};
        for (int64_t t = 0; t < ntiles; ++t) {
            float* bx = &boxes[t * 16];
            int best = -1;
            float bestd = INFINITY;
            for (int in = 0; in < kTile; ++in) {
                const int32_t row = perm[t * kTile + in];
                if (row < 0) continue;
                float dist = 0.f;
                for (int c = 0; c < kNC; ++c) {
                    const int oc = chi_comp(c);
                    const size_t e = ((size_t)t * kNC + c) * kTile + in;
                    const float v = encoding == CLEDB_ENC_I16 ? f16_bits_to_f32(reinterpret_cast<const uint16_t*>(tiles.data())[e])
                                                              : reinterpret_cast<const float*>(tiles.data())[e];
                    const float dd = (v - 0.5f * (bx[c] + bx[5 + c])) * h.slot.scale[oc] * wdim[c];
                    dist += dd * dd;
                }
                if (dist < bestd) { bestd = dist; best = in; }          // NaN distances never win
            }
            if (best < 0) continue;
            for (int c = 0; c < kNC; ++c) {
                const size_t e = ((size_t)t * kNC + c) * kTile + best;
                bx[10 + c] = encoding == CLEDB_ENC_I16 ? f16_bits_to_f32(reinterpret_cast<const uint16_t*>(tiles.data())[e])
                                                       : reinterpret_cast<const float*>(tiles.data())[e];
            }
        }
    }

    size_t sbox_off = 0;                           // floats: the group bounds follow the tile bounds in one allocation
    if (blocked) {
        sbox_off = boxes.size();
        int64_t sb = 0;
        for (int q = 0; q < kParts; ++q) {
            h.slot.part_sb0[q] = (int32_t)sb;
            const int64_t ng = (h.slot.part_ntiles[q] + kSuper - 1) / kSuper;
            for (int64_t g = 0; g < ng; ++g) {
                float bx[12];
                for (int c = 0; c < kNC; ++c) { bx[c] = INFINITY; bx[5 + c] = -INFINITY; }
                bx[10] = bx[11] = 0.f;
                const int64_t t0 = h.slot.part_tile0[q] + g * kSuper, t1 = std::min<int64_t>(t0 + kSuper, h.slot.part_tile0[q] + h.slot.part_ntiles[q]);
                for (int64_t t = t0; t < t1; ++t)
                    for (int c = 0; c < kNC; ++c) {
                        bx[c] = std::min(bx[c], boxes[t * 16 + c]);
                        bx[5 + c] = std::max(bx[5 + c], boxes[t * 16 + 5 + c]);
                    }
                boxes.insert(boxes.end(), bx, bx + 12);
            }
            sb += ng;
        }
    }

    // ---- upload
    auto fail = [&](const char* what) {
        cudaGetLastError();
        free_db(h);
        ctx->err = std::string("cledb_db_put: ") + what;
        return CLEDB_ERR_NOMEM;
    };
    auto fail_cuda = [&](const char* what) {
        const cudaError_t e = cudaGetLastError();
        free_db(h);
        ctx->err = std::string("cledb_db_put: ") + what + " (" + cudaGetErrorString(e) + ")";
        return CLEDB_ERR_CUDA;
    };
    if (cudaMalloc(&h.d_tiles, tiles.size()) != cudaSuccess) return fail("out of device memory (tiles)");
    if (cudaMalloc(&h.d_vcomp, vcomp.size()) != cudaSuccess) return fail("out of device memory (V components)");
    if (cudaMalloc(&h.d_perm, perm.size() * 4) != cudaSuccess) return fail("out of device memory (perm)");
    if (cudaMalloc(&h.d_invperm, invperm.size() * 4) != cudaSuccess) return fail("out of device memory (invperm)");
    if (nnan > 0 && cudaMalloc(&h.d_kept_rank, kept_rank.size() * 4) != cudaSuccess) return fail("out of device memory (rank)");
    if (cudaMalloc(&h.d_row_kpos, row_kpos.size() * 4) != cudaSuccess) return fail("out of device memory (kpos)");
    if (blocked) {
        if (cudaMalloc(&h.d_boxes, boxes.size() * 4) != cudaSuccess) return fail("out of device memory (tile bounds)");
        if (cudaMemcpy(h.d_boxes, boxes.data(), boxes.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    }
    if (cudaMemcpy(h.d_tiles, tiles.data(), tiles.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    if (cudaMemcpy(h.d_vcomp, vcomp.data(), vcomp.size(), cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    if (cudaMemcpy(h.d_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    if (cudaMemcpy(h.d_invperm, invperm.data(), invperm.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    if (nnan > 0) if (cudaMemcpy(h.d_kept_rank, kept_rank.data(), kept_rank.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    if (cudaMemcpy(h.d_row_kpos, row_kpos.data(), row_kpos.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail_cuda("upload failed");
    h.bytes = (int64_t)(tiles.size() + vcomp.size() + perm.size() * 4 + invperm.size() * 4 + kept_rank.size() * 4 + row_kpos.size() * 4);
    h.slot.tiles = h.d_tiles;
    h.slot.vcomp = h.d_vcomp;
    h.slot.perm = h.d_perm;
    h.slot.kept_rank = h.d_kept_rank;
    h.slot.row_pos = h.d_invperm;
    h.slot.row_kpos = h.d_row_kpos;
    h.slot.n = n;
    h.slot.encoding = encoding;
    h.slot.live = 1;
    h.slot.blocked = blocked ? 1 : 0;
    h.slot.boxes = h.d_boxes;
    h.slot.sboxes = h.d_boxes ? h.d_boxes + sbox_off : nullptr;

    // the new database is complete on the device: only now is an existing database of this id replaced, so that a
    // failed put (format, out of memory) leaves the old one usable
    if (ctx->slot_of_id.count(db_id) && cledb_db_drop(ctx, db_id) != CLEDB_OK) { free_db(h); return CLEDB_ERR_CUDA; }
    int slot;
    if (!ctx->free_slots.empty()) { slot = ctx->free_slots.back(); ctx->free_slots.pop_back(); ctx->dbs[slot] = h; }
    else { slot = (int)ctx->dbs.size(); ctx->dbs.push_back(h); }
    ctx->slot_of_id[db_id] = slot;
    ctx->tables_dirty = true;
    return CLEDB_OK;
}

extern "C" int cledb_db_info(cledb_ctx* ctx, int db_id, int64_t info[16]) {
    if (!ctx || !info) return CLEDB_ERR_ARG;
    auto it = ctx->slot_of_id.find(db_id);
    if (it == ctx->slot_of_id.end()) { ctx->err = "cledb_db_info: unknown database id"; return CLEDB_ERR_NODB; }
    const HostDb& h = ctx->dbs[it->second];
    info[0] = h.slot.n;
    info[1] = h.slot.part_cnt[0];
    info[2] = h.slot.part_cnt[1];
    info[3] = h.slot.part_cnt[2];
    info[4] = h.nnan;
    info[5] = h.slot.encoding;
    info[6] = h.bytes;
    info[7] = kTile;
    for (int c = 0; c < 8; ++c) info[8 + c] = h.shift[c];
    return CLEDB_OK;
}

extern "C" int cledb_db_decoded(cledb_ctx* ctx, int db_id, float* out) {
    if (!ctx || !out) return CLEDB_ERR_ARG;
    auto it = ctx->slot_of_id.find(db_id);
    if (it == ctx->slot_of_id.end()) { ctx->err = "cledb_db_decoded: unknown database id"; return CLEDB_ERR_NODB; }
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = sync_tables(ctx, ctx->stream);
    if (rc) return rc;
    const int64_t n = ctx->dbs[it->second].slot.n;
    float* d = nullptr;
    if (cudaMalloc(&d, (size_t)n * 32) != cudaSuccess) { cudaGetLastError(); ctx->err = "cledb_db_decoded: out of device memory"; return CLEDB_ERR_NOMEM; }
    rc = launch_decode_db(ctx, it->second, d, ctx->stream);
    if (rc == CLEDB_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("cledb_db_decoded: ") + cudaGetErrorString(e); rc = CLEDB_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}

extern "C" int cledb_gather_rows(cledb_ctx* ctx, int db_id, const int32_t* idx, int64_t n, float* rows_out) {
    if (!ctx || !idx || !rows_out || n < 0) return CLEDB_ERR_ARG;
    auto it = ctx->slot_of_id.find(db_id);
    if (it == ctx->slot_of_id.end()) { ctx->err = "cledb_gather_rows: unknown database id"; return CLEDB_ERR_NODB; }
    if (n == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = sync_tables(ctx, ctx->stream);
    if (rc) return rc;
    if ((rc = ensure_workspace(ctx, ctx->io, (size_t)n * 36 + 512))) return rc;
    int32_t* d_idx = reinterpret_cast<int32_t*>(ctx->io.ptr);
    float* d_rows = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ctx->io.ptr) + ((size_t)n * 4 + 255) / 256 * 256);
    CLEDB_CUDA(ctx, cudaMemcpyAsync(d_idx, idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_gather_rows(ctx, it->second, d_idx, n, d_rows, ctx->stream))) return rc;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(rows_out, d_rows, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CLEDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// database selection per pixel: sdb_findelongationanddens (CLEDB_PREPINV.py:458-468)
// ------------------------------------------------------------------------------------------------------
namespace cledb {
// One thread per pixel; the table is a few thousand rows that every lane reads at the same address (broadcast, cached).
__global__ void __launch_bounds__(256) k_db_select(int64_t npix, const float* __restrict__ yobs, const float* __restrict__ dobs,
                                                   int nfiles, const float* __restrict__ table, int32_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float y = yobs[p], d = dobs[p];
    if (isnan(y)) { out[p] = CLEDB_SELECT_NAN; return; }              // np.min is NaN, argwhere(x == NaN) is empty
    float best = CUDART_INF_F;
    for (int i = 0; i < nfiles; ++i) best = fminf(best, fabsf(__fsub_rn(__fdiv_rn(table[3 * i + 1], 100.0f), y)));
    int first = -1, last = -1;
    for (int i = 0; i < nfiles; ++i) {
        if (fabsf(__fsub_rn(__fdiv_rn(table[3 * i + 1], 100.0f), y)) == best) {
            if (first < 0) first = i;
            last = i;
        }
    }
    if (first < 0) { out[p] = CLEDB_SELECT_NAN; return; }             // only with NaN heights in the table
    if (last <= first) { out[p] = CLEDB_SELECT_EMPTY; return; }       // argmin of an empty slice raises
    int arg = first;
    float bd = fabsf(__fsub_rn(__fdiv_rn(table[3 * first + 2], 100.0f), d));
    if (!isnan(bd)) {                                                 // np.argmin returns the first NaN if there is one
        for (int i = first + 1; i < last; ++i) {
            const float v = fabsf(__fsub_rn(__fdiv_rn(table[3 * i + 2], 100.0f), d));
            if (isnan(v)) { arg = i; break; }
            if (v < bd) { bd = v; arg = i; }
        }
    }
    out[p] = (int32_t)table[3 * arg];
}
}  // namespace cledb

extern "C" int cledb_db_select_dev(cledb_ctx* ctx, int64_t npix, const float* yobs, const float* dobs, int nfiles,
                                   const float* table, int32_t* file_out, void* stream) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || nfiles < 1 || !yobs || !dobs || !table || !file_out) { ctx->err = "cledb_db_select: bad argument"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    k_db_select<<<(int)((npix + 255) / 256), 256, 0, st>>>(npix, yobs, dobs, nfiles, table, file_out);
    ctx->launches++;
    CLEDB_CUDA(ctx, cudaGetLastError());
    return CLEDB_OK;
}

extern "C" int cledb_db_select(cledb_ctx* ctx, int64_t npix, const float* yobs, const float* dobs, int nfiles, const float* table,
                               int32_t* file_out) {
    if (!ctx) return CLEDB_ERR_ARG;
    if (npix < 0 || nfiles < 1 || !yobs || !dobs || !table || !file_out) { ctx->err = "cledb_db_select: bad argument"; return CLEDB_ERR_ARG; }
    if (npix == 0) return CLEDB_OK;
    CLEDB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)npix;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_y = 0, o_d = up(n * 4), o_t = o_d + up(n * 4), o_o = o_t + up((size_t)nfiles * 12), total = o_o + up(n * 4);
    int rc = ensure_workspace(ctx, ctx->io, total);
    if (rc) return rc;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx->io.ptr);
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_y, yobs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_d, dobs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CLEDB_CUDA(ctx, cudaMemcpyAsync(w + o_t, table, (size_t)nfiles * 12, cudaMemcpyHostToDevice, ctx->stream));
    rc = cledb_db_select_dev(ctx, npix, reinterpret_cast<float*>(w + o_y), reinterpret_cast<float*>(w + o_d), nfiles,
                             reinterpret_cast<float*>(w + o_t), reinterpret_cast<int32_t*>(w + o_o), ctx->stream);
    if (rc) return rc;
    CLEDB_CUDA(ctx, cudaMemcpyAsync(file_out, w + o_o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CLEDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CLEDB_OK;
}
