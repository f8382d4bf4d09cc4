/* cledb_b200.h - C ABI of the B200-native CLEDB 2-line database search.
 *
 * The reference (arparaschiv/solar-coronal-inversion, CLEDB 0.7.0-beta) is pure Python and has no FFI layer;
 * its boundary for this path is the Python function surface of CLEDB_PROC / CLEDB_PREPINV.  This header is
 * the C-ABI a binding for that surface loads (ctypes.CDLL in solar-coronal-inversion_b200/cledb_b200/native.py;
 * INTEGRATION.md shows the stub a reference maintainer would add).  Each entry point names the reference
 * code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No C++ or torch types cross the boundary, no exceptions.
 *   - return 0 on success, a negative CLEDB_ERR_* otherwise; cledb_last_error() gives the text.
 *   - one context per CUDA device and per thread of control; calls on one context are serialised by the
 *     caller; different contexts are independent.  ONE call in flight per context: the "_dev" entry points take a
 *     stream, but every call on a context shares the context's workspaces (and may regrow them), so a second call on
 *     the same context - on whatever stream - may only be issued after the first has been enqueued, and calls on
 *     different streams of one context must be ordered by the caller (events).  Use one context per concurrent stream.
 *   - "host" entry points take C-contiguous host buffers (pageable or pinned), copy in, run, copy out and
 *     return when the results are in the output buffers.  "_dev" entry points take device pointers that are
 *     valid on the context's device and enqueue on the given cudaStream_t (passed as void*; NULL = the
 *     context's own stream) without waiting for the results.  With the default search (exact tile pruning) a "_dev" call
 *     is a pure enqueue - no host synchronisation, capturable in a CUDA graph once the workspaces have their size; only the
 *     brute-force search (pruning off, or the fp64 validation kernel) synchronises the stream once internally, after its
 *     classification kernel, to size its work items from the bucket counts.  Asynchronous errors: a searched pixel whose
 *     database id is not resident gets status CLEDB_PIX_NULL and raises a flag that the next synchronising call on the
 *     context reports as CLEDB_ERR_NODB (every host entry point, cledb_synchronize).
 *   - all arrays are row-major; float = IEEE binary32.
 */
#ifndef CLEDB_B200_H
#define CLEDB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cledb_ctx cledb_ctx;

#define CLEDB_OK            0
#define CLEDB_ERR_ARG      -1   /* bad argument (null pointer, size, unknown id, nsearch out of range)   */
#define CLEDB_ERR_CUDA     -2   /* a CUDA runtime call failed                                             */
#define CLEDB_ERR_NODB     -3   /* a pixel refers to a database id that was never put                     */
#define CLEDB_ERR_DBFORMAT -4   /* database violates the layout contract (component 0 != 1, non-finite)   */
#define CLEDB_ERR_NOMEM    -5   /* device or host allocation failed                                       */
#define CLEDB_ERR_COMM     -6   /* multi-GPU: NCCL missing / failed, or no communicator on the context    */

#define CLEDB_MODE_IQUV 0       /* cledb_matchiquv, CLEDB_PROC.py:921  */
#define CLEDB_MODE_IQUD 1       /* cledb_matchiqud, CLEDB_PROC.py:1094 */

#define CLEDB_ENC_F32 0         /* database kept as float32 in HBM (lossless)                             */
#define CLEDB_ENC_I16 1         /* database kept as int16 codes in HBM (see cledb_codec_*)                */

/* Exact tile pruning (default on; cledb_set_pruning or the environment variable CLEDB_B200_PRUNING=0 turn it off).  A
 * database put while pruning is on is stored, inside each sign class, as the leaves of a k-d partition of its entries in the
 * space of the five chi^2 components: a 256-entry tile is a small box of that space, stored with its component bounds
 * (and the bounds of every group of 16 tiles).  The search orders the pixels by the tile that fits them best and skips,
 * per pixel, every tile whose chi^2 lower bound exceeds the pixel's running threshold.  The bound is exact in floating
 * point (fma and the accumulation chain are monotone), so idx / chi / kpos / rows are bit-identical to the plain search;
 * only the amount of arithmetic changes (about 2 % of the (pixel, tile) pairs are evaluated on BASELINE config 2).
 * Pruning off = the brute-force evaluation of every (pixel, entry) pair, the kernel the FP32 roofline is quoted on.
 * The fp64 validation kernel always searches by brute force. */
#define CLEDB_PREC_FP32 0       /* production kernel: fp32 FMA residuals                                  */
#define CLEDB_PREC_FP64 1       /* validation kernel: the reference's operation order in f32/f64          */

#define CLEDB_MAX_NSEARCH 32

/* per-pixel status written by cledb_match (why a pixel was or was not searched) */
#define CLEDB_PIX_SEARCHED   0
#define CLEDB_PIX_NULL       1  /* NaN / all-zero observation: CLEDB_PROC.py:944-948, 1118-1122          */
#define CLEDB_PIX_ALLNAN     2  /* every chi^2 is NaN in the reference (0/0, x/0 with weight 0, nf == 0)  */
#define CLEDB_PIX_REFRAISES  3  /* the reference raises for this pixel (empty candidate set / NaN max)   */
#define CLEDB_PIX_ALLINF     4  /* every chi^2 is +inf in the reference (zero sigma on a fitted component) */

/* ---- library / context ----------------------------------------------------------------------------- */
const char* cledb_version(void);
int         cledb_create(int device, cledb_ctx** ctx);
int         cledb_destroy(cledb_ctx* ctx);
const char* cledb_last_error(cledb_ctx* ctx);                 /* ctx may be NULL: last error of cledb_create */
/* props[0]=SM count, [1]=SM clock kHz, [2]=total HBM bytes, [3]=L2 bytes, [4]=cc major*10+minor,
 * [5]=resident search CTAs per SM, [6]=search kernel registers/thread, [7]=search kernel smem bytes/CTA */
int         cledb_device_props(cledb_ctx* ctx, int64_t props[8]);
int         cledb_synchronize(cledb_ctx* ctx);
/* number of kernels this library has launched on the context since creation (bench.py "gpu_launches") */
int64_t     cledb_launch_count(cledb_ctx* ctx);
/* CUDA-event timing of the search kernel alone: begin resets, every cledb_match[_dev] call records one
 * sample, collect synchronises and returns up to cap durations in ms (returns the number written). */
int         cledb_profile_begin(cledb_ctx* ctx);
int         cledb_profile_collect(cledb_ctx* ctx, float* ms, int cap);

/* FFMA-only microkernel: measured FP32 CUDA-core peak of this device in TFLOP/s (best of 5 after warm-up).
 * It is the roofline denominator of the search kernel, which is bound by the FP32 pipe, not by HBM. */
int         cledb_measure_fp32_peak(cledb_ctx* ctx, double* tflops);

/* ---- database residency ---------------------------------------------------------------------------- *
 * Replaces the per-pixel pickling of database[db_enc[xx,yy]] into every pool task (CLEDB_PROC.py:304-307):
 * a database is uploaded once, partitioned by sign(V1) (the prefilter of CLEDB_PROC.py:974), optionally
 * int16-encoded, and stays in HBM until dropped.  db is the [n,8] float32 array that sdb_preprocess produces
 * (CLEDB_PREPINV.py:306-317): component 0 must be exactly 1. */
int cledb_db_put(cledb_ctx* ctx, int db_id, const float* db, int64_t n, int encoding);
int cledb_db_drop(cledb_ctx* ctx, int db_id);
/* info[0]=n, [1]=rows with V1>0, [2]=V1<0, [3]=V1==0, [4]=V1 NaN (never candidates), [5]=encoding,
 * [6]=HBM bytes, [7]=entries per tile, [8..15]=power-of-two shift of components 0..7 (int16 encoding) */
int cledb_db_info(cledb_ctx* ctx, int db_id, int64_t info[16]);
/* the database exactly as the kernels see it after decoding, [n,8] float32 (rows with NaN V1 come back as
 * [1,0,0,NaN,0,0,0,0]); this is the array the oracle is given in parity tests */
int cledb_db_decoded(cledb_ctx* ctx, int db_id, float* out);
/* decoded rows by original index, rows_out[n,8]; idx < 0 gives zeros (database_sel[ixr,:], CLEDB_PROC.py:1064) */
int cledb_gather_rows(cledb_ctx* ctx, int db_id, const int32_t* idx, int64_t n, float* rows_out);

/* cledb_db_select: sdb_findelongationanddens (CLEDB_PREPINV.py:458-468) for every pixel of a map at once.
 *   table[nfiles,3] float32 = dbynumbers of sdb_fileingest (CLEDB_PREPINV.py:402-412): file index, 100*height,
 *   100*log10(density), in the order of the sorted file list.  For pixel p the files whose |height/100 - yobs[p]| is
 *   minimal are found, and among the range [first match, last match) - the reference's slice never reaches the last
 *   file of a height - the one whose |density/100 - dobs[p]| is smallest (first on ties); all in float32 as NumPy does.
 *   file_out[p] = table[.,0] of that file, or CLEDB_SELECT_NAN (yobs is NaN: the reference raises IndexError) or
 *   CLEDB_SELECT_EMPTY (a single file matches the height: the reference raises ValueError on the empty slice). */
#define CLEDB_SELECT_NAN   -1
#define CLEDB_SELECT_EMPTY -2
int cledb_db_select(cledb_ctx* ctx, int64_t npix, const float* yobs, const float* dobs, int nfiles, const float* table,
                    int32_t* file_out);
int cledb_db_select_dev(cledb_ctx* ctx, int64_t npix, const float* yobs, const float* dobs, int nfiles, const float* table,
                        int32_t* file_out, void* stream);

/* ---- the search ------------------------------------------------------------------------------------ *
 * Replaces, for every pixel at once, the candidate selection, chi^2 evaluation and partial sort of
 * cledb_matchiquv / cledb_matchiqud (CLEDB_PROC.py:944-1013 / 1118-1188): for pixel p the nsearch smallest
 * reduced chi^2 over the candidate rows of database db_id[p] (IQUV: rows whose sign(V1) equals sign(sobs[p,3]);
 * IQUD: all rows with non-NaN V1), ties to the lowest original row index.
 *   sobs[npix,8], snr[npix,8] float32, db_id[npix] int32
 *   idx_out [npix,nsearch]  original row index, -1 = empty slot
 *   chi_out [npix,nsearch]  reduced chi^2 (float32)             chi64_out: same in float64 (may be NULL)
 *   kpos_out[npix,nsearch]  position of the row inside the pixel's candidate list (needed to replay the
 *                           reference's cledb_partsort index quirk, CLEDB_PROC.py:1620-1625); may be NULL
 *   rows_out[npix,nsearch,8] decoded database rows of the winners (may be NULL)
 *   status_out[npix] uint8  CLEDB_PIX_*                         ncand_out[npix] int64 candidates evaluated (may be NULL)
 * Thresholds (maxchisq), field strength, geometry and issue flags are O(nsearch) work on these outputs and
 * stay in the host binding (CLEDB_PROC.py:1027-1082). */
int cledb_match(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                int32_t* idx_out, float* chi_out, double* chi64_out, int32_t* kpos_out,
                float* rows_out, uint8_t* status_out, int64_t* ncand_out);
int cledb_match_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                    const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                    int32_t* idx_out, float* chi_out, double* chi64_out, int32_t* kpos_out,
                    float* rows_out, uint8_t* status_out, int64_t* ncand_out, void* stream);
/* tuning knobs: pixels per warp work item (1..40; the pruned search uses at most 32 and ignores the chunk count) and how
 * many chunks each candidate list is cut into (1..8); 0 = choose from the pixel count (default).  Results do not depend on
 * them. */
int cledb_set_tuning(cledb_ctx* ctx, int pixels_per_item, int split);
/* search variant (results are bit-identical in all three):
 *   0  plain layout, brute force: every (pixel, entry) pair evaluated, entries in a stride-permuted order
 *   1  k-d tile layout with exact tile pruning (default)
 *   2  k-d tile layout, seeded brute force: every (pixel, entry) pair is evaluated - the FP32-roofline kernel - but each
 *      pixel starts from the exact top-nsearch of its seed tile, so its threshold is near final from the first tile on and
 *      the exact re-evaluation of flagged (pixel, tile) pairs drops from ~54 to ~10 per pixel
 * The layout is fixed when a database is put (0: plain, 1 or 2: k-d tiles); every resident database must have the same
 * layout when a search runs; 1 and 2 can be switched between searches without putting the databases again. */
int cledb_set_pruning(cledb_ctx* ctx, int on);
/* after a pruned search: evaluated[0] = (pixel, tile) pairs evaluated, evaluated[1] = pairs skipped by the bound; counts of the
 * last search launched (cledb_invert sends maps of 256 Ki pixels and more through in 128 Ki-pixel chunks: the last chunk) */
int cledb_pruning_stats(cledb_ctx* ctx, int64_t evaluated[2]);

/* ---- search + solution building in one call --------------------------------------------------------- *
 * Replaces everything cledb_invproc does per pixel (CLEDB_PROC.py:304-319 with the bodies of cledb_matchiquv /
 * cledb_matchiqud, :944-1087 / :1118-1252): the search above, then - on the device - the replay of the
 * cledb_partsort index quirk (:1620-1625, unless strict_topk), the maxchisq cut-offs (:1027-1044 / :1202-1219),
 * the field strength (:1049-1059, :1225), cledb_phys/cledb_physcle/cledb_params (:1634-1749) and cledb_quderotate
 * (:1575-1588).  Only PyCELP-type databases (dbtype 1, density taken from the observation) are built on the
 * device; CLE-type headers return CLEDB_ERR_ARG and the binding builds those solutions from cledb_match.
 *
 * Transcendentals are NOT evaluated on the device, so that the numbers are the caller's own libm/NumPy bits:
 *   grid->sin_bphi/cos_bphi[nbphi], sin_btheta/cos_btheta[nbtheta]: sin/cos of the float32 grid angles
 *   rot_cos[npix], rot_sin[npix]: cos(-2*aobs), sin(-2*aobs) of every pixel (float32)
 *   yobs[npix], dobs[npix] float32; dopp: B_POS of every pixel for IQUD (element p at dopp[p*dopp_stride],
 *   float32 or float64 as dopp_is_f64 says; NULL for IQUV); maxchisq, bcalc as in ctrlparams
 *   invout[npix,nsearch,11], sfound[npix,nsearch,8] float32; status_out[npix] (may be NULL)
 *   issue_out[npix] int32 (may be NULL): the issue codes of cledb_invproc (:323-355) this pixel raises, OR-ed together
 *     512  = Van-Vleck proximity of the LAST solution (53 deg <= vartheta_B <= 56 deg; `flg = 0` sits inside the reference's
 *            loop over solutions, :324, so only slot nsearch-1 decides), evaluated in float64
 *     1024 = chi^2 of the best solution > 10 (:351)          2048 = snr[p,0] < 1 (:353)
 *     1    = vartheta_B lies within 2e-3 deg of 53 or 56: the reference evaluates the chain in float32 with the caller's
 *            NumPy, whose last bits decide such a pixel - the binding re-evaluates those (about 1 pixel in 50 000)
 *   The binding ORs (issue_out[p] & ~1) into both lines of its issuemask: a code already set is not added twice. */
typedef struct cledb_dbgrid {
    int32_t dbtype, ngx, nbphi, nbtheta;               /* dbhdr[14], [2], [3], [4]  (CLEDB_PREPINV.py:482-508) */
    float   gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax;
    const float* sin_bphi;   const float* cos_bphi;    /* [nbphi]   */
    const float* sin_btheta; const float* cos_btheta;  /* [nbtheta] */
} cledb_dbgrid;
int cledb_invert(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                 const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                 const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin,
                 const void* dopp, int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid,
                 double maxchisq, int bcalc, int strict_topk,
                 float* invout, float* sfound, uint8_t* status_out, int32_t* issue_out);
/* device-pointer variant: every array, including the four tables inside *grid, lives on the context's device */
int cledb_invert_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                     const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                     const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin,
                     const void* dopp, int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid,
                     double maxchisq, int bcalc, int strict_topk,
                     float* invout, float* sfound, uint8_t* status_out, int32_t* issue_out, void* stream);

/* ---- reduced-database mode (reduced = True) -------------------------------------------------------- *
 * Replaces cledb_getsubsetiquv / cledb_getsubsetiqud and the search over their per-pixel output (CLEDB_PROC.py:1401-1568,
 * 951-963, 1125-1134): for every phi index only the nsearch theta rows whose tan(theta)*sin(phi) is closest to the
 * observed tan(Phi_B), and the nsearch closest to tan(pi - Phi_B), are searched (for every gx; duplicates and the all-zero
 * rows of the phi = 0 plane included, in the reference's order).  The subset is enumerated on the device, never built.
 *   outer      = ngx (PyCELP type) or ned*ngx (CLE type); every database used must have outer*nbphi*nbtheta rows
 *   tan_btheta = tan(bthetamin + l*(bthetamax-bthetamin)/nbtheta), sin_bphi = sin(bphimin + k*(bphimax-bphimin)/nbphi),
 *                float64, evaluated by the caller (:1417-1422)
 *   tie_rank   = rank of theta index l in the caller's argsort of an all-equal array: the reference's presort uses
 *                NumPy's default (unstable) argsort and the phi = 0 plane is one big tie
 *   tan_phib   = [npix,2] float64: tan(Phi_B) and tan(pi - Phi_B) of every pixel (:1428-1430 / :1516-1518)
 * Outputs as cledb_match, with idx the row of the FULL database (:1070-1071) and kpos the position inside the pixel's
 * filtered subset; nan_slots_out[npix] = how many leading output slots the reference leaves empty because that many
 * candidates have a NaN chi^2 (its selection sort picks NaNs first). */
typedef struct cledb_reduced {
    int32_t outer, nbphi, nbtheta;
    const double*  tan_btheta;   /* [nbtheta] */
    const double*  sin_bphi;     /* [nbphi]   */
    const int32_t* tie_rank;     /* [nbtheta] */
    const double*  tan_phib;     /* [npix,2]  */
} cledb_reduced;
int cledb_match_reduced(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                        const float* sobs, const float* snr, const int32_t* db_id, int nsearch, const cledb_reduced* red,
                        int32_t* idx_out, float* chi_out, double* chi64_out, int32_t* kpos_out, float* rows_out,
                        uint8_t* status_out, int32_t* nan_slots_out);
/* device-pointer variant: every array, the four tables inside *red included, lives on the context's device; enqueued on the
 * stream (NULL = the context's own).  Every output is required (chi64_out only for the fp64 search). */
int cledb_match_reduced_dev(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                            const float* sobs, const float* snr, const int32_t* db_id, int nsearch, const cledb_reduced* red,
                            int32_t* idx_out, float* chi_out, double* chi64_out, int32_t* kpos_out, float* rows_out,
                            uint8_t* status_out, int32_t* nan_slots_out, void* stream);
/* the same followed by the solution building of cledb_invert */
int cledb_invert_reduced(cledb_ctx* ctx, int mode, int precision, int64_t npix,
                         const float* sobs, const float* snr, const int32_t* db_id, int nsearch, const cledb_reduced* red,
                         const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin,
                         const void* dopp, int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid,
                         double maxchisq, int bcalc, int strict_topk,
                         float* invout, float* sfound, uint8_t* status_out, int32_t* issue_out);

/* ---- multi-GPU: one process and one context per GPU, pixels sharded, databases replicated ------------ *
 * The reference dispatches one pool task per pixel (CLEDB_PROC.py:304-313): pixels are independent, the search has no exchange
 * step, and the only collective is the gather of the finished maps.  NCCL is bound at run time (the libnccl.so.2 already
 * loaded into the process, else the system's; CLEDB_B200_NCCL overrides), single-GPU use needs none.
 *   cledb_comm_unique_id   rank 0 creates the 128-byte id; the launcher ships it to the other ranks (torch.distributed
 *                          broadcast, MPI, a file ...)
 *   cledb_comm_init        collective over all ranks; nranks == 1 needs no id and no NCCL.  cledb_comm_destroy is collective too
 *                          (the ranks unmap each other's memory before anyone frees it)
 *   cledb_comm_info        info[0]=rank, [1]=nranks, [2]=NCCL version code, [3]=1 if an NCCL communicator is live
 *   cledb_comm_transport   how the gathers move data.  1 = PEER MEMORY (default where it works): cledb_comm_init maps a small flag
 *                          array of every rank into every process with CUDA IPC, cledb_comm_window does the same for the
 *                          receive window, and the gather is the library's own kernel storing each record once, straight at
 *                          its place in every rank's window over NVLink / NVSwitch, bracketed by two rounds of flags
 *                          (free / done epochs, st.release.sys / ld.acquire.sys) - no staging copy, no scatter pass.
 *                          0 = ncclAllGather into a receive buffer + k_unshard.  want = -1 queries; *active receives the
 *                          transport in effect.  Asking for 1 where IPC is unavailable returns CLEDB_ERR_COMM and the reason.
 *                          Every rank must use the same transport (CLEDB_B200_PEER=0 in the environment starts with NCCL;
 *                          CLEDB_B200_NO_IPC=1 makes the IPC export fail, which exercises the fall-back of all ranks to NCCL).
 *                          One gather in flight per communicator: the flag epochs count the operations in the order they are
 *                          enqueued, which must be the order they run in - enqueue gathers on one stream, or order the streams
 *                          with events (cledb_invert_gather_dev does that for its second stream).
 *   cledb_comm_window      collective: the receive window of this rank holds at least `bytes` bytes (device memory, 256-byte
 *                          aligned; re-allocated and re-mapped on every rank when it has to grow - earlier pointers and
 *                          contents are then gone).  One window per communicator; cledb_invert_sharded keeps its full maps at
 *                          the start of the same window.
 *   cledb_comm_push_blocks peer transport: bound of the grid of the push kernels of the following cledb_allgather_dev /
 *                          cledb_gather_maps_dev calls (0 = no bound, the default).  A gather enqueued on a second stream beside
 *                          a search should leave the SMs to the search: 24 blocks (of 1024 threads) reach the NVLink rate.
 *   cledb_allgather_dev    every rank contributes `bytes` bytes at `send`; `recv` receives nranks*bytes in rank order
 *                          (device pointers, enqueued on the stream).  `recv` inside the window + peer transport (bytes and
 *                          pointers multiples of 16): every rank stores its block into every rank's window (k_push_block);
 *                          otherwise ncclAllGather over NVLink / NVSwitch
 *   cledb_partition        pixels ordered by key[p] in 0..nkeys-1 (stable; the binding passes db_enc) and cut into nranks
 *                          contiguous ranges of equal total weight (weight_of_key[key]: rows of that database, NULL = 1):
 *                          order_out[npix], bounds_out[nranks+1]; rank r owns order_out[bounds[r] : bounds[r+1]], so it
 *                          needs only the databases of its range resident.  Host arrays, pure host code, the same result
 *                          on every rank.
 *   cledb_shard_layout     byte offsets of the packed records of a shard of <= maxcount pixels:
 *                          layout[0..3] = offsets of invout[maxcount,k,11], sfound[maxcount,k,8], status[maxcount],
 *                          issue[maxcount] int32; layout[4] = bytes per rank (the all-gather count)
 *   cledb_unshard_dev      scatters the gathered records of all ranks to their pixels: recv[nranks][layout] ->
 *                          invout[npix,k,11], sfound[npix,k,8], status[npix], issue[npix] (device pointers; status and
 *                          issue may be NULL)
 *   cledb_map_layout       byte offsets of the full maps of npix pixels inside the window: layout[0..3] = invout[npix,k,11],
 *                          sfound[npix,k,8], status[npix], issue[npix] int32; layout[4] = bytes
 *   cledb_gather_maps_dev  peer transport: the all-gather and the scatter in ONE kernel (k_push_records).  This rank's packed
 *                          records (`send`, cledb_shard_layout of maxcount) of its nlocal pixels my_order_dev[nlocal] (their
 *                          flat indices in the map) are stored at their pixels' places in the maps block `maps` (a pointer
 *                          into the local window, the same offset on every rank) of every rank (root < 0) or of `root`
 *                          only.  When the call's work has run on the stream, every record of every rank is in place.
 *   cledb_invert_gather_dev  peer transport: search + solutions + gather of this rank's shard as a PIPELINE (device pointers as in
 *                          cledb_invert_dev, grid tables on the device).  The shard goes through in `pieces` contiguous pieces
 *                          (0 = chosen from npix and the rank count: 2 from six ranks on when a piece still holds 64 Ki pixels, else 1; the same
 *                          on every rank - every rank runs `pieces` pushes, also one with no pixels); the records of piece i
 *                          cross NVLink - a k_push_records of a few blocks on the communicator's second stream - while piece
 *                          i+1 is searched on the caller's stream; the caller's stream then waits for the last push.
 *   cledb_invert_sharded   cledb_invert for one map of npix pixels over all ranks: every rank passes the inputs of ITS
 *                          nlocal pixels (those of order[bounds[rank] : bounds[rank+1]], in that order; db_id = the ids
 *                          under which this rank put their databases), searches and builds them, writes the packed
 *                          records into the send buffer; then cledb_invert_gather_dev (peer transport) or ONE ncclAllGather +
 *                          k_unshard; ranks with root < 0 or root == rank receive the full maps invout[npix,k,11] ... in
 *                          their host buffers (the others may pass NULL outputs).  Collective: every rank must call it. */
#define CLEDB_COMM_ID_BYTES 128
int cledb_comm_unique_id(void* id128);
int cledb_comm_init(cledb_ctx* ctx, int rank, int nranks, const void* id128);
int cledb_comm_destroy(cledb_ctx* ctx);
int cledb_comm_info(cledb_ctx* ctx, int32_t info[4]);
int cledb_comm_transport(cledb_ctx* ctx, int want, int32_t* active);
int cledb_comm_push_blocks(cledb_ctx* ctx, int blocks);
int cledb_comm_window(cledb_ctx* ctx, int64_t bytes, void** ptr_out);
int cledb_allgather_dev(cledb_ctx* ctx, const void* send, void* recv, int64_t bytes, void* stream);
int cledb_partition(int64_t npix, const int32_t* key, int nkeys, const double* weight_of_key, int nranks,
                    int32_t* order_out, int64_t* bounds_out);
int cledb_shard_layout(int64_t maxcount, int nsearch, int64_t layout[5]);
int cledb_unshard_dev(cledb_ctx* ctx, int nranks, const int64_t* bounds_dev, const int32_t* order_dev, int64_t npix,
                      int64_t maxcount, int nsearch, const void* recv, float* invout, float* sfound, uint8_t* status,
                      int32_t* issue, void* stream);
int cledb_map_layout(int64_t npix, int nsearch, int64_t layout[5]);
int cledb_gather_maps_dev(cledb_ctx* ctx, int64_t npix, int64_t nlocal, const int32_t* my_order_dev, int64_t maxcount,
                          int nsearch, const void* send, void* maps, int root, void* stream);
int cledb_invert_gather_dev(cledb_ctx* ctx, int mode, int precision, int64_t nlocal,
                            const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                            const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin,
                            const void* dopp, int dopp_is_f64, const cledb_dbgrid* grid_dev,
                            double maxchisq, int bcalc, int strict_topk,
                            int64_t npix, const int32_t* my_order_dev, void* maps, int root, int pieces, void* stream);
int cledb_invert_sharded(cledb_ctx* ctx, int mode, int precision, int64_t nlocal,
                         const float* sobs, const float* snr, const int32_t* db_id, int nsearch,
                         const float* yobs, const float* dobs, const float* rot_cos, const float* rot_sin,
                         const void* dopp, int dopp_is_f64, int64_t dopp_stride, const cledb_dbgrid* grid,
                         double maxchisq, int bcalc, int strict_topk,
                         int64_t npix, const int32_t* order, const int64_t* bounds, int root,
                         float* invout, float* sfound, uint8_t* status_out, int32_t* issue_out);

/* ---- pinned host memory for the binding's staging buffers (cudaHostAlloc / cudaFreeHost) ------------- */
int cledb_host_alloc(int64_t bytes, void** out);
int cledb_host_free(void* p);

/* ---- elementwise neighbours ------------------------------------------------------------------------ *
 * cledb_blos: blos_proc_1pix (CLEDB_PROC.py:874-914) for npix pixels of one line.
 *   iquv[npix,4]; kfac = 1e9*h*c/mu_B/lambda_ref^2, g_eff, ffac = F_factor as Python floats (they are
 *   rounded to float32 exactly where NumPy's weak-scalar promotion rounds them)
 *   out[npix,4] = [B_deg+, B_deg-, B_classic, azimuth]; flags[npix,2] = issue codes 32 / 64 raised */
int cledb_blos(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac,
               float* out, uint8_t* flags);
int cledb_blos_dev(cledb_ctx* ctx, int64_t npix, const float* iquv, double kfac, double g_eff, double ffac,
                   float* out, uint8_t* flags, void* stream);
/* cledb_qurotate: obs_qurotate + obs_calcheight (CLEDB_PREPINV.py:1277-1350, 897-926).
 *   xs[nx], ys[ny]: the per-row / per-column coordinates crval'+i*cdelt evaluated by the caller in the
 *   scalar types the header carries (float64), or hpxy[2,nx,ny] float32 (Cryo-NIRSP grid) when non-NULL;
 *   in[nx,ny,8] -> out[nx,ny,8] (Q,U of both lines rotated by 2*aobs), aobs[nx,ny], yobs[nx,ny] */
int cledb_qurotate(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref,
                   const float* hpxy, const float* in, float* out, float* aobs, float* yobs);
int cledb_qurotate_dev(cledb_ctx* ctx, int nx, int ny, const double* xs, const double* ys, double linpolref,
                       const float* hpxy, const float* in, float* out, float* aobs, float* yobs, void* stream);

/* cledb_obs_integrate: obs_integrate_work_1pix (CLEDB_PREPINV.py:1164-1270) for npix spectra of one line: continuum
 * subtraction, wavelength integration of I, Q, U, Stokes V from the median of V / (d(I/Itot)/dlambda), signal-to-noise
 * ratio and the SNR issue codes 1 / 2.  float64 arithmetic in NumPy's order, including its pairwise summation, so the
 * outputs are bit-identical to the reference's for float64 input spectra.
 *   spec[npix,nw,4] float64;  sp: what the caller derives from the wavelength axis with its own NumPy
 *     wcont  index of the clean-continuum window (:968-976);  lo, hi = int(0.25*nw), int(0.75*nw) (:1210)
 *     dwv    mean dispersion (:1046);  the coefficients of numpy.gradient on that axis: uniform != 0 -> h2 = 2*dx,
 *            else ga/gb/gc[nw-2] (interior points 1..nw-2); dx_first, dx_last for the two edges
 *   tot[npix,4], snr[npix,4], background[npix,4] float32 (cast as the reference's output arrays do), issue[npix] int32;
 *   pixels without counts or with a non-finite sample give zeros (:1168, :1267); status[npix]: 0 ok/empty, 1 = the
 *   reference raises for this pixel (no sample inside the 0.3 % .. 99.7 % window of the running sum). */
typedef struct cledb_spectral {
    int32_t nw, wcont, lo, hi, uniform;
    double  dwv, h2, dx_first, dx_last;
    const double* ga; const double* gb; const double* gc;
} cledb_spectral;
int cledb_obs_integrate(cledb_ctx* ctx, int64_t npix, const double* spec, const cledb_spectral* sp,
                        float* tot, float* snr, float* background, int32_t* issue, uint8_t* status);

/* cledb_obs_dens: obs_dens_work_1pix (CLEDB_PREPINV.py:1436-1479) for npix pixels: density from the Fe XIII 1074/1079
 * intensity ratio through the CHIANTI look-up table.  The reference builds a quadratic interpolating spline of
 * (ratio -> density) for the first table height above the pixel, per pixel; here the caller builds the nh splines once
 * (knots[nh,nknots], coefs[nh,ncoef], degree 2, scipy.interpolate.interp1d(kind="quadratic")._spline) and the device
 * evaluates them with de Boor's recurrence in scipy's operation order (extrapolating outside the table).
 *   a_obs, b_obs = Stokes I of the two lines, yobs = height, float32;  heights[nh] float64
 *   dens_out[npix] float32 = density as the reference stores it before its log10 (1 for b_obs == 0, 0 for a NaN/inf
 *   ratio), issue_out[npix] = 4 where no table height lies above the pixel (the last row is used). */
int cledb_obs_dens(cledb_ctx* ctx, int64_t npix, const float* a_obs, const float* b_obs, const float* yobs,
                   int nh, const double* heights, int nknots, const double* knots, int ncoef, const double* coefs,
                   float* dens_out, int32_t* issue_out);

/* ---- int16 database codec (pure host code, usable without a GPU) ------------------------------------ *
 * code = IEEE binary16 bit pattern of x * 2^shift (round to nearest even), stored as int16;
 * decode = float(binary16) * 2^-shift, exact in float32.  cledb_codec_shift picks the shift that puts the
 * largest finite |x[i*stride]| just below 2^15.  The legacy log-scale int16 scheme of the reference
 * (sdb_dcompress, CLEDB_PREPINV.py:1618-1644) is disabled upstream; this codec is defined by this library. */
int  cledb_codec_shift(const float* x, int64_t n, int64_t stride);
void cledb_codec_encode(const float* x, int64_t n, int shift, int16_t* codes);
void cledb_codec_decode(const int16_t* codes, int64_t n, int shift, float* x);

#ifdef __cplusplus
}
#endif
#endif /* CLEDB_B200_H */
