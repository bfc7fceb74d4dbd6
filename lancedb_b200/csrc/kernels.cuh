// kernels.cuh -- kernel argument blocks and host launchers shared by the C-ABI layer.
#pragma once

#include "common.cuh"

namespace lgpu {

// ---------------- scan (K2+K3) geometry ------------------------------------------
constexpr int SCAN_G = 8;                         // queries per tile
constexpr uint32_t SCAN_ROWS_TILE_MID = 4 * 32 * 12;   // exact kernel (scan2.cu): 2 x 128 scanner threads x 12 rows
constexpr uint32_t SCAN3_ROWS_TILE = 8 * 32 * 6;       // filter kernel (scan3.cu): 256 scanner threads x 6 rows
constexpr int SCAN_LUT_HALF = 32768;              // exact kernel: [256 c][8 s][4 g] f32
constexpr int SCAN_LUT_BYTES = 2 * SCAN_LUT_HALF; // [2 h] halves

__host__ __device__ __forceinline__ uint32_t scan_nrb(uint32_t n, uint32_t rows_tile)
{
    return (n + rows_tile - 1) / rows_tile;
}
__host__ __device__ __forceinline__ uint32_t scan_rb_rows(uint32_t n, uint32_t nrb)
{
    return nrb ? (((n + nrb - 1) / nrb + 31u) & ~31u) : 0u;
}

// One scan tile, precomputed by the regroup step (group.cu) so the persistent scan CTAs fetch a tile with a
// single 128-byte read: partition p, rows [row0, row0+nrows), ng (1..8) queries q[] (probe slots slot[]) whose
// distances go to dist_out + out[g].  ng == 0 marks "no tile" inside the kernel's shared-memory copy.
struct alignas(16) TileDesc {
    uint32_t p, row0, nrows, ng;
    uint32_t q[SCAN_G];
    uint32_t slot[SCAN_G];        // probe slot (q * nprobes + j) of each query
    uint32_t out[SCAN_G];         // offset (floats) of the slot's distance segment; sub-batches keep it < 2^32
    // the partition's constants, copied in by the regroup step: read from the shared-memory copy of the descriptor, a
    // tile's first code-word load and its row-constant load go out at once instead of behind a dependent global read
    uint32_t n_p, npad;           // rows / padded rows of partition p
    uint32_t code_base8;          // code_base[p] / 8
    uint32_t part_off32;          // part_off[p] (row position of the partition's first row)
};
static_assert(sizeof(TileDesc) == 128, "TileDesc layout");

// one surviving row of the filter scan's candidate mode
struct alignas(16) CandRec {
    float lb;                     // lower bound L of the row's distance
    uint32_t p, row;              // partition and row inside it
    uint32_t pad;
};
constexpr uint32_t CAND_NO_THR = 0xffffffffu;
constexpr uint32_t CAND_TOPK_MAX = 128;      // largest k (or k * refine_factor) the candidate mode serves (LGPU_CAND_KMAX lowers it)
constexpr uint32_t CAND_CAP_MAX = 2048;      // largest per-query candidate list

struct ScanArgs {
    // index (device)
    const float *centroids;       // [nlist][dim]
    const float *cb_tiled;        // [nch][256][8][dsub]
    const unsigned char *codes;   // skewed code streams, see retile.cu
    const uint64_t *code_base;    // [nlist] byte offset of partition p's stream block
    const uint32_t *part_n;       // [nlist]
    const uint32_t *part_npad;    // [nlist] rows rounded up to 32
    uint32_t dim, m, nch, metric, nlist;
    uint32_t rows_tile;           // SCAN_ROWS_TILE_MID (exact) or SCAN3_ROWS_TILE (filter)
    unsigned long long fzero2;    // packed (+0.f, +0.f); opaque to ptxas on purpose (scan_common.cuh)
    // batch (device)
    const float *queries;         // [B][dim] (normalised for cosine)
    const uint32_t *total_tiles;  // [1]
    uint32_t *tile_counter;       // [1], zeroed before launch
    float *dist_out;
    const TileDesc *tile_desc;    // [total_tiles]
    const uint32_t *gate;         // optional [1] (exact kernel as fix-up pass): 0 = nothing to do
    // filter pass (scan3.cu, tables.cu); unused by the exact kernel
    const uint4 *qt;              // [B][nch][256] x (8 x u16): quantised per-query tables, rotated (tables.cu)
    const float *qt_step;         // [B] quantisation step
    const float *qt_base;         // [B] sum_i min_i (- (m - 1) for dot)
    const float *probe_A;         // [B*nprobes] |q - c_p|^2 - |q|^2   (nullptr for dot)
    const float *row_R;           // [nrows] 2 * codeword(row) . c_p  (nullptr for dot)
    const uint64_t *part_off;     // [nlist+1]
    // filter pass, candidate mode (cand != nullptr): instead of one f32 per (row, query) in dist_out, the scanners
    // keep a running per-query threshold and append only the rows that can still be among the exact top-k
    uint32_t nprobes, topk;       // probe slots per query; k (the number of neighbours the threshold is for)
    uint32_t *thr;                // [B] ordered-uint key (f32_key) of tau_q, CAND_NO_THR = none yet; atomicMin
    const float *slack;           // [B] scale (W + 2E): a row survives iff L <= tau_q + slack_q
    uint32_t *cand_cnt;           // [B] appended candidates (may exceed cand_cap: the overflow is dropped and flagged)
    CandRec *cand;                // [B][cand_cap]
    uint32_t cand_cap;
    uint32_t *cand_key;           // [B][cand_cap] f32_key(lb) of every appended record, 0xffffffff where none is yet: what
                                  // the scanners re-read to tighten tau_q from the list itself (scan3.cu)
    uint32_t *cand_last;          // [B] list length at the query's last list-based tightening
};
bool scan_dsub_supported(uint32_t dsub);
// exact kernel (scan2.cu): residual -> f32 table chunk -> sequential code scan, bit-identical to the oracle;
// needs a.tile_desc built with rows_tile == SCAN_ROWS_TILE_MID
void launch_scan2(const ScanArgs &a, uint32_t dsub, int grid, cudaStream_t st);
// filter kernel (scan3.cu): lower bounds from 16-bit per-query tables; rows_tile == SCAN3_ROWS_TILE
void launch_scan3(const ScanArgs &a, int grid, cudaStream_t st);

// ---------------- tiny batches: one CTA per (query, probed partition) pair (small.cu) ----------------
struct SmallScanArgs {
    const float *centroids; const float *cb_tiled; const unsigned char *codes; const uint64_t *code_base;
    const uint32_t *part_n; const uint32_t *part_npad;
    uint32_t dim, m, nch, metric, nlist, nprobes;
    const float *queries;         // [B][dim] (normalised for cosine)
    const uint64_t *probes;       // [B*nprobes]
    uint64_t seg_stride;          // floats per probe slot in dist_out (>= the largest partition, multiple of 4)
    uint64_t *seg_off;            // [B*nprobes], written here: slot * seg_stride
    float *dist_out;
};
size_t small_scan_smem(uint32_t m, uint32_t dim);
void launch_small_scan(const SmallScanArgs &a, uint32_t dsub, uint32_t slots, cudaStream_t st);

// ---------------- batch preparation (grouping probes by partition) ----------------
struct GroupArgs {
    const uint64_t *probes;       // [B*nprobes] partition ids (u64 from the selector)
    uint32_t B, nprobes, nlist, rows_tile;
    const uint32_t *part_n;
    const uint32_t *part_npad;    // (tile descriptors) padded rows per partition
    const uint64_t *code_base;    // (tile descriptors) byte offset of each partition's code blocks, multiple of 8
    const uint64_t *part_off;     // (tile descriptors) row position of each partition's first row, < 2^32
    uint32_t *part_cnt;           // [nlist] zeroed before
    uint32_t *slot_pos;           // [B*nprobes]
    uint64_t *seg_local;          // [B*nprobes] offset inside the query's block
    uint64_t *qtot;               // [B]
    uint64_t *seg_off;            // [B*nprobes]
    uint32_t *qlist_off;          // [nlist]
    uint32_t *tile_off;           // [nlist+1] first tile of each partition's first query group (class A)
    uint32_t *tile_off_b;         // [nlist+1] first tile of each partition's remaining groups (class B, after all of A)
    uint32_t *qlist;              // [B*nprobes]
    uint32_t *total_tiles;        // [1]
    uint32_t *tile_counter;       // [1]
    unsigned long long *scanned_rows;  // [1] sum over probe slots of n_p (roofline bytes / m)
    const uint32_t *only;         // optional [B]: regroup only the flagged queries (fix-up pass)
    const uint32_t *gate;         // optional [1]: 0 = no query is flagged, the whole fix-up pass returns at once
    TileDesc *tile_desc;          // optional [max_tiles] tile descriptors for the streaming scan kernel
    uint32_t max_tiles;           // capacity of tile_desc (host bound on the tile count)
};
void launch_group(const GroupArgs &a, cudaStream_t st);

// ---------------- exact distance matrix (K1 coarse, flat v1) ----------------------
// D[q][c] for q < B, c < N.  mode 0: L2 (squared), 1: dot distance 1 - x.y,
// 2: cosine 1 - x.y/|x|/|y| (xnorm[B] = |x|, ysqrt[N] = |y| required).
// `only` (optional, [B]): rows whose flag is 0 are skipped (used by the tensor-core paths' fix-up pass)
void launch_dist_matrix(const float *Q, const float *C, uint32_t B, uint64_t N, uint32_t d, int mode,
                        const float *xnorm, const float *ysqrt, float *D, uint64_t ldD, cudaStream_t st,
                        const uint32_t *only = nullptr, const uint32_t *gate = nullptr);
// coarse step after the tensor-core GEMM, one kernel: from S[B][ld] = |x|^2 - 2 bf16(q).bf16(x) (N columns) to the k
// columns with the smallest EXACT l2 distance (lance lane order), ascending by (distance, column); flags[q] = 1 when
// the candidate band overflowed and the caller must redo the query with the exact kernels
void launch_coarse_finish(const float *S, uint64_t ld, uint32_t B, uint32_t N, const float *Q, const float *C,
                          const float *qn2, float xmax, uint32_t d, uint32_t k, uint64_t *out_ids, float *out_dist,
                          uint32_t *out_cnt, uint32_t *flags, uint32_t *gate, cudaStream_t st,
                          const uint64_t *list_pos = nullptr, const uint32_t *list_cnt = nullptr);
// list mode (list_pos / list_cnt given, N == ld == the list capacity): S[q] holds the list_cnt[q] scores the GEMM's
// filtering epilogue admitted for query q and list_pos[q] their columns; a list longer than its capacity flags the query
// row norms |x| = sqrt(dot(x,x)) in lance order; out[n]
void launch_row_norms(const float *X, uint64_t n, uint32_t d, float *out, cudaStream_t st);
// out[q] = x[q] / |x[q]|
void launch_normalize(const float *X, uint32_t B, uint32_t d, float *out, cudaStream_t st);
// exact distances of (query, stored row) pairs: out[q][c] = dist(Q[q], V[pos[q][c]]),
// pos == UINT64_MAX => +inf
void launch_pair_distance(const float *Q, const float *V, const uint64_t *pos, uint32_t B, uint32_t nc,
                          uint32_t d, int metric, float *out, cudaStream_t st);

// ---------------- top-k select (K4) ------------------------------------------------
constexpr uint32_t SELECT_KMAX = 2048;
// one top-k entry as it crosses NVLink in the partition-sharded search: a single ncclAllGather of
// [B][k] of these per rank, consumed in place by the merge (select mode 2)
struct alignas(16) TopkRecord {
    uint64_t id;      // _rowid, UINT64_MAX = unused slot
    float dist;       // _distance
    uint32_t pad;
};
static_assert(sizeof(TopkRecord) == 16, "TopkRecord layout");
struct SelectArgs {
    int mode;                     // 0: IVF distance segments, 1: dense row (ids = column / col_ids),
                                  // 2: strided candidate lists with per-entry ids
    // mode 0
    const float *dist;            // dist_out
    const uint64_t *seg_off;      // [B*nprobes]
    const uint64_t *probes;       // [B*nprobes]
    uint32_t nprobes, nlist;      // probe ids >= nlist are unused slots
    const uint32_t *part_n;
    const uint64_t *part_off;     // [nlist+1] storage row offsets
    const uint64_t *row_ids;      // [nrows]
    // mode 1 / 2
    const float *dense;           // values
    uint64_t ncols;               // candidates per query
    uint64_t row_stride;          // mode 1: ld; mode 2: stride between queries
    uint64_t inner, outer_stride; // mode 2: col -> (col/inner)*outer_stride + q*row_stride + col%inner
    const uint64_t *col_ids;      // mode 1: optional id per column (NULL => column index)
    const uint64_t *cand_ids;     // mode 2: id per entry (same addressing as dense)
    const uint64_t *cand_pos;     // mode 2: optional pos per entry
    const uint32_t *ncols_q;      // mode 2: optional [B] candidates of each query (<= ncols)
    const TopkRecord *cand_rec;   // mode 2: optional packed (id, dist) entries instead of dense + cand_ids
    // common
    uint32_t B, k;
    int has_lower, has_upper;
    float lower, upper;
    uint64_t *out_ids;            // [B][k]
    float *out_dist;              // [B][k]
    uint32_t *out_count;          // [B]
    TopkRecord *out_rec;          // optional [B][k]: packed output instead of out_ids / out_dist
    uint64_t *out_pos;            // optional [B][k] storage position (mode 0) / column (mode 1)
    const uint32_t *only;         // optional [B]: queries whose flag is 0 are left untouched
    const uint32_t *gate;         // optional [1]: 0 = no query is flagged (the kernel returns at once)
    // prefilter (query.rs:489-507): optional row-id allow-list bitmap; a candidate whose id has bit 0 (or is
    // >= allow_bits) is dropped before it can enter the top-k
    const uint32_t *allow;
    uint64_t allow_bits;
};
void launch_select(const SelectArgs &a, cudaStream_t st);
// flags[q] = cnt[q] < k (maximum_nprobes widening: the queries that did not find k rows)
void launch_count_below(const uint32_t *cnt, uint32_t B, uint32_t k, uint32_t *flags, cudaStream_t st);
// rec[i] = (ids[i], dist[i]) for i < n
void launch_pack_records(const uint64_t *ids, const float *dist, uint64_t n, TopkRecord *out, cudaStream_t st);

// ---------------- tensor-core shortlist (gemm.cu) ------------------------------------
bool gemm_shape_supported(uint32_t d);
// X f32 [n][d] -> bf16 [n][d]; norm2[n] = |x|^2 (f32) if norm2 != nullptr
void launch_to_bf16(const float *X, uint64_t n, uint32_t d, void *Xb, float *norm2, cudaStream_t st);
// out[q][x] = xnorm2[x] - 2 * bf16(Q[q]) . bf16(X[x])   (tcgen05 + TMA), q < B, x < N
// optional filtering epilogue: instead of writing the dense score matrix, append the columns whose score
// is <= thr[q] to the query's candidate list (count may exceed cap: those appends are dropped)
struct GemmFilter {
    const float *thr;             // [B]; nullptr = dense output
    uint32_t *count;              // [B], zeroed by the caller
    uint64_t *cand_pos;           // [B][cap] column (storage row) of each candidate
    uint64_t *cand_ids;           // [B][cap] its id (col_ids[x] or x)
    const uint64_t *col_ids;      // optional
    uint32_t cap;
    float *cand_s;                // optional [B][cap]: the admitted score itself (coarse step: second-level threshold)
};
void launch_gemm_dist(const void *Qb, const void *Xb, const float *xnorm2, uint32_t B, uint64_t N, uint32_t d,
                      float *out, uint64_t ld_out, int num_sms, cudaStream_t st, const GemmFilter *filter = nullptr);
// filtering epilogue on an existing dense score matrix D[B][ld] (see gemm.cu)
void launch_filter_dense(const float *D, uint64_t ld, uint32_t B, uint64_t N, const GemmFilter &flt, cudaStream_t st);
void launch_sample_threshold(const float *approx, const uint32_t *cnt, const float *qnorm2, float xmax, uint32_t d,
                             uint32_t B, uint32_t k, float *thr, cudaStream_t st);
// the same threshold straight from dense sample scores D[B][ld] (ns <= 4096 columns; false = shape not handled)
bool launch_sample_kth_threshold(const float *D, uint64_t ld, uint32_t ns, const float *qnorm2, float xmax, uint32_t d,
                                 uint32_t B, uint32_t k, float *thr, cudaStream_t st);
void launch_overflow_flags(const uint32_t *count, uint32_t cap, uint32_t B, uint32_t *flags, cudaStream_t st);
// flags[q] = 1 when the approximate shortlist of query q cannot be proven to contain the exact top-k:
// approx[q][0..kp) ascending, cnt[q] entries valid; proven iff cnt < kp or approx[kp-1] > approx[k-1] + 2E_q,
// E_q = 2^-7 (1+2^-8) |q| xmax + 4 d 2^-24 (|q| + xmax)^2
void launch_band_check(const float *approx, const uint32_t *cnt, const float *qnorm2, float xmax, uint32_t d,
                       uint32_t B, uint32_t k, uint32_t kp, uint32_t *flags, cudaStream_t st);

// ---------------- filter + verify form of the PQ scan (tables.cu, scan3.cu) -------------
// Quantised per-query tables.  qt[q][ch][c] = 8 x u16, position j = n_q[i][c] for sub-space i = 8 ch + ((j + c) & 7):
//   T_q[i][c] in [min_i + step n, min_i + step (n + 1)),  n <= qmax = floor(65535 / m)   (0 for i >= m)
// step[q] = max_i (max_c T - min_c T) / qmax, base[q] = sum_i min_i (- (m - 1) for dot),
// sbound[q] = sum_i max_c |T|, bad[q] = 1 when the table is not finite (the query takes the exact path).
// mm: scratch [B][8 nch][2].
void launch_query_tables_q16(const float *Q, const float *cb_tiled, const float *cb_n2, uint32_t B, uint32_t dim,
                             uint32_t m, uint32_t nch, uint32_t dsub, int metric, float *mm, uint4 *qt, float *step,
                             float *base, float *sbound, uint32_t *bad, cudaStream_t st);
// out[e] = |cb_tiled entry e|^2 for the nch * 256 * 8 tiled codebook entries (open time)
void launch_cb_norms(const float *cb_tiled, uint32_t nch, uint32_t dsub, float *out, cudaStream_t st);
// R[row] = 2 * codeword(row) . centroid(partition(row)); *rmax_bits = float bits of max |R|
void launch_row_const(const unsigned char *codes, const uint64_t *code_base, const uint32_t *part_npad,
                      const uint64_t *part_off, uint32_t nlist, uint64_t nrows, const float *centroids,
                      const float *cb_tiled, uint32_t dim, uint32_t m, uint32_t dsub, float *R, int *rmax_bits,
                      cudaStream_t st);
// exact PQ distances (oracle order) of (query, storage row) pairs; pos == UINT64_MAX -> +inf
void launch_pq_rescore(const float *Q, const uint64_t *pos, uint32_t B, uint32_t nc, const unsigned char *codes,
                       const uint64_t *code_base, const uint32_t *part_npad, const uint64_t *part_off, uint32_t nlist,
                       const float *centroids, const float *cb_tiled, uint32_t dim, uint32_t m, uint32_t dsub, int metric,
                       const uint32_t *ncols_q, float *out, cudaStream_t st);   // ncols_q: optional [B] pairs of each query (<= nc)
// probe_A[slot] = coarse_dist[slot] - |q|^2, amax[q] = max_j coarse + |q|^2, qn2[q] = |q|^2
// (probe_A / amax may be null: dot has no residual)
void launch_probe_terms(const float *probe_dist, const float *Q, uint32_t B, uint32_t nprobes, uint32_t dim,
                        float *probe_A, float *amax, float *qn2, cudaStream_t st);
// the band of every query, for the candidate mode: slack[q] = scale (W + 2E) (same W, E as launch_band_check3);
// also resets thr[q] = CAND_NO_THR, cand_cnt[q] = cand_last[q] = 0 and every key of cand_key to 0xffffffff
void launch_cand_prepare(const float *step, const float *sbound, const float *amax, const int *rmax_bits, const float *qn2,
                         float cb2, float scale, uint32_t m, uint32_t B, float *slack, uint32_t *thr, uint32_t *cand_cnt,
                         uint32_t *cand_last, uint32_t *cand_key, uint32_t cand_cap, cudaStream_t st);
// candidate mode, after the scan: per query, drop the candidates above the final threshold, re-score the survivors
// exactly (oracle arithmetic, as launch_pq_rescore), and write the k best by (_distance, _rowid) -- ids, distances,
// count and (optional) storage positions.  flags[q] = 1 when the query must be redone by the exact kernels (list
// overflow, or bad[q]); such a query's outputs are overwritten by the fix-up pass.
struct FinalizeArgs {
    const float *Q;               // [B][dim] (normalised for cosine)
    const CandRec *cand; const uint32_t *cand_cnt; uint32_t cand_cap;
    const uint32_t *cand_key;     // [B][cand_cap] keys of the records (coalesced copy of CandRec::lb)
    const uint32_t *thr; const float *slack; const uint32_t *bad;
    const unsigned char *codes; const uint64_t *code_base; const uint32_t *part_npad; const uint64_t *part_off;
    const uint64_t *row_ids; const float *centroids; const float *cb_tiled;
    uint32_t B, dim, m, dsub, k; int metric;
    uint64_t *out_ids; float *out_dist; uint32_t *out_count; uint64_t *out_pos; uint32_t *flags;
    unsigned long long *stats;    // optional [4]: candidates appended, survivors re-scored, queries flagged, queries
    // scratch of the three finalize kernels
    uint2 *work;                  // [B * cand_cap] survivors to re-score: (query, candidate index << 16 | survivor slot)
    uint32_t *work_cnt;           // [2]: survivors; gate word of the fix-up pass (set when a query is flagged)
    uint32_t *surv_cnt;           // [B] survivors per query
    float *ex_dist; uint64_t *ex_id; uint64_t *ex_pos;   // [B][cand_cap] exact distance / row id / storage position
    int num_sms;
};
void launch_cand_finalize(const FinalizeArgs &a, cudaStream_t st);
// flags[q] = 1 when the shortlist of q (lower bounds `lb` ascending, [B][kp], cnt valid) cannot be proven to hold
// the exact top-k: proven iff cnt < kp or lb[kp-1] > lb[k-1] + scale (W + 2E),  W = m step (1 + 2^-10),
// E = 2^-15 ceil(m/96) (sbound + amax + rmax + m + 2 (qn2 + cb2)); also 1 when bad[q].  amax / rmax_bits may be null
// (dot).  cb2 = sum_i max_c |codebook_i[c]|^2.  surv (optional, [B]): the length of the ascending prefix
// lb <= lb[k-1] + scale (W + 2E) -- the only rows that can be among the exact top-k, hence the only ones to re-score.
void launch_band_check3(const float *lb, const uint32_t *cnt, const float *step, const float *sbound, const float *amax,
                        const int *rmax_bits, const uint32_t *bad, const float *qn2, float cb2, float scale, uint32_t m,
                        uint32_t B, uint32_t k, uint32_t kp, uint32_t *flags, uint32_t *gate, uint32_t *surv, cudaStream_t st);

// ---------------- index build (build.cu) ----------------------------------------------
// codes[row][i] = argmin_c entry(row's residual sub-vector i, codebook_i[c]) (ties: lowest c); X normalised for cosine
void launch_pq_encode(const float *X, const uint32_t *parts, const float *centroids, const float *codebook,
                      uint64_t n, uint32_t dim, uint32_t m, int metric, unsigned char *codes, cudaStream_t st);

// ---------------- index training (kmeans.cu) ----------------------------------------------
// one Lloyd update: rows bucketed by `assign` (u64 centre id per row, >= k = unassigned), every non-empty centre replaced
// by the mean of its rows (f64 sums in ascending row order).  counts [k], offsets [k+1], cursor [k], rows [n]: scratch
void launch_kmeans_update(const uint64_t *assign, const float *x, uint64_t n, uint32_t dim, uint32_t k, uint32_t *counts,
                          uint32_t *offsets, uint32_t *cursor, uint32_t *rows, float *centroids, cudaStream_t st);
// *out = sum of the finite dist[i]
void launch_kmeans_inertia(const float *dist, uint64_t n, double *out, cudaStream_t st);
// one Lloyd update of all m PQ codebooks: codebook[i][c] = mean of the sub-vectors i of the rows with codes[row][i] == c
void launch_pq_update(const float *x, const unsigned char *codes, uint64_t n, uint32_t dim, uint32_t m, double *sums,
                      uint32_t *counts, float *codebook, cudaStream_t st);

// ---------------- index re-layout (open time) --------------------------------------
void launch_retile_codes(const unsigned char *codes, int layout, const uint64_t *part_off, uint32_t nlist,
                         uint64_t nrows, uint32_t m, uint32_t nch, const uint64_t *code_base,
                         const uint32_t *part_npad, unsigned char *out, cudaStream_t st);
void launch_retile_codebook(const float *codebook, uint32_t m, uint32_t dsub, uint32_t nch, float *out,
                            cudaStream_t st);

}  // namespace lgpu
