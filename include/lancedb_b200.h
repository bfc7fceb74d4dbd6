/*
 * lancedb_b200.h -- C ABI of the B200-native LanceDB vector-query hot path.
 *
 * This is the drop-in boundary: the entry points the Rust
 * `lancedb` crate would bind through `extern "C"` at the single place it hands a
 * vector query to lance today --
 *     rust/lancedb/src/table/query.rs:219-249   ds.scan() / scanner.nearest(..) /
 *                                                minimum_nprobes / maximum_nprobes
 *     rust/lancedb/src/table/query.rs:251-316   limit+offset, distance_range,
 *                                                use_index, refine, distance_metric
 *     rust/lancedb/src/table/query.rs:327, :121 create_plan() + execute_plan()
 * -- i.e. everything below `NativeTable::create_plan` for `AnyQuery::VectorQuery`
 * (rust/lancedb/src/table.rs:3301-3315).  INTEGRATION.md shows the Rust binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns an lgpu_status; on failure `lgpu_last_error()`
 *     (thread-local, valid until the next call on that thread) carries the message
 *     and the status maps onto `lancedb::Error`
 *     (rust/lancedb/src/error.rs:57-130): INVALID_INPUT -> Error::InvalidInput,
 *     RUNTIME/OOM -> Error::Runtime, TIMEOUT -> Error::Timeout;
 *   - the caller owns all host buffers; inputs are borrowed for the call only,
 *     outputs are caller-allocated; the library owns device memory behind the
 *     opaque handles; index arrays are copied to HBM at open (the analogue of
 *     `prewarm_index`, rust/lancedb/src/table.rs:3283-3286);
 *   - handles are thread-safe and `lgpu_*_search*` is re-entrant (each call takes
 *     a private stream + workspace), matching `BaseTable: Send + Sync`
 *     (rust/lancedb/src/table.rs:549);
 *   - no CPU fallback exists: without a CUDA device every compute entry point
 *     fails with LGPU_RUNTIME.
 */
#ifndef LANCEDB_B200_H
#define LANCEDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGPU_ABI_VERSION 2

typedef enum {
    LGPU_OK = 0,
    LGPU_INVALID_INPUT = 1,  /* lancedb::Error::InvalidInput */
    LGPU_RUNTIME = 2,        /* lancedb::Error::Runtime (CUDA failure, no device) */
    LGPU_TIMEOUT = 3,        /* lancedb::Error::Timeout */
    LGPU_OOM = 4             /* device or host allocation failure */
} lgpu_status;

/* lancedb::DistanceType, rust/lancedb/src/lib.rs:236-260 (hamming is u8-only and
 * not on this path: rust/lancedb/src/table/query.rs:230-236) */
typedef enum { LGPU_L2 = 0, LGPU_COSINE = 1, LGPU_DOT = 2 } lgpu_metric;

typedef enum {
    LGPU_CODES_ROW_MAJOR = 0,             /* [nrows][m], rows partition-contiguous */
    LGPU_CODES_PARTITION_TRANSPOSED = 1   /* per partition [m][n_p] (lance in-memory form) */
} lgpu_codes_layout;

typedef struct lgpu_index lgpu_index;   /* an IVF_PQ index resident in HBM */
typedef struct lgpu_flat lgpu_flat;     /* a raw vector column resident in HBM */

/* The arrays of one IVF_PQ index (lance v2 `IvfPq`,
 * rust/lancedb/src/table/create_index.rs:283-303, :772): IVF centroids, PQ codebook
 * (8-bit, 256 centroids per sub-vector), PQ codes and row ids grouped by partition. */
typedef struct {
    uint32_t abi_version;        /* LGPU_ABI_VERSION */
    uint32_t dim;
    uint32_t nlist;              /* IVF partitions */
    uint32_t m;                  /* PQ sub-vectors; dim % m == 0 */
    uint32_t nbits;              /* must be 8 */
    int32_t  metric;             /* lgpu_metric the index was trained with */
    int32_t  codes_layout;       /* lgpu_codes_layout */
    int32_t  device;             /* CUDA device ordinal */
    uint64_t nrows;
    const float    *centroids;    /* [nlist][dim] */
    const float    *codebook;     /* [m][256][dim/m] */
    const uint64_t *part_offsets; /* [nlist+1] row offset of each partition */
    const uint8_t  *codes;        /* nrows*m bytes, layout above */
    const uint64_t *row_ids;      /* [nrows] `_rowid` of each stored row */
    const float    *vectors;      /* optional [nrows][dim] raw vectors in the same row
                                     order (enables refine_factor); NULL if absent */
} lgpu_index_desc;

/* One vector query request = lancedb::query::VectorQueryRequest
 * (rust/lancedb/src/query.rs:1066-1114) reduced to what reaches the kernels. */
typedef struct {
    uint32_t k;              /* limit + offset (table/query.rs:231); default 10 */
    uint32_t nprobes;        /* maximum_nprobes (== minimum_nprobes == 20 by default) */
    uint32_t refine_factor;  /* 0 = no refine (query.rs:1302-1332) */
    int32_t  has_lower;      /* distance_range lower bound, inclusive */
    int32_t  has_upper;      /* distance_range upper bound, exclusive */
    float    lower;
    float    upper;
    uint32_t flags;          /* reserved, 0 */
    uint32_t max_nprobes;    /* maximum_nprobes (query.rs:1250-1275): under a prefilter, queries that found fewer
                                than k rows in their `nprobes` nearest partitions are searched again over their
                                max_nprobes nearest; 0 or <= nprobes = no widening */
    uint32_t timeout_ms;     /* QueryExecutionOptions::timeout (query.rs:626-658, utils/mod.rs:328-393): the
                                host-buffer entry points return LGPU_TIMEOUT, leaving the outputs untouched,
                                when the results are not ready this many ms after the call started; 0 = none */
} lgpu_search_params;

/* ---- library ----------------------------------------------------------- */
const char *lgpu_last_error(void);
uint32_t    lgpu_abi_version(void);
int         lgpu_device_count(int *count);

/* ---- IVF_PQ (ANNIvfPartitionExec -> ANNIvfSubIndexExec -> TopK) ---------- */
int  lgpu_index_open(const lgpu_index_desc *desc, lgpu_index **out);
void lgpu_index_close(lgpu_index *ix);
/* bytes of HBM held by the index (codes + ids + centroids + codebook + vectors) */
int  lgpu_index_device_bytes(const lgpu_index *ix, uint64_t *bytes);
/* sum over a batch of the PQ code bytes its probes must scan (the roofline's
 * algorithmic bytes); valid for the most recent search on the
 * calling thread */
int  lgpu_last_scanned_code_bytes(uint64_t *bytes);

/* Search a batch of B queries held in HOST memory.  queries: [B][dim] f32.
 * out_ids/out_dist: [B][k] (`_rowid` / `_distance`, ascending by
 * (_distance,_rowid)); out_count: [B] valid entries per query; unused slots are
 * UINT64_MAX / +inf.  H2D and D2H copies happen inside the call. */
int lgpu_search(lgpu_index *ix, const float *queries, uint32_t B,
                const lgpu_search_params *params,
                uint64_t *out_ids, float *out_dist, uint32_t *out_count);

/* Asynchronous form of lgpu_search (SURVEY.md 8b "Threading": today the Python binding parks the blocking call on
 * spawn_blocking, python/src/runtime.rs:113-119).  Returns once the copies and kernels are enqueued on a private
 * stream; every buffer must stay valid (and should be page-locked for the copies to overlap) until
 * lgpu_ticket_wait returns.  Two or more tickets in flight pipeline: batch i+1's H2D overlaps batch i's kernels. */
typedef struct lgpu_ticket lgpu_ticket;
int lgpu_search_async(lgpu_index *ix, const float *queries, uint32_t B,
                      const lgpu_search_params *params,
                      uint64_t *out_ids, float *out_dist, uint32_t *out_count, lgpu_ticket **ticket);
int lgpu_ticket_poll(lgpu_ticket *ticket, int *done);   /* *done = 1 when the results are in place */
int lgpu_ticket_wait(lgpu_ticket *ticket);              /* blocks, frees the ticket, returns the call's status */

/* Single-vector search that may share a batch with concurrent callers (SURVEY.md 8b "Threading": the reference serves
 * many one-vector queries from tokio workers, each through its own plan -- rust/lancedb/src/table/query.rs:201-215).
 * Calls with identical parameters arriving within a short window (LGPU_COALESCE_US, default 50 us) are gathered by the
 * first arrival into ONE lgpu_search; every caller gets exactly the rows a solitary lgpu_search(B = 1) would return.
 * query: [dim]; out_ids / out_dist: [k]; out_count: [1]. */
int lgpu_search_coalesced(lgpu_index *ix, const float *query, const lgpu_search_params *params,
                          uint64_t *out_ids, float *out_dist, uint32_t *out_count);

/* Prefiltered search: the reference's default filter mode ("filtering will be performed
 * before the vector search", rust/lancedb/src/query.rs:489-507; the row-id allow-list the
 * scalar filter produced is what lance hands to the ANN nodes as a pre-filter [lance,
 * recalled]).  `allow` is a host bitmap over row ids: bit (r & 31) of word r >> 5 set = row
 * id r may be returned; ids >= allow_bits are excluded.  Excluded rows are dropped before the
 * top-k (and before refine), so up to k allowed rows come back from the probed partitions. */
int lgpu_search_filtered(lgpu_index *ix, const float *queries, uint32_t B,
                         const lgpu_search_params *params,
                         const uint32_t *allow, uint64_t allow_bits,
                         uint64_t *out_ids, float *out_dist, uint32_t *out_count);

/* Same, all five buffers in DEVICE memory of the index's device; enqueued on
 * `cuda_stream` (a cudaStream_t, may be 0) and NOT synchronised on return. */
int lgpu_search_device(lgpu_index *ix, const float *d_queries, uint32_t B,
                       const lgpu_search_params *params,
                       uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count,
                       void *cuda_stream);

/* Merge `nlists` per-rank top-k lists per query (device buffers laid out
 * [nlists][B][k], as produced by an all-gather of lgpu_search_device outputs over
 * partition-sharded indexes) into the global top-k by (_distance,_rowid). */
int lgpu_merge_topk_device(int device, uint32_t nlists, uint32_t B, uint32_t k,
                           const uint64_t *d_ids, const float *d_dist,
                           uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count,
                           void *cuda_stream);

/* ---- partition-sharded search across GPUs (the reference is single-process, so there is no
 * reference interface to replace -- this is what north_star adds for an index larger than one GPU's HBM).
 * One process (or thread) per GPU.  Centroids and codebook are replicated, every partition's codes and row ids
 * live on exactly one rank (non-owned partitions are empty in that rank's lgpu_index), every rank receives the
 * same query batch, and the only data-path collective is ONE ncclAllGather of [B][k] 16-byte (_rowid, _distance)
 * records per batch, merged on every rank by (_distance, _rowid).  NCCL is bound at run time (dlopen of
 * libnccl.so.2), so hosts that never shard need no NCCL.  The host distributes the unique id over whatever
 * channel it already has (the Rust crate: its RPC layer; the Python mirror: torch.distributed / a file). ---- */
#define LGPU_COMM_ID_BYTES 128
typedef struct lgpu_comm lgpu_comm;
/* rank 0: create the group's id (an ncclUniqueId), to be handed to every rank */
int  lgpu_comm_unique_id(void *id_out, size_t id_bytes);
/* every rank, collectively: join the group on CUDA device `device` */
int  lgpu_comm_init(const void *unique_id, size_t id_bytes, int rank, int world, int device, lgpu_comm **out);
void lgpu_comm_destroy(lgpu_comm *comm);
/* collective: every rank passes the same B queries (host buffers) and its own shard; every rank receives the
 * global top-k.  refine_factor must be 0. */
int  lgpu_search_sharded(lgpu_index *shard, lgpu_comm *comm, const float *queries, uint32_t B,
                         const lgpu_search_params *params,
                         uint64_t *out_ids, float *out_dist, uint32_t *out_count);
/* same with device buffers, enqueued on `cuda_stream`, not synchronised */
int  lgpu_search_sharded_device(lgpu_index *shard, lgpu_comm *comm, const float *d_queries, uint32_t B,
                                const lgpu_search_params *params,
                                uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count,
                                void *cuda_stream);
/* device time (ms) of the local search, the all-gather and the merge of the most recent sharded call made with
 * profiling on (lgpu_set_profiling).  times: [3] */
int  lgpu_comm_last_stage_ms(lgpu_comm *comm, float *times);

/* ---- index build passes (reference: IVF_PQ build through lance, parameters at
 * rust/lancedb/src/table/create_index.rs:283-303, rust/lancedb/src/index/vector.rs:246-319; "GPU support in
 * building vector index", python/python/lancedb/table.py:2883-2937).  k-means training stays with the caller;
 * these are the two passes over every row, bit-consistent with the search kernels.  All pointers are HOST
 * memory; vectors are raw rows (normalised internally for LGPU_COSINE). ------------------------------- */
/* out_parts[r] = the partition find_partitions(vectors[r], nprobes = 1) returns */
int lgpu_ivf_assign(const float *centroids, uint32_t nlist, uint32_t dim, int metric,
                    const float *vectors, uint64_t n, int device, uint32_t *out_parts);
/* out_codes[r][i] (row-major [n][m]) = the codeword of sub-space i with the smallest distance-table entry
 * for row r's residual (row - centroid[parts[r]]; the row itself for LGPU_DOT); ties go to the lowest code */
int lgpu_pq_encode(const float *centroids, const float *codebook, uint32_t nlist, uint32_t dim, uint32_t m,
                   int metric, const float *vectors, const uint32_t *parts, uint64_t n, int device,
                   unsigned char *out_codes);

/* k-means TRAINING on the GPU (the Lloyd loops of the IVF_PQ build; parameters max_iterations / sample_rate at
 * rust/lancedb/src/index/vector.rs:286-297).  x: HOST [n][dim] training rows (already sampled, normalised for cosine).
 * centroids: in = initial centres (e.g. k sampled rows), out = trained centres.  Assignment is the search's own coarse
 * step (exact nearest centre), empty clusters keep their previous centre.  *inertia_out (optional) = sum of squared
 * distances of the rows to their nearest trained centre. */
int lgpu_kmeans_train(const float *x, uint64_t n, uint32_t dim, float *centroids, uint32_t k, uint32_t iters, int device,
                      double *inertia_out);
/* the 256-entry PQ codebooks of all m sub-spaces: x = HOST [n][dim] rows to quantise (residuals for l2 / cosine, raw rows
 * for dot); codebook [m][256][dim/m] in = initial codewords, out = trained */
int lgpu_pq_train(const float *x, uint64_t n, uint32_t dim, uint32_t m, float *codebook, uint32_t iters, int device);

/* ---- flat / brute force (LanceRead -> KNNVectorDistance -> TopK) --------- */
int  lgpu_flat_open(const float *vectors, uint64_t nrows, uint32_t dim,
                    const uint64_t *row_ids /* NULL => 0..nrows-1 */, int device,
                    lgpu_flat **out);
void lgpu_flat_close(lgpu_flat *fl);
int  lgpu_flat_search(lgpu_flat *fl, int metric, const float *queries, uint32_t B,
                      const lgpu_search_params *params,
                      uint64_t *out_ids, float *out_dist, uint32_t *out_count);
/* flat search under a row-id allow-list (same bitmap as lgpu_search_filtered) */
int  lgpu_flat_search_filtered(lgpu_flat *fl, int metric, const float *queries, uint32_t B,
                               const lgpu_search_params *params,
                               const uint32_t *allow, uint64_t allow_bits,
                               uint64_t *out_ids, float *out_dist, uint32_t *out_count);
int  lgpu_flat_search_device(lgpu_flat *fl, int metric, const float *d_queries, uint32_t B,
                             const lgpu_search_params *params,
                             uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count,
                             void *cuda_stream);

/* ---- per-stage access (parity localisation and kernel benchmarks) -------- */
/* coarse stage only: the nprobes nearest partitions of each query and their
 * distances (host buffers, [B][nprobes]) */
int lgpu_debug_coarse(lgpu_index *ix, const float *queries, uint32_t B, uint32_t nprobes,
                      uint32_t *out_parts, float *out_dists);
/* final PQ distances of every row of partition `part` for one query (host
 * buffers; out has n_p floats) -- exercises the LUT build + code scan kernels */
int lgpu_debug_partition_distances(lgpu_index *ix, const float *query, uint32_t part,
                                   float *out);
/* the tensor-core shortlist GEMM alone: out[q][x] = |x|^2 - 2 bf16(Q[q]).bf16(X[x]) (host buffers,
 * out is [B][N] f32); dim must be a multiple of 8 */
int lgpu_debug_gemm(const float *queries, const float *vectors, uint32_t B, uint64_t N, uint32_t dim,
                    int device, float *out);
/* per-kernel device time (ms) of the most recent lgpu_search* call made with
 * LGPU_PROFILE=1 in the environment: coarse, select-probes, group, scan, top-k,
 * refine, total.  times: [7] */
int lgpu_last_stage_ms(float *times);
/* filter-scan counters of the most recent profiled lgpu_search* call on this thread (first sub-batch): candidates the
 * scanners appended, survivors re-scored exactly, queries sent to the exact fix-up pass, queries.  stats: [4] */
int lgpu_last_filter_stats(uint64_t *stats);
/* kernels this process has launched through the library so far (eager launches and graph replays alike) */
int lgpu_kernel_launch_count(uint64_t *count);
/* switch per-stage CUDA-event timing (and the scanned-bytes counter) on/off for the
 * calling process; overrides LGPU_PROFILE.  While on, lgpu_search_device synchronises
 * the stream before returning. */
int lgpu_set_profiling(int enabled);

#ifdef __cplusplus
}
#endif
#endif
