// api.cu -- the extern "C" boundary (include/lancedb_b200.h): handle management,
// workspaces, and the host-side orchestration of one batched vector query:
//   [cosine: normalise] -> K1 exact centroid distances -> select nprobes ->
//   regroup probe slots by partition -> K2+K3 fused LUT build + code scan ->
//   K4 top-k by (_distance,_rowid) -> [refine: exact re-rank] .
// Everything runs on one stream per call with no host round trip in between.
// The reference-side equivalent is NativeTable::create_plan + execute_plan
// (rust/lancedb/src/table/query.rs:131-328, :121).
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>
#include <unistd.h>

namespace lgpu {

static thread_local std::string g_err;
static thread_local float g_stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};
static thread_local uint64_t g_scanned_bytes = 0;
static thread_local uint64_t g_filter_stats[4] = {0, 0, 0, 0};
void set_error(const std::string &msg) { g_err = msg; }

// every kernel the library launches (eagerly, into a stream capture, or through a graph replay) is counted
static std::atomic<uint64_t> g_kernel_launches{0};
static thread_local bool g_capturing = false;        // launches made while capturing are counted per replay instead
static thread_local uint64_t g_captured_launches = 0;
bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("LGPU_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

void count_launches(uint64_t n)
{
    if (g_capturing) g_captured_launches += n;
    else g_kernel_launches.fetch_add(n, std::memory_order_relaxed);
}

static std::atomic<int> g_profiling{-1};
static bool profiling_enabled()
{
    int v = g_profiling.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("LGPU_PROFILE");
        int want = (e && e[0] == '1') ? 1 : 0;
        g_profiling.compare_exchange_strong(v, want);     // lgpu_set_profiling may have won the race: keep its value
        v = g_profiling.load(std::memory_order_relaxed);
    }
    return v == 1;
}
static size_t workspace_budget()
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("LGPU_WS_BYTES");
        v = e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)8 << 30);
        if (v < ((size_t)1 << 20)) v = (size_t)1 << 20;
    }
    return v;
}

// bumped by every device (re)allocation in the process.  A captured CUDA graph bakes workspace pointers in, so a
// graph is only replayed while the epoch still equals the one recorded at capture (Workspace::graph_epoch): any
// entry point, on any thread, that grows a buffer invalidates every captured graph (they re-capture on next use).
static std::atomic<uint64_t> g_alloc_epoch{0};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    void ensure(size_t n)
    {
        if (n <= bytes) return;
        g_alloc_epoch++;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        size_t want = n + n / 8;
        LGPU_CUDA(cudaMalloc(&p, want));
        bytes = want;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct Workspace {
    cudaStream_t stream = nullptr;     // private stream (host-buffer entry points)
    cudaStream_t aux = nullptr;        // side stream: the per-query tables are built while the coarse step runs
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int stats_mode = 0;                 // profiling: 1 = the last sub-batch ran the candidate mode, 2 = the dense filter
    bool aux_open = false;              // an EAGER fork onto `aux` has not been joined yet (a call failed half-way)
    cudaEvent_t done = nullptr;        // last use, for cross-stream reuse
    cudaEvent_t ev[8] = {};
    DevBuf q, qn, xnorm, D, probes, probe_dist, probe_cnt;
    DevBuf part_cnt, slot_pos, seg_local, qtot, seg_off, qlist_off, tile_off, tile_off_b, qlist, scalars, tile_desc, allow;
    DevBuf dist_out, out_ids, out_dist, out_count;
    DevBuf t_ids, t_dist, t_pos, t_cnt, t_exact;
    DevBuf widen;                       // maximum_nprobes widening: queries that found fewer than k rows
    DevBuf qb, qn2, flags;              // tensor-core shortlist: bf16 queries, |q|^2, unproven-query flags
    // CUDA graph of one host-buffer search (lgpu_search): the ~15 launches of a batch replayed as one
    cudaGraphExec_t graph = nullptr;
    uint64_t graph_key[4] = {0, 0, 0, 0};
    uint64_t graph_epoch = 0;           // g_alloc_epoch when `graph` was captured
    uint64_t graph_kernels = 0;         // kernel launches one replay stands for
    int graph_state = 0;                // 0: next call runs eagerly (warm-up), 1: capture, 2: replay, -1: disabled
    DevBuf sbound, probe_A, amax;       // tensor-core shortlists: thresholds / counters; filter scan: bounds, per-probe scalars
    DevBuf qt, qt_mm, qt_step, qt_base, qt_bad;   // filter scan (scan3.cu): quantised per-query tables
    DevBuf s_ids, s_lb, s_pos, s_cnt, s_exact;    // filter scan (dense mode): shortlist by lower bound, exact re-score
    DevBuf c_stats;                               // candidate-mode counters (profiling only)
    DevBuf c_work, c_wcnt, c_surv, c_exd, c_exi, c_exp;   // candidate mode: survivor work list and exact results
    DevBuf c_thr, c_slack, c_cnt, c_rec, c_key, c_last;          // filter scan (candidate mode): thresholds, bands, candidate lists
    Workspace()
    {
        LGPU_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        LGPU_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        LGPU_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        LGPU_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        LGPU_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        for (auto &e : ev) LGPU_CUDA(cudaEventCreate(&e));
    }
    ~Workspace()
    {
        if (graph) cudaGraphExecDestroy(graph);
        if (stream) cudaStreamDestroy(stream);
        if (aux) cudaStreamDestroy(aux);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (done) cudaEventDestroy(done);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
    }
};

struct WorkspacePool {
    std::mutex mu;
    std::vector<Workspace *> free_list;
    Workspace *take()
    {
        {
            std::lock_guard<std::mutex> g(mu);
            if (!free_list.empty()) { Workspace *w = free_list.back(); free_list.pop_back(); return w; }
        }
        return new Workspace();
    }
    void give(Workspace *w) { std::lock_guard<std::mutex> g(mu); free_list.push_back(w); }
    ~WorkspacePool() { for (auto *w : free_list) delete w; }
};

}  // namespace lgpu

using namespace lgpu;

// ---- live-handle registry: every entry point resolves its handle through it, so a handle that was closed (or
// that belongs to the parent of a fork()) is rejected instead of dereferenced, and close waits for the calls
// still inside the handle (BaseTable is Send + Sync: rust/lancedb/src/table.rs:549; queries are re-executable
// from any tokio worker: rust/lancedb/src/query.rs:954-955).
namespace lgpu {
static std::mutex g_live_mu;
static std::condition_variable g_live_cv;
static std::unordered_set<const void *> g_live;
static std::atomic<bool> g_cuda_touched{false};   // this process has created a CUDA context through the library
static std::atomic<bool> g_fork_poisoned{false};  // we are the child of a fork() taken after that
// fork(): the reference rebuilds its tokio runtime in the child (python/src/runtime.rs:68-81); a CUDA context has
// the same constraint and cannot be rebuilt, so the child drops every handle and refuses GPU work.
static void atfork_prepare() { g_live_mu.lock(); }
static void atfork_parent() { g_live_mu.unlock(); }
static void atfork_child()
{
    g_live.clear();                                  // the parent's handles are not valid here (leaked, never freed)
    if (g_cuda_touched.load()) g_fork_poisoned.store(true);
    g_live_mu.unlock();
}
static void register_handle(const void *h)
{
    static std::once_flag once;
    std::call_once(once, [] { pthread_atfork(atfork_prepare, atfork_parent, atfork_child); });
    std::lock_guard<std::mutex> g(g_live_mu);
    g_live.insert(h);
}
template <class H> struct HandleRef {
    H *h;
    HandleRef(H *p, const char *what) : h(nullptr)
    {
        std::lock_guard<std::mutex> g(g_live_mu);
        if (!p || !g_live.count(p)) {
            set_error(std::string(what) + " handle is null, closed, or was opened in another process (fork)");
            throw Failure{LGPU_INVALID_INPUT};
        }
        p->refs.fetch_add(1);
        h = p;
    }
    ~HandleRef()
    {
        if (h && h->refs.fetch_sub(1) == 1) { std::lock_guard<std::mutex> g(g_live_mu); g_live_cv.notify_all(); }
    }
    H *operator->() const { return h; }
};
// unregister and wait until no call is inside the handle any more; false = it was not a live handle
template <class H> static bool retire_handle(H *p)
{
    std::unique_lock<std::mutex> g(g_live_mu);
    if (!p || !g_live.erase(p)) return false;
    g_live_cv.wait(g, [&] { return p->refs.load() == 0; });
    return true;
}
}  // namespace lgpu

// micro-batcher of concurrent single-vector calls (SURVEY.md 8b "Threading": many tokio workers each with one query
// vector).  Callers with identical parameters that arrive within a short window ride one batched search: the first
// arrival leads -- waits for the window (or a full batch), takes the queue, runs lgpu_search on the gathered
// queries, scatters the rows -- the others sleep on the condition variable until their row is filled in.
namespace lgpu {
struct PendingQuery {
    const float *q; uint64_t *ids; float *dist; uint32_t *cnt;
    int status = LGPU_OK; bool done = false; std::string err;
};
struct Coalescer {
    std::mutex mu;
    std::condition_variable cv;
    struct Lane { lgpu_search_params params; std::vector<PendingQuery *> queue; bool leader = false; };
    std::vector<Lane *> lanes;                           // one per distinct parameter set seen (a handful)
    ~Coalescer() { for (auto *l : lanes) delete l; }
};
}  // namespace lgpu

struct lgpu_index {
    std::atomic<int> refs{0};
    Coalescer coalescer;
    int device = 0, num_sms = 0;
    uint32_t dim = 0, nlist = 0, m = 0, dsub = 0, nch = 0;
    uint32_t max_nrb = 1;          // row blocks (of 1536 rows) of the largest partition: bounds the tile count
    int metric = 0;
    uint64_t nrows = 0, device_bytes = 0;
    DevBuf centroids, cb_tiled, codes, code_base, part_n, part_npad, part_off, row_ids, vectors;
    DevBuf row_R, rmax_bits;            // filter scan: per-row constant 2 b.c and max |R| (tables.cu)
    DevBuf cb_n2;                       // filter scan: |b|^2 of every tiled codebook entry
    float cb2 = 0.f;                    // sum_i max_c |codebook_i[c]|^2 (error budget of the filter's table entries)
    bool has_tables = false;
    DevBuf cent_b, cent_n2;             // bf16 centroids + |c|^2 for the tensor-core coarse step
    DevBuf cent_sb, cent_sn2;           // every COARSE_SAMPLE_STRIDE-th centroid (bf16 + |c|^2): the threshold sample
    uint32_t cent_ns = 0;               // rows of the sample (0: none)
    float cent_max = 0.f;
    bool has_tc = false;
    bool has_vectors = false;
    std::vector<uint64_t> pad_prefix;   // prefix sums of pad4(n_p) sorted descending
    std::vector<uint32_t> h_part_n;
    WorkspacePool pool;
};

struct lgpu_flat {
    std::atomic<int> refs{0};
    int device = 0;
    uint64_t nrows = 0;
    uint32_t dim = 0;
    DevBuf vectors, row_ids, ysqrt;
    DevBuf vec_b, vec_n2;               // bf16 rows + |x|^2 for the tensor-core path
    float vec_max = 0.f;
    int num_sms = 0;
    bool has_tc = false;
    bool has_ids = false, has_norms = false;
    std::mutex mu;
    WorkspacePool pool;
};

// ---- NCCL, bound at run time (dlopen): single-GPU hosts need no libnccl, and inside a process that already
// loaded one (torch) the same library instance is used ----
namespace lgpu {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
static NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("LGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = "libnccl.so.2 not found (set LGPU_NCCL_LIB)"; return; }
        auto sym = [&](const char *n) { void *f = dlsym(api.lib, n); if (!f) api.error = std::string("missing NCCL symbol ") + n; return f; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    if (!api.error.empty()) { set_error("NCCL unavailable: " + api.error); throw Failure{LGPU_RUNTIME}; }
    return api;
}
#define LGPU_NCCL(expr)                                                                          \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            ::lgpu::set_error(std::string(#expr) + ": " + ::lgpu::nccl_api().GetErrorString(_r)); \
            throw ::lgpu::Failure{LGPU_RUNTIME};                                                 \
        }                                                                                        \
    } while (0)
}  // namespace lgpu

// one rank of a partition-sharded search group: an NCCL communicator plus the gather buffers
struct lgpu_comm {
    std::atomic<int> refs{0};
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    std::mutex mu;                       // collectives of one communicator are issued one call at a time
    DevBuf send, recv;                   // [B][k] / [world][B][k] TopkRecord
    DevBuf l_ids, l_dist, l_cnt;         // the local (per-shard) top-k before the exchange
    cudaEvent_t ev[3] = {};              // local search done / all-gather done / merge done (stage timing)
    float last_ms[3] = {0, 0, 0};        // local search, all-gather, merge of the most recent profiled call
};

namespace {

struct WsLease {
    WorkspacePool &pool;
    Workspace *ws;
    cudaStream_t st;
    WsLease(WorkspacePool &p, cudaStream_t user, bool use_user) : pool(p), ws(p.take())
    {
        st = use_user ? user : ws->stream;
        // the workspace may still be in use by an earlier call on another stream
        if (cudaStreamWaitEvent(st, ws->done, 0) != cudaSuccess) cudaGetLastError();
    }
    ~WsLease()
    {
        // side-stream work of a call that failed half-way.  Only after an eager fork: an event whose last record sits
        // inside a captured graph cannot be waited on outside the capture (cudaErrorInvalidValue, which would then be
        // reported by the next cudaGetLastError() of an unrelated launch).
        if (ws->aux_open) {
            if (cudaStreamWaitEvent(st, ws->ev_join, 0) != cudaSuccess) cudaGetLastError();
            ws->aux_open = false;
        }
        if (cudaEventRecord(ws->done, st) != cudaSuccess) cudaGetLastError();
        pool.give(ws);
    }
};

// Deadline of one call (QueryExecutionOptions::timeout -> TimeoutStream, rust/lancedb/src/utils/mod.rs:328-393,
// used at rust/lancedb/src/query.rs:1452-1457).  Kernels cannot be cancelled, so the deadline is enforced where
// the host waits: between sub-batches and before the results are copied back.  On expiry the call returns
// LGPU_TIMEOUT without touching the caller's output buffers; work already enqueued drains into the workspace,
// which stays fenced by its `done` event until then.
struct Deadline {
    bool armed = false;
    std::chrono::steady_clock::time_point at;
    explicit Deadline(uint32_t timeout_ms)
    {
        if (timeout_ms) { armed = true; at = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms); }
    }
    bool expired() const { return armed && std::chrono::steady_clock::now() >= at; }
    // wait for everything enqueued on `st` so far, or for the deadline
    void wait(cudaStream_t st, cudaEvent_t ev) const
    {
        if (!armed) return;
        LGPU_CUDA(cudaEventRecord(ev, st));
        for (;;) {
            cudaError_t e = cudaEventQuery(ev);
            if (e == cudaSuccess) return;
            if (e != cudaErrorNotReady) LGPU_CUDA(e);
            if (expired()) { set_error("Query timeout"); throw Failure{LGPU_TIMEOUT}; }
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
};

// prefilter: device bitmap over row ids (nullptr = no filter)
struct RowFilter {
    const uint32_t *bits = nullptr;
    uint64_t nbits = 0;
};

// LGPU_EXACT_SCAN=1 sends every query through the exact kernel (scan2.cu) instead of filter + verify
// (scan3.cu + tables.cu).  Results are bit-identical either way; the switch exists for A/B timing and tests.
static bool exact_scan_forced()
{
    const char *e = getenv("LGPU_EXACT_SCAN");
    return e && e[0] == '1';
}

// Mode switches of the IVF path, re-read on every call (tests flip them between calls; getenv is ~100 ns) and folded
// into the CUDA-graph key so a captured launch sequence is never replayed under a different mode.
struct ScanModes {
    bool exact, dense_forced;
    uint32_t cand_kmax, cap_env, small_slots, coarse_list_min;
    uint64_t signature() const
    {
        return ((uint64_t)exact | (uint64_t)dense_forced << 1 | (uint64_t)cand_kmax << 8 | (uint64_t)cap_env << 20 |
                (uint64_t)small_slots << 36 | (uint64_t)(coarse_list_min & 0xffffu) << 48) * 0x9e3779b97f4a7c15ull;
    }
};
static ScanModes scan_modes()
{
    auto num = [](const char *name, uint32_t dflt) { const char *e = getenv(name); return e ? (uint32_t)atoi(e) : dflt; };
    ScanModes m;
    m.exact = exact_scan_forced();
    m.dense_forced = getenv("LGPU_DENSE_FILTER") != nullptr;
    m.cand_kmax = std::min<uint32_t>(CAND_TOPK_MAX, num("LGPU_CAND_KMAX", CAND_TOPK_MAX));
    m.cap_env = num("LGPU_CAND_CAP", 0u);
    m.small_slots = num("LGPU_SMALL_SLOTS", 1024u);
    m.coarse_list_min = num("LGPU_COARSE_LIST_MIN", 8192u);     // nlist from which the coarse GEMM filters (0: never)
    return m;
}

static bool tc_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("LGPU_NO_TENSOR_CORE"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// bf16 copy + squared norms + max norm of a row-major f32 matrix (open time)
static void prepare_tc_operand(const float *X, uint64_t n, uint32_t d, DevBuf &Xb, DevBuf &n2, float &xmax, cudaStream_t st)
{
    Xb.ensure(std::max<size_t>((size_t)n * d * 2, 16));
    n2.ensure(std::max<size_t>((size_t)n * 4, 16));
    launch_to_bf16(X, n, d, Xb.p, n2.as<float>(), st);
    std::vector<float> h(n);
    LGPU_CUDA(cudaMemcpyAsync(h.data(), n2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    LGPU_CUDA(cudaStreamSynchronize(st));
    float m = 0.f;
    for (float v : h) m = std::max(m, v);
    xmax = std::sqrt(m) * 1.0001f;
}

// Tensor-core shortlist + exact re-score (squared L2 only): the k best of the N rows of X for each
// of B queries by exact lance-order distance, ids from col_ids (or the row index), ascending by
// (distance, id).  Dbuf: [B][ld] f32 scratch.  Queries whose shortlist cannot be proven complete
// (band_check) are redone by the exact kernels in the same stream, without a host round trip.
static void tc_topk_l2(Workspace *ws, cudaStream_t st, int num_sms, const float *Q, uint32_t B, const float *X,
                       const void *Xb, const float *xnorm2, float xmax, uint64_t N, uint32_t d,
                       const uint64_t *col_ids, uint32_t k, uint32_t kp, uint64_t *out_ids, float *out_dist,
                       uint32_t *out_cnt, float *Dbuf, uint64_t ld)
{
    ws->qb.ensure((size_t)B * d * 2); ws->qn2.ensure((size_t)B * 4); ws->flags.ensure((size_t)B * 4);
    ws->t_ids.ensure((size_t)B * kp * 8); ws->t_dist.ensure((size_t)B * kp * 4);
    ws->t_pos.ensure((size_t)B * kp * 8); ws->t_cnt.ensure((size_t)B * 4); ws->t_exact.ensure((size_t)B * kp * 4);
    launch_to_bf16(Q, B, d, ws->qb.p, ws->qn2.as<float>(), st);
    launch_gemm_dist(ws->qb.p, Xb, xnorm2, B, N, d, Dbuf, ld, num_sms, st);
    SelectArgs sa{};
    sa.mode = 1; sa.dense = Dbuf; sa.ncols = N; sa.row_stride = ld; sa.col_ids = col_ids;
    sa.B = B; sa.k = kp;
    sa.out_ids = ws->t_ids.as<uint64_t>(); sa.out_dist = ws->t_dist.as<float>();
    sa.out_count = ws->t_cnt.as<uint32_t>(); sa.out_pos = ws->t_pos.as<uint64_t>();
    launch_select(sa, st);
    if (kp >= N) LGPU_CUDA(cudaMemsetAsync(ws->flags.p, 0, (size_t)B * 4, st));   // every row is a candidate
    else launch_band_check(ws->t_dist.as<float>(), ws->t_cnt.as<uint32_t>(), ws->qn2.as<float>(), xmax, d, B, k, kp,
                           ws->flags.as<uint32_t>(), st);
    launch_pair_distance(Q, X, ws->t_pos.as<uint64_t>(), B, kp, d, LGPU_L2, ws->t_exact.as<float>(), st);
    SelectArgs sb{};
    sb.mode = 2; sb.dense = ws->t_exact.as<float>(); sb.cand_ids = ws->t_ids.as<uint64_t>();
    sb.ncols = kp; sb.inner = kp; sb.row_stride = kp; sb.outer_stride = 0;
    sb.B = B; sb.k = k; sb.out_ids = out_ids; sb.out_dist = out_dist; sb.out_count = out_cnt;
    launch_select(sb, st);
    // fix-up pass (no-ops unless a query was flagged)
    launch_dist_matrix(Q, X, B, N, d, 0, nullptr, nullptr, Dbuf, ld, st, ws->flags.as<uint32_t>());
    SelectArgs sc{};
    sc.mode = 1; sc.dense = Dbuf; sc.ncols = N; sc.row_stride = ld; sc.col_ids = col_ids;
    sc.B = B; sc.k = k; sc.out_ids = out_ids; sc.out_dist = out_dist; sc.out_count = out_cnt;
    sc.only = ws->flags.as<uint32_t>();
    launch_select(sc, st);
}

// Large-N variant of tc_topk_l2 (flat search): a tensor-core pass over a row sample fixes, per query, a
// score threshold that provably admits every true top-k row; the pass over all rows then runs with the
// filtering epilogue (no dense score matrix), the few admitted rows are re-scored exactly, and queries whose
// candidate list overflowed are redone by the exact kernels.
static void tc_topk_l2_filtered(Workspace *ws, cudaStream_t st, int num_sms, const float *Q, uint32_t B, const float *X,
                                const void *Xb, const float *xnorm2, float xmax, uint64_t N, uint32_t d,
                                const uint64_t *col_ids, uint32_t k, uint64_t *out_ids, float *out_dist,
                                uint32_t *out_cnt, float *Dbuf, uint64_t ld, uint64_t min_sample = 65536,
                                uint32_t cap = 1024)
{
    const uint64_t Ns = std::min<uint64_t>(N, std::max<uint64_t>(min_sample, N / 8));
    const uint64_t lds = (Ns + 3) & ~3ull;               // Dbuf is [B][ld >= lds]
    ws->qb.ensure((size_t)B * d * 2); ws->qn2.ensure((size_t)B * 4); ws->flags.ensure((size_t)B * 4);
    ws->t_ids.ensure((size_t)B * cap * 8); ws->t_dist.ensure((size_t)B * std::max<uint32_t>(k, 32) * 4);
    ws->t_pos.ensure((size_t)B * cap * 8); ws->t_cnt.ensure((size_t)B * 4); ws->t_exact.ensure((size_t)B * cap * 4);
    ws->probe_A.ensure((size_t)B * 4);                   // thr[q]
    ws->amax.ensure((size_t)B * 4);                      // candidate counters
    ws->sbound.ensure((size_t)B * std::max<uint32_t>(k, 32) * 8);   // sample ids (unused)
    launch_to_bf16(Q, B, d, ws->qb.p, ws->qn2.as<float>(), st);
    // 1. sample pass: dense scores of the first Ns rows, k-th best per query -> threshold
    launch_gemm_dist(ws->qb.p, Xb, xnorm2, B, Ns, d, Dbuf, lds, num_sms, st);
    SelectArgs sa{};
    sa.mode = 1; sa.dense = Dbuf; sa.ncols = Ns; sa.row_stride = lds; sa.B = B; sa.k = k;
    sa.out_ids = ws->sbound.as<uint64_t>(); sa.out_dist = ws->t_dist.as<float>(); sa.out_count = ws->t_cnt.as<uint32_t>();
    launch_select(sa, st);
    launch_sample_threshold(ws->t_dist.as<float>(), ws->t_cnt.as<uint32_t>(), ws->qn2.as<float>(), xmax, d, B, k,
                            ws->probe_A.as<float>(), st);
    // 2. full pass with the filtering epilogue
    LGPU_CUDA(cudaMemsetAsync(ws->amax.p, 0, (size_t)B * 4, st));
    LGPU_CUDA(cudaMemsetAsync(ws->t_pos.p, 0xff, (size_t)B * cap * 8, st));
    LGPU_CUDA(cudaMemsetAsync(ws->t_ids.p, 0xff, (size_t)B * cap * 8, st));
    GemmFilter flt{};
    flt.thr = ws->probe_A.as<float>(); flt.count = ws->amax.as<uint32_t>(); flt.cand_pos = ws->t_pos.as<uint64_t>();
    flt.cand_ids = ws->t_ids.as<uint64_t>(); flt.col_ids = col_ids; flt.cap = cap;
    if (Ns == N) launch_filter_dense(Dbuf, lds, B, N, flt, st);      // the sample pass already scored every row
    else launch_gemm_dist(ws->qb.p, Xb, xnorm2, B, N, d, nullptr, 0, num_sms, st, &flt);
    launch_overflow_flags(ws->amax.as<uint32_t>(), cap, B, ws->flags.as<uint32_t>(), st);
    // 3. exact re-score of the admitted rows, final top-k.  (Launch shapes that skip the empty slots -- a loop
    // over the filled slots, 32- or 64-slot CTAs -- measured no faster at C2 and slower at C4 / C5 shapes.)
    launch_pair_distance(Q, X, ws->t_pos.as<uint64_t>(), B, cap, d, LGPU_L2, ws->t_exact.as<float>(), st);
    SelectArgs sb{};
    sb.mode = 2; sb.dense = ws->t_exact.as<float>(); sb.cand_ids = ws->t_ids.as<uint64_t>();
    sb.ncols = cap; sb.inner = cap; sb.row_stride = cap; sb.outer_stride = 0;
    sb.B = B; sb.k = k; sb.out_ids = out_ids; sb.out_dist = out_dist; sb.out_count = out_cnt;
    launch_select(sb, st);
    // 4. fix-up of overflowed queries (no-ops otherwise)
    launch_dist_matrix(Q, X, B, N, d, 0, nullptr, nullptr, Dbuf, ld, st, ws->flags.as<uint32_t>());
    SelectArgs sc{};
    sc.mode = 1; sc.dense = Dbuf; sc.ncols = N; sc.row_stride = ld; sc.col_ids = col_ids;
    sc.B = B; sc.k = k; sc.out_ids = out_ids; sc.out_dist = out_dist; sc.out_count = out_cnt;
    sc.only = ws->flags.as<uint32_t>();
    launch_select(sc, st);
}

void check_params(const lgpu_search_params *p)
{
    LGPU_REQUIRE(p != nullptr, "search params are null");
    LGPU_REQUIRE(p->k >= 1, "limit must be greater than 0");
    LGPU_REQUIRE(p->k <= SELECT_KMAX, "limit+offset above 2048 is not supported on the GPU path");
    if (p->refine_factor)
        LGPU_REQUIRE((uint64_t)p->k * p->refine_factor <= SELECT_KMAX, "limit*refine_factor above 2048 is not supported");
}

// one sub-batch of an IVF_PQ search, everything device-side on `st`
void ivf_sub_batch(lgpu_index *ix, Workspace *ws, cudaStream_t st, const float *d_q, uint32_t B,
                   const lgpu_search_params &sp, uint32_t nprobes, uint64_t *d_ids, float *d_dist,
                   uint32_t *d_cnt, bool prof, const uint64_t *forced_probes = nullptr, RowFilter rf = RowFilter(),
                   const uint32_t *only = nullptr)
{
    // `only` (device, [B]): redo just the flagged queries (maximum_nprobes widening) -- exact kernels, outputs of
    // the other queries are left untouched
    const uint32_t dim = ix->dim, nlist = ix->nlist;
    const uint32_t slots = B * nprobes;
    int evi = 0;
    auto mark = [&]() { if (prof) cudaEventRecord(ws->ev[evi++], st); };
    mark();
    ws->stats_mode = 0;
    // ---- queries (normalised copy for cosine) ----
    const float *qsearch = d_q;
    if (ix->metric == LGPU_COSINE) {
        ws->qn.ensure((size_t)B * dim * 4);
        launch_normalize(d_q, B, dim, ws->qn.as<float>(), st);
        qsearch = ws->qn.as<float>();
    }
    // ---- which scan: filter + verify (scan3.cu) unless the request needs every exact distance.  The filter's
    // per-query tables depend on the queries alone, so they are built on a side stream while the coarse step and the
    // regrouping (small kernels that leave most SMs idle) run on this one ----
    // the PQ top-`kk` of every query (kk = k, or k * refine_factor candidates for the exact re-rank)
    const uint32_t kk = sp.refine_factor ? sp.k * sp.refine_factor : sp.k;
    const uint32_t kp = kk <= 16 ? 32u : std::min<uint32_t>(SELECT_KMAX, 2 * kk + 32);
    // tiny batches (a single query, a micro-batch): one CTA per (query, probe) pair with the exact table in shared
    // memory (small.cu) -- 4 launches instead of ~25; LGPU_SMALL_SLOTS = 0 disables
    const ScanModes modes = scan_modes();
    const bool small_path = d_ids && !forced_probes && !only && slots <= modes.small_slots &&
                            small_scan_smem(ix->m, dim) <= 200 * 1024 &&
                            (size_t)slots * ix->pad_prefix[1] * 4 <= workspace_budget();
    const bool filter_scan = ix->has_tables && !modes.exact && !sp.has_lower && !sp.has_upper && !forced_probes &&
                             d_ids && kp > kk && ix->m <= 512 && !only && !small_path;
    if (filter_scan) {
        ws->qt.ensure((size_t)B * ix->nch * 256 * 16); ws->qt_mm.ensure((size_t)B * ix->nch * 8 * 8);
        ws->qt_step.ensure((size_t)B * 4); ws->qt_base.ensure((size_t)B * 4); ws->qt_bad.ensure((size_t)B * 4);
        ws->sbound.ensure((size_t)B * 4);
        LGPU_CUDA(cudaEventRecord(ws->ev_fork, st));
        LGPU_CUDA(cudaStreamWaitEvent(ws->aux, ws->ev_fork, 0));
        launch_query_tables_q16(qsearch, ix->cb_tiled.as<float>(), ix->cb_n2.as<float>(), B, dim, ix->m, ix->nch, ix->dsub,
                                ix->metric, ws->qt_mm.as<float>(), ws->qt.as<uint4>(), ws->qt_step.as<float>(),
                                ws->qt_base.as<float>(), ws->sbound.as<float>(), ws->qt_bad.as<uint32_t>(), ws->aux);
        LGPU_CUDA(cudaEventRecord(ws->ev_join, ws->aux));
        ws->aux_open = !g_capturing;
    }
    // ---- K1: exact centroid distances + nprobes nearest ----
    ws->probes.ensure((size_t)slots * 8);
    ws->probe_dist.ensure((size_t)slots * 4);
    ws->probe_cnt.ensure((size_t)B * 4);
    if (forced_probes) {
        LGPU_CUDA(cudaMemcpyAsync(ws->probes.p, forced_probes, (size_t)slots * 8, cudaMemcpyHostToDevice, st));
        mark();
    } else {
        const uint64_t ldc = (nlist + 3u) & ~3u;
        ws->D.ensure((size_t)B * ldc * 4);
        // the bf16 error band around the nprobes-th centroid has to fit in the shortlist, so take 3x
        // nprobes (>= 64) candidates (dense variant) or admit by threshold (filtered variant)
        const uint32_t kp = std::min<uint32_t>(SELECT_KMAX, std::max<uint32_t>(64, 3 * nprobes));
        // from ~1M (query, centroid) pairs the tensor-core shortlist wins (C2: 0.16 vs 0.19 ms)
        const bool big = (uint64_t)B * nlist >= ((uint64_t)1 << 20) || getenv("LGPU_FORCE_TC_COARSE");
        if (ix->has_tc && tc_enabled() && ix->metric != LGPU_DOT && B >= 8 && nlist >= 256 && kp > nprobes && big) {
            // tcgen05 GEMM scores + one finishing kernel per query (threshold, exact re-score in lance order, top
            // nprobes): bit-identical probe sets; queries whose candidate band overflowed are redone exactly
            mark();
            ws->qb.ensure((size_t)B * dim * 2); ws->qn2.ensure((size_t)B * 4); ws->flags.ensure((size_t)B * 4);
            ws->c_wcnt.ensure(16);
            uint32_t *cgate = ws->c_wcnt.as<uint32_t>() + 2;    // 0 = no query overflowed: the exact fix-up returns at once
            launch_to_bf16(qsearch, B, dim, ws->qb.p, ws->qn2.as<float>(), st);
            if (modes.coarse_list_min && nlist >= modes.coarse_list_min && ix->cent_ns >= 4 * nprobes && nprobes <= 64) {
                // Many lists (C5: 16384): a dense [B][nlist] score matrix is 537 MB written and read back, which bounds
                // the GEMM (tensor pipe 34 %).  Instead: (1) dense scores of a strided SAMPLE of the centroids; their
                // nprobes-th smallest + 2 E_q bounds, per query, the scores of every true probe; (2) the full GEMM runs
                // with the filtering epilogue and appends (column, score) of the few columns under that bound to the
                // query's list; (3) the finishing kernel works on the list (second-level threshold from the list's own
                // nprobes-th smallest, exact re-score, sort).  Overflowing lists are redone by the exact kernels.
                const uint32_t ns = ix->cent_ns, lds = (ns + 3u) & ~3u, lcap = 1024;
                ws->t_dist.ensure((size_t)B * nprobes * 4); ws->t_ids.ensure((size_t)B * nprobes * 8);
                ws->t_cnt.ensure((size_t)B * 4); ws->probe_A.ensure((size_t)B * 4); ws->amax.ensure((size_t)B * 4);
                ws->t_pos.ensure((size_t)B * lcap * 8); ws->t_exact.ensure((size_t)B * lcap * 4);
                launch_gemm_dist(ws->qb.p, ix->cent_sb.p, ix->cent_sn2.as<float>(), B, ns, dim, ws->D.as<float>(), lds,
                                 ix->num_sms, st);
                if (!launch_sample_kth_threshold(ws->D.as<float>(), lds, ns, ws->qn2.as<float>(), ix->cent_max, dim, B, nprobes,
                                                 ws->probe_A.as<float>(), st)) {
                    SelectArgs ss{};
                    ss.mode = 1; ss.dense = ws->D.as<float>(); ss.ncols = ns; ss.row_stride = lds; ss.B = B; ss.k = nprobes;
                    ss.out_ids = ws->t_ids.as<uint64_t>(); ss.out_dist = ws->t_dist.as<float>(); ss.out_count = ws->t_cnt.as<uint32_t>();
                    launch_select(ss, st);
                    launch_sample_threshold(ws->t_dist.as<float>(), ws->t_cnt.as<uint32_t>(), ws->qn2.as<float>(), ix->cent_max, dim,
                                            B, nprobes, ws->probe_A.as<float>(), st);
                }
                LGPU_CUDA(cudaMemsetAsync(ws->amax.p, 0, (size_t)B * 4, st));
                GemmFilter flt{};
                flt.thr = ws->probe_A.as<float>(); flt.count = ws->amax.as<uint32_t>(); flt.cand_pos = ws->t_pos.as<uint64_t>();
                flt.cand_ids = nullptr; flt.col_ids = nullptr; flt.cap = lcap; flt.cand_s = ws->t_exact.as<float>();
                launch_gemm_dist(ws->qb.p, ix->cent_b.p, ix->cent_n2.as<float>(), B, nlist, dim, nullptr, 0, ix->num_sms, st, &flt);
                launch_coarse_finish(ws->t_exact.as<float>(), lcap, B, lcap, qsearch, ix->centroids.as<float>(),
                                     ws->qn2.as<float>(), ix->cent_max, dim, nprobes, ws->probes.as<uint64_t>(),
                                     ws->probe_dist.as<float>(), ws->probe_cnt.as<uint32_t>(), ws->flags.as<uint32_t>(), cgate,
                                     st, ws->t_pos.as<uint64_t>(), ws->amax.as<uint32_t>());
            } else {
                launch_gemm_dist(ws->qb.p, ix->cent_b.p, ix->cent_n2.as<float>(), B, nlist, dim, ws->D.as<float>(), ldc,
                                 ix->num_sms, st);
                launch_coarse_finish(ws->D.as<float>(), ldc, B, nlist, qsearch, ix->centroids.as<float>(), ws->qn2.as<float>(),
                                     ix->cent_max, dim, nprobes, ws->probes.as<uint64_t>(), ws->probe_dist.as<float>(),
                                     ws->probe_cnt.as<uint32_t>(), ws->flags.as<uint32_t>(), cgate, st);
            }
            launch_dist_matrix(qsearch, ix->centroids.as<float>(), B, nlist, dim, 0, nullptr, nullptr, ws->D.as<float>(), ldc,
                               st, ws->flags.as<uint32_t>(), cgate);
            SelectArgs sc{};
            sc.mode = 1; sc.dense = ws->D.as<float>(); sc.ncols = nlist; sc.row_stride = ldc;
            sc.B = B; sc.k = nprobes; sc.out_ids = ws->probes.as<uint64_t>(); sc.out_dist = ws->probe_dist.as<float>();
            sc.out_count = ws->probe_cnt.as<uint32_t>(); sc.only = ws->flags.as<uint32_t>(); sc.gate = cgate;
            launch_select(sc, st);
        } else {
            launch_dist_matrix(qsearch, ix->centroids.as<float>(), B, nlist, dim, ix->metric == LGPU_DOT ? 1 : 0,
                               nullptr, nullptr, ws->D.as<float>(), ldc, st);
            mark();
            SelectArgs sa{};
            sa.mode = 1; sa.dense = ws->D.as<float>(); sa.ncols = nlist; sa.row_stride = ldc;
            sa.B = B; sa.k = nprobes;
            sa.out_ids = ws->probes.as<uint64_t>(); sa.out_dist = ws->probe_dist.as<float>();
            sa.out_count = ws->probe_cnt.as<uint32_t>();
            launch_select(sa, st);
        }
    }
    mark();
    if (small_path) {
        mark();                                          // (no regrouping)
        const uint64_t stride = std::max<uint64_t>(ix->pad_prefix[1], 4);
        ws->seg_off.ensure((size_t)slots * 8);
        ws->dist_out.ensure((size_t)slots * stride * 4);
        SmallScanArgs ss{};
        ss.centroids = ix->centroids.as<float>(); ss.cb_tiled = ix->cb_tiled.as<float>();
        ss.codes = ix->codes.as<unsigned char>(); ss.code_base = ix->code_base.as<uint64_t>();
        ss.part_n = ix->part_n.as<uint32_t>(); ss.part_npad = ix->part_npad.as<uint32_t>();
        ss.dim = dim; ss.m = ix->m; ss.nch = ix->nch; ss.metric = (uint32_t)ix->metric; ss.nlist = nlist; ss.nprobes = nprobes;
        ss.queries = qsearch; ss.probes = ws->probes.as<uint64_t>(); ss.seg_stride = stride;
        ss.seg_off = ws->seg_off.as<uint64_t>(); ss.dist_out = ws->dist_out.as<float>();
        launch_small_scan(ss, ix->dsub, slots, st);
        mark();
        SelectArgs sa{};
        sa.mode = 0; sa.dist = ws->dist_out.as<float>(); sa.seg_off = ws->seg_off.as<uint64_t>();
        sa.probes = ws->probes.as<uint64_t>(); sa.nprobes = nprobes; sa.nlist = nlist;
        sa.part_n = ix->part_n.as<uint32_t>(); sa.part_off = ix->part_off.as<uint64_t>();
        sa.row_ids = ix->row_ids.as<uint64_t>(); sa.B = B;
        sa.allow = rf.bits; sa.allow_bits = rf.nbits;
        sa.has_lower = sp.has_lower; sa.has_upper = sp.has_upper; sa.lower = sp.lower; sa.upper = sp.upper;
        sa.k = kk; sa.out_ids = d_ids; sa.out_dist = d_dist; sa.out_count = d_cnt;
        if (sp.refine_factor) {
            ws->t_ids.ensure((size_t)B * kk * 8); ws->t_dist.ensure((size_t)B * kk * 4);
            ws->t_pos.ensure((size_t)B * kk * 8); ws->t_cnt.ensure((size_t)B * 4); ws->t_exact.ensure((size_t)B * kk * 4);
            sa.out_ids = ws->t_ids.as<uint64_t>(); sa.out_dist = ws->t_dist.as<float>();
            sa.out_count = ws->t_cnt.as<uint32_t>(); sa.out_pos = ws->t_pos.as<uint64_t>();
        }
        launch_select(sa, st);
        mark();
        if (sp.refine_factor == 0) { mark(); return; }
        launch_pair_distance(d_q, ix->vectors.as<float>(), ws->t_pos.as<uint64_t>(), B, kk, dim, ix->metric,
                             ws->t_exact.as<float>(), st);
        SelectArgs sr{};
        sr.mode = 2; sr.dense = ws->t_exact.as<float>(); sr.cand_ids = ws->t_ids.as<uint64_t>();
        sr.ncols = kk; sr.inner = kk; sr.row_stride = kk; sr.outer_stride = 0;
        sr.B = B; sr.k = sp.k; sr.out_ids = d_ids; sr.out_dist = d_dist; sr.out_count = d_cnt;
        launch_select(sr, st);
        mark();
        return;
    }
    // ---- regroup probe slots by partition ----
    ws->part_cnt.ensure((size_t)nlist * 4);
    ws->slot_pos.ensure((size_t)slots * 4);
    ws->seg_local.ensure((size_t)slots * 8);
    ws->qtot.ensure((size_t)B * 8);
    ws->seg_off.ensure((size_t)slots * 8);
    ws->qlist_off.ensure((size_t)nlist * 4);
    ws->tile_off.ensure((size_t)(nlist + 1) * 4); ws->tile_off_b.ensure((size_t)(nlist + 1) * 4);
    ws->qlist.ensure((size_t)slots * 4);
    ws->scalars.ensure(64);
    GroupArgs ga{};
    ga.probes = ws->probes.as<uint64_t>(); ga.B = B; ga.nprobes = nprobes; ga.nlist = nlist;
    ga.part_n = ix->part_n.as<uint32_t>(); ga.part_cnt = ws->part_cnt.as<uint32_t>();
    ga.part_npad = ix->part_npad.as<uint32_t>(); ga.code_base = ix->code_base.as<uint64_t>(); ga.part_off = ix->part_off.as<uint64_t>();
    ga.slot_pos = ws->slot_pos.as<uint32_t>(); ga.seg_local = ws->seg_local.as<uint64_t>();
    ga.qtot = ws->qtot.as<uint64_t>(); ga.seg_off = ws->seg_off.as<uint64_t>();
    ga.qlist_off = ws->qlist_off.as<uint32_t>(); ga.tile_off = ws->tile_off.as<uint32_t>(); ga.tile_off_b = ws->tile_off_b.as<uint32_t>();
    ga.qlist = ws->qlist.as<uint32_t>();
    ga.total_tiles = ws->scalars.as<uint32_t>(); ga.tile_counter = ws->scalars.as<uint32_t>() + 1;
    ga.scanned_rows = reinterpret_cast<unsigned long long *>(ws->scalars.as<char>() + 16);
    // tile descriptors, sized by a host bound on the tile count (for 1536-row tiles; 3072-row tiles need fewer)
    // sum_p ceil(cnt_p / 8) * nrb_p <= (slots / 8 + #probed partitions) * max_p nrb_p
    {
        uint64_t max_tiles = ((uint64_t)slots / SCAN_G + std::min<uint64_t>(nlist, slots) + 1) * ix->max_nrb;
        LGPU_REQUIRE(max_tiles < (1ull << 31), "batch too large for one scan launch");
        ws->tile_desc.ensure((size_t)max_tiles * sizeof(TileDesc));
        ga.tile_desc = ws->tile_desc.as<TileDesc>(); ga.max_tiles = (uint32_t)max_tiles;
    }
    ga.only = only;
    ga.rows_tile = filter_scan ? SCAN3_ROWS_TILE : SCAN_ROWS_TILE_MID;
    launch_group(ga, st);
    mark();
    // ---- K2+K3 ----
    uint32_t np_eff = std::min<uint32_t>(nprobes, nlist);
    size_t cap_floats = (size_t)B * ix->pad_prefix[np_eff];
    ws->dist_out.ensure(std::max<size_t>(cap_floats, 4) * 4);
    ScanArgs sc{};
    sc.centroids = ix->centroids.as<float>(); sc.cb_tiled = ix->cb_tiled.as<float>();
    sc.codes = ix->codes.as<unsigned char>(); sc.code_base = ix->code_base.as<uint64_t>();
    sc.part_n = ix->part_n.as<uint32_t>(); sc.part_npad = ix->part_npad.as<uint32_t>();
    sc.dim = dim; sc.m = ix->m; sc.nch = ix->nch; sc.metric = (uint32_t)ix->metric; sc.nlist = nlist;
    sc.rows_tile = ga.rows_tile; sc.fzero2 = 0ull;
    sc.queries = qsearch;
    sc.total_tiles = ga.total_tiles; sc.tile_counter = ga.tile_counter;
    sc.dist_out = ws->dist_out.as<float>();
    sc.tile_desc = ga.tile_desc;
    sc.part_off = ix->part_off.as<uint64_t>();

    // K4 over the distance segments: the kk best by (_distance, _rowid), prefilter applied before the top-k
    SelectArgs sa{};
    sa.mode = 0; sa.dist = ws->dist_out.as<float>(); sa.seg_off = ga.seg_off; sa.probes = ga.probes;
    sa.nprobes = nprobes; sa.nlist = nlist; sa.part_n = ix->part_n.as<uint32_t>(); sa.part_off = ix->part_off.as<uint64_t>();
    sa.row_ids = ix->row_ids.as<uint64_t>(); sa.B = B;
    sa.allow = rf.bits; sa.allow_bits = rf.nbits;
    sa.only = only;
    // where the PQ top-kk goes: straight to the caller, or to the refine stage's candidate lists
    uint64_t *pq_ids = d_ids; float *pq_dist = d_dist; uint32_t *pq_cnt = d_cnt; uint64_t *pq_pos = nullptr;
    if (sp.refine_factor) {
        ws->t_ids.ensure((size_t)B * kk * 8); ws->t_dist.ensure((size_t)B * kk * 4);
        ws->t_pos.ensure((size_t)B * kk * 8); ws->t_cnt.ensure((size_t)B * 4);
        ws->t_exact.ensure((size_t)B * kk * 4);
        pq_ids = ws->t_ids.as<uint64_t>(); pq_dist = ws->t_dist.as<float>(); pq_cnt = ws->t_cnt.as<uint32_t>();
        pq_pos = ws->t_pos.as<uint64_t>();
    }

    if (filter_scan) {
        // per-query 16-bit tables, per-probe scalars
        const bool dot = ix->metric == LGPU_DOT;
        ws->flags.ensure((size_t)B * 4); ws->c_wcnt.ensure(16); ws->c_surv.ensure((size_t)B * 4);
        ws->s_ids.ensure((size_t)B * kp * 8); ws->s_lb.ensure((size_t)B * kp * 4); ws->s_pos.ensure((size_t)B * kp * 8);
        ws->s_cnt.ensure((size_t)B * 4); ws->s_exact.ensure((size_t)B * kp * 4);
        LGPU_CUDA(cudaStreamWaitEvent(st, ws->ev_join, 0));        // the tables, built on the side stream
        ws->aux_open = false;
        ws->qn2.ensure((size_t)B * 4);
        if (!dot) {
            ws->probe_A.ensure((size_t)slots * 4); ws->amax.ensure((size_t)B * 4);
            sc.probe_A = ws->probe_A.as<float>(); sc.row_R = ix->row_R.as<float>();
        }
        launch_probe_terms(ws->probe_dist.as<float>(), qsearch, B, nprobes, dim, dot ? nullptr : ws->probe_A.as<float>(),
                           dot ? nullptr : ws->amax.as<float>(), ws->qn2.as<float>(), st);
        sc.qt = ws->qt.as<uint4>(); sc.qt_step = ws->qt_step.as<float>(); sc.qt_base = ws->qt_base.as<float>();
        const float mscale = ix->metric == LGPU_COSINE ? 0.5f : 1.0f;
        // candidate mode (no prefilter, k <= 32): the scanners threshold the rows themselves, nothing dense is written
        // candidate capacity per query (power of two >= k); LGPU_CAND_CAP shrinks it to exercise the overflow path
        const uint32_t cap_env = modes.cap_env;
        uint32_t cap = kk <= 32 ? 512 : (kk <= 64 ? 1024 : CAND_CAP_MAX);
        {
            // long partitions put more rows inside the band around the k-th distance (clustered data: a query near a
            // big blob sees thousands of nearly equidistant rows): give them longer lists rather than the exact fix-up
            const uint64_t rows_probe = std::max<uint64_t>(1, ix->pad_prefix[np_eff] / np_eff);   // mean of the np largest
            if (rows_probe > 16384) cap = std::max<uint32_t>(cap, CAND_CAP_MAX);
            else if (rows_probe > 4096) cap = std::max<uint32_t>(cap, 1024);
        }
        bool cap_forced = false;
        if (cap_env >= 32 && cap_env <= CAND_CAP_MAX && !(cap_env & (cap_env - 1)) && cap_env >= kk) { cap = cap_env; cap_forced = true; }
        // A tile whose query has no threshold yet appends about k rows.  With few queries fanned out over many tiles
        // (small B, many probes, long partitions) most of a query's tiles run at the same moment on the 2 x SMs CTAs,
        // before any of them has published a threshold, and the list overflows whatever the scanners tighten later:
        // estimate that concurrency from the averages and take the dense mode (cheap at such B) when it is too high.
        bool cand_fits = true;
        if (!cap_forced) {
            const uint64_t rows_part = std::max<uint64_t>(1, ix->pad_prefix[np_eff] / np_eff);
            const uint64_t tiles_part = (rows_part + SCAN3_ROWS_TILE - 1) / SCAN3_ROWS_TILE;
            const uint64_t parts = std::min<uint64_t>(nlist, slots);
            const uint64_t groups_part = std::max<uint64_t>(1, (slots / parts + SCAN_G - 1) / SCAN_G);
            const double total = (double)parts * groups_part * tiles_part;
            const double conc = (double)np_eff * tiles_part * std::min(1.0, 2.0 * ix->num_sms / total);
            cand_fits = conc * kk <= 2.0 * cap;
        }
        const bool cand_mode = !rf.bits && kk <= modes.cand_kmax && !modes.dense_forced && cand_fits;
        if (cand_mode) {
            ws->c_thr.ensure((size_t)B * 4); ws->c_slack.ensure((size_t)B * 4); ws->c_cnt.ensure((size_t)B * 4);
            ws->c_rec.ensure((size_t)B * cap * sizeof(CandRec));
            ws->c_key.ensure((size_t)B * cap * 4); ws->c_last.ensure((size_t)B * 4);
            launch_cand_prepare(ws->qt_step.as<float>(), ws->sbound.as<float>(), dot ? nullptr : ws->amax.as<float>(),
                                dot ? nullptr : ix->rmax_bits.as<int>(), ws->qn2.as<float>(), ix->cb2, mscale, ix->m, B,
                                ws->c_slack.as<float>(), ws->c_thr.as<uint32_t>(), ws->c_cnt.as<uint32_t>(),
                                ws->c_last.as<uint32_t>(), ws->c_key.as<uint32_t>(), cap, st);
            sc.cand_key = ws->c_key.as<uint32_t>(); sc.cand_last = ws->c_last.as<uint32_t>();
            sc.nprobes = nprobes; sc.topk = kk; sc.thr = ws->c_thr.as<uint32_t>(); sc.slack = ws->c_slack.as<float>();
            sc.cand_cnt = ws->c_cnt.as<uint32_t>(); sc.cand = ws->c_rec.as<CandRec>(); sc.cand_cap = cap;
            launch_scan3(sc, ix->num_sms, st);
            mark();
            FinalizeArgs fa{};
            fa.Q = qsearch; fa.cand = sc.cand; fa.cand_cnt = sc.cand_cnt; fa.cand_cap = cap; fa.cand_key = sc.cand_key; fa.thr = sc.thr;
            fa.slack = sc.slack; fa.bad = ws->qt_bad.as<uint32_t>();
            fa.codes = ix->codes.as<unsigned char>(); fa.code_base = ix->code_base.as<uint64_t>();
            fa.part_npad = ix->part_npad.as<uint32_t>(); fa.part_off = ix->part_off.as<uint64_t>();
            fa.row_ids = ix->row_ids.as<uint64_t>(); fa.centroids = ix->centroids.as<float>();
            fa.cb_tiled = ix->cb_tiled.as<float>();
            fa.B = B; fa.dim = dim; fa.m = ix->m; fa.dsub = ix->dsub; fa.k = kk; fa.metric = ix->metric;
            fa.out_ids = pq_ids; fa.out_dist = pq_dist; fa.out_count = pq_cnt; fa.out_pos = pq_pos;
            fa.flags = ws->flags.as<uint32_t>();
            ws->c_work.ensure((size_t)B * cap * 8); ws->c_wcnt.ensure(16); ws->c_surv.ensure((size_t)B * 4);
            ws->c_exd.ensure((size_t)B * cap * 4); ws->c_exi.ensure((size_t)B * cap * 8); ws->c_exp.ensure((size_t)B * cap * 8);
            fa.work = ws->c_work.as<uint2>(); fa.work_cnt = ws->c_wcnt.as<uint32_t>(); fa.surv_cnt = ws->c_surv.as<uint32_t>();
            fa.ex_dist = ws->c_exd.as<float>(); fa.ex_id = ws->c_exi.as<uint64_t>(); fa.ex_pos = ws->c_exp.as<uint64_t>();
            fa.num_sms = ix->num_sms;
            if (prof) {
                ws->c_stats.ensure(32);
                LGPU_CUDA(cudaMemsetAsync(ws->c_stats.p, 0, 32, st));
                fa.stats = ws->c_stats.as<unsigned long long>();
                ws->stats_mode = 1;
            }
            launch_cand_finalize(fa, st);
            sc.cand = nullptr;                      // (the fix-up pass below is the exact kernel)
        } else {
        launch_scan3(sc, ix->num_sms, st);
        mark();
        // shortlist: the kp smallest lower bounds (with their storage positions)
        SelectArgs ss = sa;
        ss.k = kp; ss.out_ids = ws->s_ids.as<uint64_t>(); ss.out_dist = ws->s_lb.as<float>();
        ss.out_count = ws->s_cnt.as<uint32_t>(); ss.out_pos = ws->s_pos.as<uint64_t>();
        launch_select(ss, st);
        ws->stats_mode = 2;
        launch_band_check3(ws->s_lb.as<float>(), ws->s_cnt.as<uint32_t>(), ws->qt_step.as<float>(), ws->sbound.as<float>(),
                           dot ? nullptr : ws->amax.as<float>(), dot ? nullptr : ix->rmax_bits.as<int>(),
                           ws->qt_bad.as<uint32_t>(), ws->qn2.as<float>(), ix->cb2, mscale, ix->m, B, kk, kp,
                           ws->flags.as<uint32_t>(), ws->c_wcnt.as<uint32_t>() + 1, ws->c_surv.as<uint32_t>(), st);
        // exact PQ distances of the shortlist (oracle arithmetic), then the kk best of those
        launch_pq_rescore(qsearch, ws->s_pos.as<uint64_t>(), B, kp, ix->codes.as<unsigned char>(),
                          ix->code_base.as<uint64_t>(), ix->part_npad.as<uint32_t>(), ix->part_off.as<uint64_t>(), nlist,
                          ix->centroids.as<float>(), ix->cb_tiled.as<float>(), dim, ix->m, ix->dsub, ix->metric,
                          ws->c_surv.as<uint32_t>(), ws->s_exact.as<float>(), st);
        SelectArgs sb{};
        sb.mode = 2; sb.dense = ws->s_exact.as<float>(); sb.cand_ids = ws->s_ids.as<uint64_t>();
        sb.cand_pos = ws->s_pos.as<uint64_t>(); sb.ncols_q = ws->c_surv.as<uint32_t>();
        sb.ncols = kp; sb.inner = kp; sb.row_stride = kp; sb.outer_stride = 0;
        sb.B = B; sb.k = kk; sb.out_ids = pq_ids; sb.out_dist = pq_dist; sb.out_count = pq_cnt; sb.out_pos = pq_pos;
        launch_select(sb, st);
        }
        // fix-up of the queries whose shortlist could not be proven (no tiles, hence no work, unless one is
        // flagged): regroup them alone, exact scan, exact top-kk over their segments
        const uint32_t *gate = ws->c_wcnt.as<uint32_t>() + 1;   // 0 = nothing flagged: every kernel below returns at once
        ga.only = ws->flags.as<uint32_t>(); ga.gate = gate;
        ga.rows_tile = SCAN_ROWS_TILE_MID;
        launch_group(ga, st);
        sc.rows_tile = SCAN_ROWS_TILE_MID; sc.gate = gate;
        launch_scan2(sc, ix->dsub, ix->num_sms, st);
        SelectArgs sf = sa;
        sf.k = kk; sf.out_ids = pq_ids; sf.out_dist = pq_dist; sf.out_count = pq_cnt; sf.out_pos = pq_pos;
        sf.only = ws->flags.as<uint32_t>(); sf.gate = gate;
        launch_select(sf, st);
        mark();
    } else {
        launch_scan2(sc, ix->dsub, ix->num_sms, st);
        mark();
        if (!d_ids) { mark(); mark(); return; }     // debug: distances only
        sa.has_lower = sp.has_lower; sa.has_upper = sp.has_upper; sa.lower = sp.lower; sa.upper = sp.upper;
        sa.k = kk; sa.out_ids = pq_ids; sa.out_dist = pq_dist; sa.out_count = pq_cnt; sa.out_pos = pq_pos;
        launch_select(sa, st);
        mark();
    }
    if (sp.refine_factor == 0) { mark(); return; }
    // ---- refine (query.rs:1302-1332): exact distance of the k*rf candidates, re-sort ----
    launch_pair_distance(d_q, ix->vectors.as<float>(), ws->t_pos.as<uint64_t>(), B, kk, dim, ix->metric,
                         ws->t_exact.as<float>(), st);
    SelectArgs sr{};
    sr.mode = 2; sr.dense = ws->t_exact.as<float>(); sr.cand_ids = ws->t_ids.as<uint64_t>();
    sr.ncols = kk; sr.inner = kk; sr.row_stride = kk; sr.outer_stride = 0;
    sr.B = B; sr.k = sp.k; sr.out_ids = d_ids; sr.out_dist = d_dist; sr.out_count = d_cnt;
    sr.only = only;
    launch_select(sr, st);
    mark();
}

uint32_t ivf_sub_batch_size(lgpu_index *ix, uint32_t B, uint32_t nprobes)
{
    uint32_t np_eff = std::min<uint32_t>(nprobes, ix->nlist);
    size_t per_q = std::max<size_t>(ix->pad_prefix[np_eff] * 4 + (size_t)ix->nlist * 4 + (size_t)ix->nch * 256 * 8 * 4, 4);
    size_t bs = workspace_budget() / per_q;
    // tile descriptors address the distance segments with 32-bit float offsets
    bs = std::min<size_t>(bs, (size_t)0xffffffffull / std::max<size_t>(ix->pad_prefix[np_eff], 1));
    bs = std::max<size_t>(1, std::min<size_t>(bs, 65535));
    return (uint32_t)std::min<size_t>(bs, B);
}

void ivf_search_device(lgpu_index *ix, Workspace *ws, cudaStream_t st, const float *d_q, uint32_t B,
                       const lgpu_search_params &sp, uint64_t *d_ids, float *d_dist, uint32_t *d_cnt,
                       RowFilter rf = RowFilter(), const Deadline *deadline = nullptr)
{
    const uint32_t nprobes = std::min<uint32_t>(std::max<uint32_t>(sp.nprobes, 1), ix->nlist);
    const uint32_t np_widest = rf.bits ? std::max(nprobes, std::min<uint32_t>(sp.max_nprobes, ix->nlist)) : nprobes;
    const uint32_t bs = ivf_sub_batch_size(ix, B, np_widest);
    const bool prof = profiling_enabled();
    for (uint32_t q0 = 0; q0 < B; q0 += bs) {
        uint32_t b = std::min(bs, B - q0);
        if (deadline && q0 > 0) deadline->wait(st, ws->ev[7]);      // the previous sub-batch, or LGPU_TIMEOUT
        ivf_sub_batch(ix, ws, st, d_q + (size_t)q0 * ix->dim, b, sp, nprobes, d_ids + (size_t)q0 * sp.k,
                      d_dist + (size_t)q0 * sp.k, d_cnt + q0, prof && q0 == 0, nullptr, rf);
        // maximum_nprobes (query.rs:1250-1275): under a prefilter, the queries that found fewer than k rows in their
        // minimum_nprobes partitions are searched again over their maximum_nprobes nearest (no work if none is)
        const uint32_t np_max = std::min<uint32_t>(sp.max_nprobes, ix->nlist);
        if (rf.bits && np_max > nprobes) {
            LGPU_REQUIRE(np_max <= SELECT_KMAX || np_max >= ix->nlist, "maximum_nprobes above 2048 is not supported");
            ws->widen.ensure((size_t)b * 4);
            launch_count_below(d_cnt + q0, b, sp.k, ws->widen.as<uint32_t>(), st);
            ivf_sub_batch(ix, ws, st, d_q + (size_t)q0 * ix->dim, b, sp, np_max, d_ids + (size_t)q0 * sp.k,
                          d_dist + (size_t)q0 * sp.k, d_cnt + q0, false, nullptr, rf, ws->widen.as<uint32_t>());
        }
    }
    if (prof) {
        LGPU_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 6; i++) cudaEventElapsedTime(&g_stage_ms[i], ws->ev[i], ws->ev[i + 1]);
        cudaEventElapsedTime(&g_stage_ms[6], ws->ev[0], ws->ev[6]);
        unsigned long long rows = 0;
        LGPU_CUDA(cudaMemcpy(&rows, ws->scalars.as<char>() + 16, 8, cudaMemcpyDeviceToHost));
        g_scanned_bytes = (uint64_t)rows * ix->m;
        memset(g_filter_stats, 0, sizeof(g_filter_stats));
        if (ws->stats_mode == 1) LGPU_CUDA(cudaMemcpy(g_filter_stats, ws->c_stats.p, 32, cudaMemcpyDeviceToHost));
        else if (ws->stats_mode == 2) {                       // dense filter: queries the band check could not prove
            std::vector<uint32_t> fl(B);
            LGPU_CUDA(cudaMemcpy(fl.data(), ws->flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
            for (uint32_t f : fl) g_filter_stats[2] += f ? 1 : 0;
            g_filter_stats[3] = B;
        }
    }
}

void flat_search_device(lgpu_flat *fl, Workspace *ws, cudaStream_t st, int metric, const float *d_q, uint32_t B,
                        const lgpu_search_params &sp, uint64_t *d_ids, float *d_dist, uint32_t *d_cnt,
                        RowFilter rf = RowFilter(), const Deadline *deadline = nullptr)
{
    const uint64_t N = fl->nrows;
    const uint64_t ld = (N + 3) & ~3ull;
    if (metric == LGPU_COSINE) {
        std::lock_guard<std::mutex> g(fl->mu);
        if (!fl->has_norms) {
            fl->ysqrt.ensure(std::max<size_t>(N, 1) * 4);
            launch_row_norms(fl->vectors.as<float>(), N, fl->dim, fl->ysqrt.as<float>(), st);
            LGPU_CUDA(cudaStreamSynchronize(st));
            fl->has_norms = true;
        }
    }
    size_t per_q = std::max<size_t>(ld * 4, 4);
    uint32_t bs = (uint32_t)std::max<size_t>(1, std::min<size_t>(workspace_budget() / per_q, B));
    for (uint32_t q0 = 0; q0 < B; q0 += bs) {
        uint32_t b = std::min(bs, B - q0);
        if (deadline && q0 > 0) deadline->wait(st, ws->ev[7]);
        const float *q = d_q + (size_t)q0 * fl->dim;
        ws->D.ensure(std::max<size_t>((size_t)b * ld, 4) * 4);
        const float *xn = nullptr;
        if (metric == LGPU_COSINE) {
            ws->xnorm.ensure((size_t)b * 4);
            launch_row_norms(q, b, fl->dim, ws->xnorm.as<float>(), st);
            xn = ws->xnorm.as<float>();
        }
        const uint32_t kp = (uint32_t)std::min<uint64_t>(N, std::min<uint32_t>(SELECT_KMAX, std::max<uint32_t>(8 * sp.k, 256)));
        // (a prefilter goes through the exact kernels: the shortlist thresholds are fixed on unfiltered rows)
        if (fl->has_tc && tc_enabled() && metric == LGPU_L2 && !sp.has_lower && !sp.has_upper && b >= 8 && N >= 4096 &&
            !rf.bits) {
            if (N >= 262144 && !getenv("LGPU_FLAT_DENSE"))
                tc_topk_l2_filtered(ws, st, fl->num_sms, q, b, fl->vectors.as<float>(), fl->vec_b.p,
                                    fl->vec_n2.as<float>(), fl->vec_max, N, fl->dim,
                                    fl->has_ids ? fl->row_ids.as<uint64_t>() : nullptr, sp.k,
                                    d_ids + (size_t)q0 * sp.k, d_dist + (size_t)q0 * sp.k, d_cnt + q0,
                                    ws->D.as<float>(), ld);
            else
                tc_topk_l2(ws, st, fl->num_sms, q, b, fl->vectors.as<float>(), fl->vec_b.p, fl->vec_n2.as<float>(),
                           fl->vec_max, N, fl->dim, fl->has_ids ? fl->row_ids.as<uint64_t>() : nullptr, sp.k, kp,
                           d_ids + (size_t)q0 * sp.k, d_dist + (size_t)q0 * sp.k, d_cnt + q0, ws->D.as<float>(), ld);
            continue;
        }
        launch_dist_matrix(q, fl->vectors.as<float>(), b, N, fl->dim, metric == LGPU_L2 ? 0 : (metric == LGPU_DOT ? 1 : 2),
                           xn, fl->ysqrt.as<float>(), ws->D.as<float>(), ld, st);
        SelectArgs sa{};
        sa.mode = 1; sa.dense = ws->D.as<float>(); sa.ncols = N; sa.row_stride = ld;
        sa.col_ids = fl->has_ids ? fl->row_ids.as<uint64_t>() : nullptr;
        sa.B = b; sa.k = sp.k;
        sa.has_lower = sp.has_lower; sa.has_upper = sp.has_upper; sa.lower = sp.lower; sa.upper = sp.upper;
        sa.out_ids = d_ids + (size_t)q0 * sp.k; sa.out_dist = d_dist + (size_t)q0 * sp.k; sa.out_count = d_cnt + q0;
        sa.allow = rf.bits; sa.allow_bits = rf.nbits;
        launch_select(sa, st);
    }
}

template <class F> int guarded(F &&f)
{
    try { f(); return LGPU_OK; }
    catch (const Failure &e) { return e.status; }
    catch (const std::bad_alloc &) { set_error("host allocation failed"); return LGPU_OOM; }
    catch (const std::exception &e) { set_error(e.what()); return LGPU_RUNTIME; }
}

void require_device(int device)
{
    if (g_fork_poisoned.load()) {
        set_error("this process is a fork() of one that already used CUDA through lancedb_b200: the GPU context "
                  "does not survive fork; open the index in a spawned process instead");
        throw Failure{LGPU_RUNTIME};
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: lancedb_b200 has no CPU fallback");
        throw Failure{LGPU_RUNTIME};
    }
    LGPU_REQUIRE(device >= 0 && device < n, "invalid CUDA device ordinal");
    LGPU_CUDA(cudaSetDevice(device));
    g_cuda_touched.store(true);
}

// CUDA-graph replay of the host-buffer entry points is on by default; LGPU_NO_GRAPH=1 disables it.
static bool graphs_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("LGPU_NO_GRAPH"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// host-buffer wrapper: stage in, run, stage out, synchronise.  `key` identifies the launch sequence
// (shapes + parameters): the second call with the same key is captured into a CUDA graph and later calls
// replay it, which removes ~15 launch latencies from the synchronous end-to-end path.  Any device
// (re)allocation anywhere in the process since the capture, profiling mode, or a failed capture falls back to
// eager launches.
template <class Run>
void host_submit(WsLease &lease, const Deadline &deadline, const float *queries, uint32_t B, uint32_t dim, uint32_t k,
                 uint64_t *out_ids, float *out_dist, uint32_t *out_count, const uint64_t (&key)[4], Run &&run,
                 bool allow_graph, bool sync)
{
    Workspace *ws = lease.ws;
    cudaStream_t st = lease.st;
    ws->q.ensure(std::max<size_t>((size_t)B * dim, 1) * 4);
    ws->out_ids.ensure(std::max<size_t>((size_t)B * k, 1) * 8);
    ws->out_dist.ensure(std::max<size_t>((size_t)B * k, 1) * 4);
    ws->out_count.ensure(std::max<size_t>(B, 1) * 4);
    LGPU_CUDA(cudaMemcpyAsync(ws->q.p, queries, (size_t)B * dim * 4, cudaMemcpyHostToDevice, st));
    auto eager = [&] {
        run(ws, st, ws->q.as<float>(), ws->out_ids.as<uint64_t>(), ws->out_dist.as<float>(), ws->out_count.as<uint32_t>(),
            deadline);
    };
    const bool same = memcmp(key, ws->graph_key, sizeof(key)) == 0;
    if (!allow_graph || !graphs_enabled() || profiling_enabled() || ws->graph_state < 0 || deadline.armed) {
        eager();
    } else if (!same) {                                     // new shape: warm up (allocations), capture next time
        if (ws->graph) { cudaGraphExecDestroy(ws->graph); ws->graph = nullptr; }
        memcpy(ws->graph_key, key, sizeof(key));
        ws->graph_state = 0;
        eager();
        ws->graph_state = 1;
    } else if (ws->graph_state == 2 && g_alloc_epoch.load() == ws->graph_epoch) {
        LGPU_CUDA(cudaGraphLaunch(ws->graph, st));
        count_launches(ws->graph_kernels);
    } else if (ws->graph_state == 1) {
        bool ok = false;
        cudaGraph_t g = nullptr;
        const uint64_t epoch0 = g_alloc_epoch.load();
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            g_capturing = true; g_captured_launches = 0;
            try {
                eager();
                ok = cudaStreamEndCapture(st, &g) == cudaSuccess && g != nullptr && g_alloc_epoch.load() == epoch0;
            } catch (const Failure &) {
                cudaStreamEndCapture(st, &g);
                ok = false;
            }
            g_capturing = false;
        }
        if (ok && cudaGraphInstantiate(&ws->graph, g, 0) == cudaSuccess) {
            ws->graph_state = 2;
            ws->graph_epoch = epoch0;
            ws->graph_kernels = g_captured_launches;
            cudaGraphDestroy(g);
            LGPU_CUDA(cudaGraphLaunch(ws->graph, st));
            count_launches(ws->graph_kernels);
        } else if (g_alloc_epoch.load() != epoch0) {        // a buffer moved during the capture: try again next call
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            ws->graph = nullptr;
            LGPU_CUDA(cudaMemcpyAsync(ws->q.p, queries, (size_t)B * dim * 4, cudaMemcpyHostToDevice, st));
            eager();
        } else {                                            // never try again with this workspace
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            ws->graph = nullptr;
            ws->graph_state = -1;
            LGPU_CUDA(cudaMemcpyAsync(ws->q.p, queries, (size_t)B * dim * 4, cudaMemcpyHostToDevice, st));
            eager();
        }
    } else {                                                // captured, but a buffer moved since: capture again
        if (ws->graph) { cudaGraphExecDestroy(ws->graph); ws->graph = nullptr; }
        ws->graph_state = 1;
        eager();
    }
    if (sync) deadline.wait(st, ws->ev[7]);
    LGPU_CUDA(cudaMemcpyAsync(out_ids, ws->out_ids.p, (size_t)B * k * 8, cudaMemcpyDeviceToHost, st));
    LGPU_CUDA(cudaMemcpyAsync(out_dist, ws->out_dist.p, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
    LGPU_CUDA(cudaMemcpyAsync(out_count, ws->out_count.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (sync) LGPU_CUDA(cudaStreamSynchronize(st));
}

template <class Run>
void host_call(WorkspacePool &pool, const float *queries, uint32_t B, uint32_t dim, uint32_t k, uint64_t *out_ids,
               float *out_dist, uint32_t *out_count, const uint64_t (&key)[4], uint32_t timeout_ms, Run &&run,
               bool allow_graph = true)
{
    const Deadline deadline(timeout_ms);
    WsLease lease(pool, nullptr, false);
    host_submit(lease, deadline, queries, B, dim, k, out_ids, out_dist, out_count, key, run, allow_graph, true);
}

static inline void make_key(uint64_t (&key)[4], uint64_t tag, uint32_t B, const lgpu_search_params &p)
{
    uint32_t lo, hi;
    memcpy(&lo, &p.lower, 4); memcpy(&hi, &p.upper, 4);
    key[0] = tag ^ ((uint64_t)B << 32);
    key[1] = (uint64_t)p.k | ((uint64_t)p.nprobes << 32);
    key[2] = (uint64_t)p.refine_factor | ((uint64_t)(p.has_lower ? 1 : 0) << 32) | ((uint64_t)(p.has_upper ? 1 : 0) << 33) |
             ((uint64_t)(p.max_nprobes & 0x3fffffu) << 34);
    key[3] = ((uint64_t)lo | ((uint64_t)hi << 32)) ^ scan_modes().signature();
}

}  // namespace

extern "C" {

const char *lgpu_last_error(void) { return g_err.c_str(); }
uint32_t lgpu_abi_version(void) { return LGPU_ABI_VERSION; }

int lgpu_device_count(int *count)
{
    return guarded([&] {
        LGPU_REQUIRE(count != nullptr, "count is null");
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
        *count = n;
    });
}

int lgpu_index_open(const lgpu_index_desc *d, lgpu_index **out)
{
    lgpu_index *ix = nullptr;
    int rc = guarded([&] {
        LGPU_REQUIRE(d != nullptr && out != nullptr, "null argument");
        LGPU_REQUIRE(d->abi_version == LGPU_ABI_VERSION, "ABI version mismatch");
        LGPU_REQUIRE(d->dim > 0 && d->nlist > 0 && d->m > 0, "dim, nlist and m must be positive");
        LGPU_REQUIRE(d->dim % d->m == 0, "num_sub_vectors must divide the vector dimension");
        LGPU_REQUIRE(d->nbits == 8, "only 8-bit PQ codes are supported");
        LGPU_REQUIRE(d->metric == LGPU_L2 || d->metric == LGPU_COSINE || d->metric == LGPU_DOT, "unknown distance type");
        LGPU_REQUIRE(d->codes_layout == LGPU_CODES_ROW_MAJOR || d->codes_layout == LGPU_CODES_PARTITION_TRANSPOSED,
                     "unknown codes layout");
        LGPU_REQUIRE(scan_dsub_supported(d->dim / d->m),
                     "unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        LGPU_REQUIRE(d->centroids && d->codebook && d->part_offsets, "null index array");
        LGPU_REQUIRE(d->nrows == 0 || (d->codes && d->row_ids), "null codes / row_ids");
        LGPU_REQUIRE(d->nrows < (1ull << 32), "more than 2^32 rows in one GPU shard are not supported (shard the index)");
        LGPU_REQUIRE(d->part_offsets[0] == 0 && d->part_offsets[d->nlist] == d->nrows,
                     "part_offsets must start at 0 and end at nrows");
        for (uint32_t p = 0; p < d->nlist; p++) {
            LGPU_REQUIRE(d->part_offsets[p + 1] >= d->part_offsets[p], "part_offsets must be non-decreasing");
            LGPU_REQUIRE(d->part_offsets[p + 1] - d->part_offsets[p] < (1ull << 31), "partition too large");
        }
        require_device(d->device);
        ix = new lgpu_index();
        ix->device = d->device;
        cudaDeviceProp prop;
        LGPU_CUDA(cudaGetDeviceProperties(&prop, d->device));
        ix->num_sms = prop.multiProcessorCount;
        ix->dim = d->dim; ix->nlist = d->nlist; ix->m = d->m; ix->dsub = d->dim / d->m;
        ix->nch = (d->m + 7) / 8; ix->metric = d->metric; ix->nrows = d->nrows;

        const uint32_t nlist = d->nlist;
        std::vector<uint32_t> part_n(nlist), part_npad(nlist);
        std::vector<uint64_t> code_base(nlist), pads(nlist);
        uint64_t cb = 0;
        for (uint32_t p = 0; p < nlist; p++) {
            uint32_t n = (uint32_t)(d->part_offsets[p + 1] - d->part_offsets[p]);
            part_n[p] = n; part_npad[p] = (n + 31u) & ~31u;
            code_base[p] = cb;
            cb += (uint64_t)(ix->nch + 1) * part_npad[p] * 8;
            LGPU_REQUIRE((cb >> 3) < (1ull << 32), "index too large for one GPU shard (re-laid-out codes above 32 GiB)");
            pads[p] = (n + 3ull) & ~3ull;
        }
        ix->h_part_n = part_n;
        for (uint32_t p = 0; p < nlist; p++) ix->max_nrb = std::max(ix->max_nrb, scan_nrb(part_n[p], SCAN_ROWS_TILE_MID));
        std::sort(pads.begin(), pads.end(), std::greater<uint64_t>());
        ix->pad_prefix.assign(nlist + 1, 0);
        for (uint32_t p = 0; p < nlist; p++) ix->pad_prefix[p + 1] = ix->pad_prefix[p] + pads[p];

        cudaStream_t st = nullptr;
        auto up = [&](DevBuf &b, const void *src, size_t bytes) {
            b.ensure(std::max<size_t>(bytes, 16));
            if (bytes) LGPU_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
            ix->device_bytes += b.bytes;
        };
        up(ix->centroids, d->centroids, (size_t)nlist * d->dim * 4);
        up(ix->part_n, part_n.data(), (size_t)nlist * 4);
        up(ix->part_npad, part_npad.data(), (size_t)nlist * 4);
        up(ix->code_base, code_base.data(), (size_t)nlist * 8);
        up(ix->part_off, d->part_offsets, (size_t)(nlist + 1) * 8);
        up(ix->row_ids, d->row_ids, (size_t)d->nrows * 8);
        if (d->vectors) { up(ix->vectors, d->vectors, (size_t)d->nrows * d->dim * 4); ix->has_vectors = true; }
        if (gemm_shape_supported(d->dim) && d->metric != LGPU_DOT) {
            LGPU_CUDA(cudaStreamSynchronize(st));
            prepare_tc_operand(ix->centroids.as<float>(), nlist, d->dim, ix->cent_b, ix->cent_n2, ix->cent_max, st);
            ix->device_bytes += ix->cent_b.bytes + ix->cent_n2.bytes;
            if (nlist >= 1024) {
                // every 8th centroid, for the sampled threshold of the coarse step (strided: whatever order the
                // trainer left the lists in -- hierarchical k-means groups neighbours -- the sample spans all of them)
                constexpr uint32_t COARSE_SAMPLE_STRIDE = 8;
                const uint32_t ns = nlist / COARSE_SAMPLE_STRIDE;
                ix->cent_sb.ensure((size_t)ns * d->dim * 2); ix->cent_sn2.ensure((size_t)ns * 4);
                LGPU_CUDA(cudaMemcpy2DAsync(ix->cent_sb.p, (size_t)d->dim * 2, ix->cent_b.p, (size_t)COARSE_SAMPLE_STRIDE * d->dim * 2,
                                            (size_t)d->dim * 2, ns, cudaMemcpyDeviceToDevice, st));
                LGPU_CUDA(cudaMemcpy2DAsync(ix->cent_sn2.p, 4, ix->cent_n2.p, (size_t)COARSE_SAMPLE_STRIDE * 4, 4, ns,
                                            cudaMemcpyDeviceToDevice, st));
                LGPU_CUDA(cudaStreamSynchronize(st));
                ix->cent_ns = ns;
                ix->device_bytes += ix->cent_sb.bytes + ix->cent_sn2.bytes;
            }
            ix->has_tc = true;
        }
        // codebook -> [nch][256][8][dsub]
        {
            DevBuf tmp;
            size_t bytes = (size_t)d->m * 256 * ix->dsub * 4;
            tmp.ensure(bytes);
            LGPU_CUDA(cudaMemcpyAsync(tmp.p, d->codebook, bytes, cudaMemcpyHostToDevice, st));
            ix->cb_tiled.ensure((size_t)ix->nch * 256 * 8 * ix->dsub * 4);
            ix->device_bytes += ix->cb_tiled.bytes;
            launch_retile_codebook(tmp.as<float>(), d->m, ix->dsub, ix->nch, ix->cb_tiled.as<float>(), st);
            LGPU_CUDA(cudaStreamSynchronize(st));
        }
        // codes -> skewed streams
        {
            ix->codes.ensure(std::max<uint64_t>(cb, 16));
            ix->device_bytes += ix->codes.bytes;
            LGPU_CUDA(cudaMemsetAsync(ix->codes.p, 0, ix->codes.bytes, st));
            DevBuf tmp;
            size_t bytes = (size_t)d->nrows * d->m;
            if (bytes) {
                tmp.ensure(bytes);
                LGPU_CUDA(cudaMemcpyAsync(tmp.p, d->codes, bytes, cudaMemcpyHostToDevice, st));
                launch_retile_codes(tmp.as<unsigned char>(), d->codes_layout, ix->part_off.as<uint64_t>(), nlist,
                                    d->nrows, d->m, ix->nch, ix->code_base.as<uint64_t>(),
                                    ix->part_npad.as<uint32_t>(), ix->codes.as<unsigned char>(), st);
            }
            LGPU_CUDA(cudaStreamSynchronize(st));
        }
        {   // |b|^2 of the tiled codebook entries and CB2 = sum_i max_c |b|^2 (filter scan, tables.cu)
            const size_t ne = (size_t)ix->nch * 256 * 8;
            ix->cb_n2.ensure(ne * 4);
            ix->device_bytes += ix->cb_n2.bytes;
            launch_cb_norms(ix->cb_tiled.as<float>(), ix->nch, ix->dsub, ix->cb_n2.as<float>(), st);
            double cb2 = 0.0;
            for (uint32_t i = 0; i < d->m; i++) {
                double mx = 0.0;
                for (uint32_t c = 0; c < 256; c++) {
                    double n2 = 0.0;
                    const float *e = d->codebook + ((size_t)i * 256 + c) * ix->dsub;
                    for (uint32_t t = 0; t < ix->dsub; t++) n2 += (double)e[t] * e[t];
                    mx = std::max(mx, n2);
                }
                cb2 += mx;
            }
            ix->cb2 = (float)(cb2 * 1.000001);
        }
        if (d->metric != LGPU_DOT) {   // per-row constants of the filter scan (tables.cu)
            ix->row_R.ensure(std::max<size_t>((size_t)d->nrows * 4, 16));
            ix->rmax_bits.ensure(16);
            ix->device_bytes += ix->row_R.bytes;
            launch_row_const(ix->codes.as<unsigned char>(), ix->code_base.as<uint64_t>(), ix->part_npad.as<uint32_t>(),
                             ix->part_off.as<uint64_t>(), nlist, d->nrows, ix->centroids.as<float>(),
                             ix->cb_tiled.as<float>(), d->dim, d->m, ix->dsub, ix->row_R.as<float>(),
                             ix->rmax_bits.as<int>(), st);
            LGPU_CUDA(cudaStreamSynchronize(st));
        }
        ix->has_tables = true;
        register_handle(ix);
        *out = ix;
    });
    if (rc != LGPU_OK && ix) delete ix;
    return rc;
}

void lgpu_index_close(lgpu_index *ix)
{
    if (!retire_handle(ix)) return;          // null, already closed, or a parent-process handle in a fork child
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();
    delete ix;
}

int lgpu_index_device_bytes(const lgpu_index *ixh, uint64_t *bytes)
{
    return guarded([&] {
        LGPU_REQUIRE(bytes, "null argument");
        HandleRef<lgpu_index> ix(const_cast<lgpu_index *>(ixh), "index");
        *bytes = ix->device_bytes;
    });
}

int lgpu_last_scanned_code_bytes(uint64_t *bytes)
{
    return guarded([&] { LGPU_REQUIRE(bytes, "null argument"); *bytes = g_scanned_bytes; });
}

int lgpu_kernel_launch_count(uint64_t *count)
{
    return guarded([&] { LGPU_REQUIRE(count, "null argument"); *count = g_kernel_launches.load(); });
}

int lgpu_last_filter_stats(uint64_t *stats)
{
    return guarded([&] { LGPU_REQUIRE(stats, "null argument"); memcpy(stats, g_filter_stats, sizeof(g_filter_stats)); });
}

int lgpu_set_profiling(int enabled)
{
    g_profiling.store(enabled ? 1 : 0);
    return LGPU_OK;
}

int lgpu_last_stage_ms(float *times)
{
    return guarded([&] { LGPU_REQUIRE(times, "null argument"); memcpy(times, g_stage_ms, sizeof(g_stage_ms)); });
}

static void check_ivf_call(lgpu_index *ix, const void *q, uint32_t B, const lgpu_search_params *p,
                           const void *a, const void *b, const void *c)
{
    check_params(p);
    LGPU_REQUIRE(p->nprobes >= 1, "minimum_nprobes must be greater than 0");
    LGPU_REQUIRE(p->nprobes <= SELECT_KMAX || p->nprobes >= ix->nlist, "nprobes above 2048 is not supported");
    LGPU_REQUIRE(B == 0 || (q && a && b && c), "null buffer");
    LGPU_REQUIRE(p->refine_factor == 0 || ix->has_vectors,
                 "refine_factor needs the raw vectors: open the index with desc.vectors");
}

int lgpu_search(lgpu_index *ixh, const float *queries, uint32_t B, const lgpu_search_params *params,
                uint64_t *out_ids, float *out_dist, uint32_t *out_count)
{
    return guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        check_ivf_call(ix.h, queries, B, params, out_ids, out_dist, out_count);
        if (B == 0) return;
        require_device(ix->device);
        uint64_t key[4];
        make_key(key, 0x1f5ull, B, *params);
        host_call(ix->pool, queries, B, ix->dim, params->k, out_ids, out_dist, out_count, key, params->timeout_ms,
                  [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                      const Deadline &dl) { ivf_search_device(ix.h, ws, st, dq, B, *params, di, dd, dc, RowFilter(), &dl); });
    });
}

int lgpu_search_filtered(lgpu_index *ixh, const float *queries, uint32_t B, const lgpu_search_params *params,
                         const uint32_t *allow, uint64_t allow_bits, uint64_t *out_ids, float *out_dist,
                         uint32_t *out_count)
{
    return guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        check_ivf_call(ix.h, queries, B, params, out_ids, out_dist, out_count);
        LGPU_REQUIRE(allow != nullptr || allow_bits == 0, "allow bitmap is null");
        if (B == 0) return;
        require_device(ix->device);
        uint64_t key[4];
        make_key(key, 0x1f6ull, B, *params);
        key[0] ^= allow_bits << 8;
        host_call(ix->pool, queries, B, ix->dim, params->k, out_ids, out_dist, out_count, key, params->timeout_ms,
                  [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                      const Deadline &dl) {
                      const size_t words = (size_t)((allow_bits + 31) / 32);
                      ws->allow.ensure(std::max<size_t>(words, 1) * 4);
                      if (words) LGPU_CUDA(cudaMemcpyAsync(ws->allow.p, allow, words * 4, cudaMemcpyHostToDevice, st));
                      RowFilter rf; rf.bits = ws->allow.as<uint32_t>(); rf.nbits = allow_bits;
                      ivf_search_device(ix.h, ws, st, dq, B, *params, di, dd, dc, rf, &dl);
                  }, false);
    });
}

int lgpu_search_coalesced(lgpu_index *ixh, const float *query, const lgpu_search_params *params, uint64_t *out_ids,
                          float *out_dist, uint32_t *out_count)
{
    static const uint32_t window_us = getenv("LGPU_COALESCE_US") ? (uint32_t)atoi(getenv("LGPU_COALESCE_US")) : 50u;
    constexpr size_t MAX_BATCH = 256;
    PendingQuery me{query, out_ids, out_dist, out_count};
    std::vector<PendingQuery *> batch;
    lgpu_search_params p{};
    int rc = guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        check_ivf_call(ix.h, query, 1, params, out_ids, out_dist, out_count);
        p = *params;
        Coalescer &co = ix->coalescer;
        std::unique_lock<std::mutex> lk(co.mu);
        Coalescer::Lane *lane = nullptr;
        for (auto *l : co.lanes) if (memcmp(&l->params, &p, sizeof(p)) == 0) { lane = l; break; }
        if (!lane) { lane = new Coalescer::Lane(); lane->params = p; co.lanes.push_back(lane); }
        lane->queue.push_back(&me);
        if (lane->leader) {                               // follower: the lane's leader will fill our row in
            if (lane->queue.size() >= MAX_BATCH) co.cv.notify_all();
            co.cv.wait(lk, [&] { return me.done; });
            return;
        }
        lane->leader = true;                              // leader: collect for one window, then search the batch
        co.cv.wait_for(lk, std::chrono::microseconds(window_us), [&] { return lane->queue.size() >= MAX_BATCH; });
        batch.swap(lane->queue);
        lane->leader = false;
        lk.unlock();
        const uint32_t B = (uint32_t)batch.size();
        const uint32_t dim = ix->dim, k = p.k;
        std::vector<float> q((size_t)B * dim);
        std::vector<uint64_t> ids((size_t)B * k);
        std::vector<float> dist((size_t)B * k);
        std::vector<uint32_t> cnt(B);
        for (uint32_t i = 0; i < B; i++) memcpy(q.data() + (size_t)i * dim, batch[i]->q, (size_t)dim * 4);
        const int brc = lgpu_search(ix.h, q.data(), B, &p, ids.data(), dist.data(), cnt.data());
        const std::string berr = brc == LGPU_OK ? std::string() : std::string(lgpu_last_error());
        lk.lock();
        for (uint32_t i = 0; i < B; i++) {
            PendingQuery *pq = batch[i];
            pq->status = brc; pq->err = berr;
            if (brc == LGPU_OK) {
                memcpy(pq->ids, ids.data() + (size_t)i * k, (size_t)k * 8);
                memcpy(pq->dist, dist.data() + (size_t)i * k, (size_t)k * 4);
                *pq->cnt = cnt[i];
            }
            pq->done = true;
        }
        lk.unlock();
        co.cv.notify_all();
    });
    if (rc != LGPU_OK) return rc;
    if (me.status != LGPU_OK) { set_error(me.err); return me.status; }
    return LGPU_OK;
}

int lgpu_search_device(lgpu_index *ixh, const float *d_queries, uint32_t B, const lgpu_search_params *params,
                       uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count, void *cuda_stream)
{
    return guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        check_ivf_call(ix.h, d_queries, B, params, d_out_ids, d_out_dist, d_out_count);
        if (B == 0) return;
        require_device(ix->device);
        WsLease lease(ix->pool, (cudaStream_t)cuda_stream, true);
        ivf_search_device(ix.h, lease.ws, lease.st, d_queries, B, *params, d_out_ids, d_out_dist, d_out_count);
    });
}

// ---- asynchronous completion (SURVEY.md 8b "Threading": a tokio worker must not be blocked for the length of a
// search -- python/src/runtime.rs:113-119 uses spawn_blocking for that today).  lgpu_search_async stages the
// queries, enqueues the search and the copy-back on a private stream and returns; lgpu_ticket_wait blocks until
// the results are in the caller's buffers.  Two tickets in flight use two workspaces/streams, so batch i+1's
// H2D overlaps batch i's kernels. ----
struct lgpu_ticket {
    HandleRef<lgpu_index> ix;
    WsLease lease;
    Deadline deadline;
    cudaEvent_t done = nullptr;
    lgpu_ticket(lgpu_index *h, uint32_t timeout_ms)
        : ix(h, "index"), lease(ix->pool, nullptr, false), deadline(timeout_ms) {}
    ~lgpu_ticket() { if (done) cudaEventDestroy(done); }
};

int lgpu_search_async(lgpu_index *ixh, const float *queries, uint32_t B, const lgpu_search_params *params,
                      uint64_t *out_ids, float *out_dist, uint32_t *out_count, lgpu_ticket **ticket)
{
    lgpu_ticket *t = nullptr;
    int rc = guarded([&] {
        LGPU_REQUIRE(ticket != nullptr, "ticket is null");
        LGPU_REQUIRE(params != nullptr, "search params are null");
        {   // validate before taking a workspace
            HandleRef<lgpu_index> ix(ixh, "index");
            check_ivf_call(ix.h, queries, B, params, out_ids, out_dist, out_count);
            require_device(ix->device);
        }
        t = new lgpu_ticket(ixh, params->timeout_ms);
        LGPU_CUDA(cudaEventCreateWithFlags(&t->done, cudaEventDisableTiming));
        if (B > 0) {
            uint64_t key[4];
            make_key(key, 0x1f5ull, B, *params);
            lgpu_index *ix = t->ix.h;
            const lgpu_search_params sp = *params;
            host_submit(t->lease, t->deadline, queries, B, ix->dim, sp.k, out_ids, out_dist, out_count, key,
                        [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                            const Deadline &) { ivf_search_device(ix, ws, st, dq, B, sp, di, dd, dc); },
                        true, false);
        }
        LGPU_CUDA(cudaEventRecord(t->done, t->lease.st));
        *ticket = t;
    });
    if (rc != LGPU_OK && t) { cudaStreamSynchronize(t->lease.st); delete t; }
    return rc;
}

int lgpu_ticket_poll(lgpu_ticket *t, int *done)
{
    return guarded([&] {
        LGPU_REQUIRE(t && done, "null argument");
        cudaError_t e = cudaEventQuery(t->done);
        if (e != cudaSuccess && e != cudaErrorNotReady) LGPU_CUDA(e);
        *done = e == cudaSuccess ? 1 : 0;
    });
}

/* blocks until the call's results are in the caller's buffers, then frees the ticket.  With a timeout armed the
 * status is LGPU_TIMEOUT when the deadline passed first -- the wait still runs to completion, because a DMA into
 * the caller's buffers may not be left in flight. */
int lgpu_ticket_wait(lgpu_ticket *t)
{
    if (!t) { set_error("ticket is null"); return LGPU_INVALID_INPUT; }
    int rc = guarded([&] {
        bool late = false;
        if (t->deadline.armed) {
            for (;;) {
                cudaError_t e = cudaEventQuery(t->done);
                if (e == cudaSuccess) break;
                if (e != cudaErrorNotReady) LGPU_CUDA(e);
                if (t->deadline.expired()) { late = true; break; }
                std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
        }
        LGPU_CUDA(cudaEventSynchronize(t->done));
        if (late) { set_error("Query timeout"); throw Failure{LGPU_TIMEOUT}; }
    });
    delete t;
    return rc;
}

int lgpu_merge_topk_device(int device, uint32_t nlists, uint32_t B, uint32_t k, const uint64_t *d_ids,
                           const float *d_dist, uint64_t *d_out_ids, float *d_out_dist, uint32_t *d_out_count,
                           void *cuda_stream)
{
    return guarded([&] {
        LGPU_REQUIRE(nlists >= 1 && k >= 1 && k <= SELECT_KMAX, "bad merge shape");
        LGPU_REQUIRE(B == 0 || (d_ids && d_dist && d_out_ids && d_out_dist && d_out_count), "null buffer");
        if (B == 0) return;
        require_device(device);
        SelectArgs sb{};
        sb.mode = 2; sb.dense = d_dist; sb.cand_ids = d_ids;
        sb.ncols = (uint64_t)nlists * k; sb.inner = k; sb.row_stride = k; sb.outer_stride = (uint64_t)B * k;
        sb.B = B; sb.k = k; sb.out_ids = d_out_ids; sb.out_dist = d_out_dist; sb.out_count = d_out_count;
        launch_select(sb, (cudaStream_t)cuda_stream);
    });
}

int lgpu_flat_open(const float *vectors, uint64_t nrows, uint32_t dim, const uint64_t *row_ids, int device,
                   lgpu_flat **out)
{
    lgpu_flat *fl = nullptr;
    int rc = guarded([&] {
        LGPU_REQUIRE(out != nullptr && dim > 0, "null argument / zero dimension");
        LGPU_REQUIRE(nrows == 0 || vectors != nullptr, "null vectors");
        require_device(device);
        fl = new lgpu_flat();
        fl->device = device; fl->nrows = nrows; fl->dim = dim;
        fl->vectors.ensure(std::max<size_t>((size_t)nrows * dim * 4, 16));
        if (nrows) LGPU_CUDA(cudaMemcpy(fl->vectors.p, vectors, (size_t)nrows * dim * 4, cudaMemcpyHostToDevice));
        cudaDeviceProp prop;
        LGPU_CUDA(cudaGetDeviceProperties(&prop, device));
        fl->num_sms = prop.multiProcessorCount;
        if (gemm_shape_supported(dim) && nrows >= 4096) {
            prepare_tc_operand(fl->vectors.as<float>(), nrows, dim, fl->vec_b, fl->vec_n2, fl->vec_max, nullptr);
            fl->has_tc = true;
        }
        if (row_ids && nrows) {
            fl->row_ids.ensure((size_t)nrows * 8);
            LGPU_CUDA(cudaMemcpy(fl->row_ids.p, row_ids, (size_t)nrows * 8, cudaMemcpyHostToDevice));
            fl->has_ids = true;
        }
        register_handle(fl);
        *out = fl;
    });
    if (rc != LGPU_OK && fl) delete fl;
    return rc;
}

void lgpu_flat_close(lgpu_flat *fl)
{
    if (!retire_handle(fl)) return;
    cudaSetDevice(fl->device);
    cudaDeviceSynchronize();
    delete fl;
}

static void check_flat_call(lgpu_flat *fl, int metric, const void *q, uint32_t B, const lgpu_search_params *p,
                            const void *a, const void *b, const void *c)
{
    LGPU_REQUIRE(metric == LGPU_L2 || metric == LGPU_COSINE || metric == LGPU_DOT, "unknown distance type");
    check_params(p);
    LGPU_REQUIRE(B == 0 || (q && a && b && c), "null buffer");
}

int lgpu_flat_search(lgpu_flat *flh, int metric, const float *queries, uint32_t B, const lgpu_search_params *params,
                     uint64_t *out_ids, float *out_dist, uint32_t *out_count)
{
    return guarded([&] {
        HandleRef<lgpu_flat> fl(flh, "flat");
        check_flat_call(fl.h, metric, queries, B, params, out_ids, out_dist, out_count);
        if (B == 0) return;
        require_device(fl->device);
        uint64_t key[4];
        make_key(key, 0xf1a7ull + (uint64_t)metric, B, *params);
        host_call(fl->pool, queries, B, fl->dim, params->k, out_ids, out_dist, out_count, key, params->timeout_ms,
                  [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                      const Deadline &dl) { flat_search_device(fl.h, ws, st, metric, dq, B, *params, di, dd, dc, RowFilter(), &dl); });
    });
}

int lgpu_flat_search_filtered(lgpu_flat *flh, int metric, const float *queries, uint32_t B,
                              const lgpu_search_params *params, const uint32_t *allow, uint64_t allow_bits,
                              uint64_t *out_ids, float *out_dist, uint32_t *out_count)
{
    return guarded([&] {
        HandleRef<lgpu_flat> fl(flh, "flat");
        check_flat_call(fl.h, metric, queries, B, params, out_ids, out_dist, out_count);
        LGPU_REQUIRE(allow != nullptr || allow_bits == 0, "allow bitmap is null");
        if (B == 0) return;
        require_device(fl->device);
        uint64_t key[4];
        make_key(key, 0xf1b7ull + (uint64_t)metric, B, *params);
        host_call(fl->pool, queries, B, fl->dim, params->k, out_ids, out_dist, out_count, key, params->timeout_ms,
                  [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                      const Deadline &dl) {
                      const size_t words = (size_t)((allow_bits + 31) / 32);
                      ws->allow.ensure(std::max<size_t>(words, 1) * 4);
                      if (words) LGPU_CUDA(cudaMemcpyAsync(ws->allow.p, allow, words * 4, cudaMemcpyHostToDevice, st));
                      RowFilter rf; rf.bits = ws->allow.as<uint32_t>(); rf.nbits = allow_bits;
                      flat_search_device(fl.h, ws, st, metric, dq, B, *params, di, dd, dc, rf, &dl);
                  }, false);
    });
}

int lgpu_flat_search_device(lgpu_flat *flh, int metric, const float *d_queries, uint32_t B,
                            const lgpu_search_params *params, uint64_t *d_out_ids, float *d_out_dist,
                            uint32_t *d_out_count, void *cuda_stream)
{
    return guarded([&] {
        HandleRef<lgpu_flat> fl(flh, "flat");
        check_flat_call(fl.h, metric, d_queries, B, params, d_out_ids, d_out_dist, d_out_count);
        if (B == 0) return;
        require_device(fl->device);
        WsLease lease(fl->pool, (cudaStream_t)cuda_stream, true);
        flat_search_device(fl.h, lease.ws, lease.st, metric, d_queries, B, *params, d_out_ids, d_out_dist, d_out_count);
    });
}

// ---- partition-sharded multi-GPU search: one process per GPU, one ncclAllGather per batch ----
int lgpu_comm_unique_id(void *id_out, size_t id_bytes)
{
    return guarded([&] {
        LGPU_REQUIRE(id_out && id_bytes >= LGPU_COMM_ID_BYTES, "id buffer must hold LGPU_COMM_ID_BYTES bytes");
        static_assert(LGPU_COMM_ID_BYTES == sizeof(ncclUniqueId), "unique id size");
        ncclUniqueId id;
        LGPU_NCCL(nccl_api().GetUniqueId(&id));
        memcpy(id_out, &id, sizeof(id));
    });
}

int lgpu_comm_init(const void *unique_id, size_t id_bytes, int rank, int world, int device, lgpu_comm **out)
{
    lgpu_comm *c = nullptr;
    int rc = guarded([&] {
        LGPU_REQUIRE(unique_id && out && id_bytes >= LGPU_COMM_ID_BYTES, "null argument / short unique id");
        LGPU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank must be in [0, world)");
        require_device(device);
        NcclApi &api = nccl_api();
        c = new lgpu_comm();
        c->device = device; c->rank = rank; c->world = world;
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        LGPU_NCCL(api.CommInitRank(&c->comm, world, id, rank));
        for (auto &e : c->ev) LGPU_CUDA(cudaEventCreate(&e));
        register_handle(c);
        *out = c;
    });
    if (rc != LGPU_OK && c) { if (c->comm) nccl_api().CommDestroy(c->comm); delete c; }
    return rc;
}

void lgpu_comm_destroy(lgpu_comm *c)
{
    if (!retire_handle(c)) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    try { if (c->comm) nccl_api().CommDestroy(c->comm); } catch (const Failure &) {}
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    delete c;
}

// local top-k on this rank's shard -> pack -> ONE all-gather of [B][k] 16-byte records -> merge, all on `st`
static void sharded_search_device(lgpu_index *ix, lgpu_comm *c, Workspace *ws, cudaStream_t st, const float *d_q,
                                  uint32_t B, const lgpu_search_params &sp, uint64_t *d_ids, float *d_dist,
                                  uint32_t *d_cnt, const Deadline *deadline)
{
    LGPU_REQUIRE(c->device == ix->device, "communicator and index live on different devices");
    LGPU_REQUIRE(sp.refine_factor == 0, "refine_factor is not supported on the sharded path (raw vectors are not sharded)");
    std::lock_guard<std::mutex> g(c->mu);
    const size_t n = (size_t)B * sp.k;
    c->l_ids.ensure(n * 8); c->l_dist.ensure(n * 4); c->l_cnt.ensure((size_t)B * 4);
    c->send.ensure(n * sizeof(TopkRecord)); c->recv.ensure(n * sizeof(TopkRecord) * c->world);
    const bool prof = profiling_enabled();
    ivf_search_device(ix, ws, st, d_q, B, sp, c->l_ids.as<uint64_t>(), c->l_dist.as<float>(), c->l_cnt.as<uint32_t>(),
                      RowFilter(), deadline);
    launch_pack_records(c->l_ids.as<uint64_t>(), c->l_dist.as<float>(), n, c->send.as<TopkRecord>(), st);
    if (prof) cudaEventRecord(c->ev[0], st);
    LGPU_NCCL(nccl_api().AllGather(c->send.p, c->recv.p, n * sizeof(TopkRecord), ncclUint8, c->comm, st));
    if (prof) cudaEventRecord(c->ev[1], st);
    SelectArgs sb{};
    sb.mode = 2; sb.cand_rec = c->recv.as<TopkRecord>();
    sb.ncols = (uint64_t)c->world * sp.k; sb.inner = sp.k; sb.row_stride = sp.k; sb.outer_stride = (uint64_t)B * sp.k;
    sb.B = B; sb.k = sp.k; sb.out_ids = d_ids; sb.out_dist = d_dist; sb.out_count = d_cnt;
    launch_select(sb, st);
    if (prof) {
        cudaEventRecord(c->ev[2], st);
        LGPU_CUDA(cudaStreamSynchronize(st));
        c->last_ms[0] = g_stage_ms[6];
        cudaEventElapsedTime(&c->last_ms[1], c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&c->last_ms[2], c->ev[1], c->ev[2]);
    }
}

int lgpu_search_sharded(lgpu_index *ixh, lgpu_comm *ch, const float *queries, uint32_t B,
                        const lgpu_search_params *params, uint64_t *out_ids, float *out_dist, uint32_t *out_count)
{
    return guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        HandleRef<lgpu_comm> c(ch, "communicator");
        check_ivf_call(ix.h, queries, B, params, out_ids, out_dist, out_count);
        if (B == 0) return;
        require_device(ix->device);
        uint64_t key[4];
        make_key(key, 0x5a4dull, B, *params);
        host_call(ix->pool, queries, B, ix->dim, params->k, out_ids, out_dist, out_count, key, params->timeout_ms,
                  [&](Workspace *ws, cudaStream_t st, const float *dq, uint64_t *di, float *dd, uint32_t *dc,
                      const Deadline &dl) { sharded_search_device(ix.h, c.h, ws, st, dq, B, *params, di, dd, dc, &dl); },
                  false);
    });
}

int lgpu_search_sharded_device(lgpu_index *ixh, lgpu_comm *ch, const float *d_queries, uint32_t B,
                               const lgpu_search_params *params, uint64_t *d_out_ids, float *d_out_dist,
                               uint32_t *d_out_count, void *cuda_stream)
{
    return guarded([&] {
        HandleRef<lgpu_index> ix(ixh, "index");
        HandleRef<lgpu_comm> c(ch, "communicator");
        check_ivf_call(ix.h, d_queries, B, params, d_out_ids, d_out_dist, d_out_count);
        if (B == 0) return;
        require_device(ix->device);
        WsLease lease(ix->pool, (cudaStream_t)cuda_stream, true);
        sharded_search_device(ix.h, c.h, lease.ws, lease.st, d_queries, B, *params, d_out_ids, d_out_dist, d_out_count,
                              nullptr);
    });
}

int lgpu_comm_last_stage_ms(lgpu_comm *ch, float *times)
{
    return guarded([&] {
        LGPU_REQUIRE(times, "null argument");
        HandleRef<lgpu_comm> c(ch, "communicator");
        memcpy(times, c->last_ms, sizeof(c->last_ms));
    });
}

int lgpu_debug_coarse(lgpu_index *ixh, const float *queries, uint32_t B, uint32_t nprobes, uint32_t *out_parts,
                      float *out_dists)
{
    return guarded([&] {
        LGPU_REQUIRE(queries && out_parts && out_dists && B > 0 && nprobes > 0, "bad argument");
        HandleRef<lgpu_index> ix(ixh, "index");
        require_device(ix->device);
        nprobes = std::min(nprobes, ix->nlist);
        WsLease lease(ix->pool, nullptr, false);
        Workspace *ws = lease.ws; cudaStream_t st = lease.st;
        ws->q.ensure((size_t)B * ix->dim * 4);
        LGPU_CUDA(cudaMemcpyAsync(ws->q.p, queries, (size_t)B * ix->dim * 4, cudaMemcpyHostToDevice, st));
        const float *qs = ws->q.as<float>();
        if (ix->metric == LGPU_COSINE) {
            ws->qn.ensure((size_t)B * ix->dim * 4);
            launch_normalize(qs, B, ix->dim, ws->qn.as<float>(), st);
            qs = ws->qn.as<float>();
        }
        ws->D.ensure((size_t)B * ix->nlist * 4);
        ws->probes.ensure((size_t)B * nprobes * 8); ws->probe_dist.ensure((size_t)B * nprobes * 4);
        ws->probe_cnt.ensure((size_t)B * 4);
        launch_dist_matrix(qs, ix->centroids.as<float>(), B, ix->nlist, ix->dim, ix->metric == LGPU_DOT ? 1 : 0,
                           nullptr, nullptr, ws->D.as<float>(), ix->nlist, st);
        SelectArgs sa{};
        sa.mode = 1; sa.dense = ws->D.as<float>(); sa.ncols = ix->nlist; sa.row_stride = ix->nlist;
        sa.B = B; sa.k = nprobes; sa.out_ids = ws->probes.as<uint64_t>(); sa.out_dist = ws->probe_dist.as<float>();
        sa.out_count = ws->probe_cnt.as<uint32_t>();
        launch_select(sa, st);
        std::vector<uint64_t> tmp((size_t)B * nprobes);
        LGPU_CUDA(cudaMemcpyAsync(tmp.data(), ws->probes.p, tmp.size() * 8, cudaMemcpyDeviceToHost, st));
        LGPU_CUDA(cudaMemcpyAsync(out_dists, ws->probe_dist.p, tmp.size() * 4, cudaMemcpyDeviceToHost, st));
        LGPU_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < tmp.size(); i++) out_parts[i] = (uint32_t)tmp[i];
    });
}

int lgpu_debug_gemm(const float *Q, const float *X, uint32_t B, uint64_t N, uint32_t d, int device, float *out)
{
    return guarded([&] {
        LGPU_REQUIRE(Q && X && out && B > 0 && N > 0, "bad argument");
        LGPU_REQUIRE(gemm_shape_supported(d), "dimension must be a multiple of 8");
        require_device(device);
        cudaDeviceProp prop;
        LGPU_CUDA(cudaGetDeviceProperties(&prop, device));
        DevBuf q, x, qb, xb, xn2, o;
        const uint64_t ld = (N + 3) & ~3ull;
        q.ensure((size_t)B * d * 4); x.ensure((size_t)N * d * 4); qb.ensure((size_t)B * d * 2); xb.ensure((size_t)N * d * 2);
        xn2.ensure((size_t)N * 4); o.ensure((size_t)B * ld * 4);
        LGPU_CUDA(cudaMemcpy(q.p, Q, (size_t)B * d * 4, cudaMemcpyHostToDevice));
        LGPU_CUDA(cudaMemcpy(x.p, X, (size_t)N * d * 4, cudaMemcpyHostToDevice));
        launch_to_bf16(q.as<float>(), B, d, qb.p, nullptr, nullptr);
        launch_to_bf16(x.as<float>(), N, d, xb.p, xn2.as<float>(), nullptr);
        launch_gemm_dist(qb.p, xb.p, xn2.as<float>(), B, N, d, o.as<float>(), ld, prop.multiProcessorCount, nullptr);
        LGPU_CUDA(cudaDeviceSynchronize());
        LGPU_CUDA(cudaMemcpy2D(out, (size_t)N * 4, o.p, (size_t)ld * 4, (size_t)N * 4, B, cudaMemcpyDeviceToHost));
    });
}

int lgpu_debug_partition_distances(lgpu_index *ixh, const float *query, uint32_t part, float *out)
{
    return guarded([&] {
        LGPU_REQUIRE(query && out, "null argument");
        HandleRef<lgpu_index> ix(ixh, "index");
        LGPU_REQUIRE(part < ix->nlist, "partition out of range");
        require_device(ix->device);
        WsLease lease(ix->pool, nullptr, false);
        Workspace *ws = lease.ws; cudaStream_t st = lease.st;
        ws->q.ensure((size_t)ix->dim * 4);
        LGPU_CUDA(cudaMemcpyAsync(ws->q.p, query, (size_t)ix->dim * 4, cudaMemcpyHostToDevice, st));
        uint64_t forced = part;
        lgpu_search_params sp{};
        sp.k = 1; sp.nprobes = 1;
        ivf_sub_batch(ix.h, ws, st, ws->q.as<float>(), 1, sp, 1, nullptr, nullptr, nullptr, false, &forced);
        uint32_t n = ix->h_part_n[part];
        if (n) LGPU_CUDA(cudaMemcpyAsync(out, ws->dist_out.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        LGPU_CUDA(cudaStreamSynchronize(st));
    });
}

// ---- index build passes (build.cu): host buffers in, host buffers out, chunked over rows ----
int lgpu_ivf_assign(const float *centroids, uint32_t nlist, uint32_t dim, int metric, const float *vectors,
                    uint64_t n, int device, uint32_t *out_parts)
{
    return guarded([&] {
        LGPU_REQUIRE(centroids && nlist > 0 && dim > 0, "null centroids / empty shape");
        LGPU_REQUIRE(metric == LGPU_L2 || metric == LGPU_COSINE || metric == LGPU_DOT, "unknown distance type");
        LGPU_REQUIRE(n == 0 || (vectors && out_parts), "null vectors / output");
        if (n == 0) return;
        require_device(device);
        cudaStream_t st = nullptr;
        const uint64_t ld = (nlist + 3ull) & ~3ull;
        const uint64_t CH = std::max<uint64_t>(256, std::min<uint64_t>(65536, ((uint64_t)1 << 28) / (ld * 4)));
        DevBuf cent, x, xn, D, ids, dist, cnt;
        cent.ensure((size_t)nlist * dim * 4);
        LGPU_CUDA(cudaMemcpyAsync(cent.p, centroids, (size_t)nlist * dim * 4, cudaMemcpyHostToDevice, st));
        x.ensure((size_t)CH * dim * 4); D.ensure((size_t)CH * ld * 4);
        ids.ensure((size_t)CH * 8); dist.ensure((size_t)CH * 4); cnt.ensure((size_t)CH * 4);
        if (metric == LGPU_COSINE) xn.ensure((size_t)CH * dim * 4);
        std::vector<uint64_t> h_ids(CH);
        for (uint64_t r0 = 0; r0 < n; r0 += CH) {
            const uint32_t b = (uint32_t)std::min<uint64_t>(CH, n - r0);
            LGPU_CUDA(cudaMemcpyAsync(x.p, vectors + r0 * dim, (size_t)b * dim * 4, cudaMemcpyHostToDevice, st));
            const float *q = x.as<float>();
            if (metric == LGPU_COSINE) { launch_normalize(q, b, dim, xn.as<float>(), st); q = xn.as<float>(); }
            // the search path's own coarse step (find_partitions with nprobes = 1)
            launch_dist_matrix(q, cent.as<float>(), b, nlist, dim, metric == LGPU_DOT ? 1 : 0, nullptr, nullptr,
                               D.as<float>(), ld, st);
            SelectArgs sa{};
            sa.mode = 1; sa.dense = D.as<float>(); sa.ncols = nlist; sa.row_stride = ld; sa.B = b; sa.k = 1;
            sa.out_ids = ids.as<uint64_t>(); sa.out_dist = dist.as<float>(); sa.out_count = cnt.as<uint32_t>();
            launch_select(sa, st);
            LGPU_CUDA(cudaMemcpyAsync(h_ids.data(), ids.p, (size_t)b * 8, cudaMemcpyDeviceToHost, st));
            LGPU_CUDA(cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < b; i++) {
                LGPU_REQUIRE(h_ids[i] < nlist, "a vector has no finite centroid distance (NaN input?)");
                out_parts[r0 + i] = (uint32_t)h_ids[i];
            }
        }
    });
}

// nearest centre of every row (device buffers): the search's coarse step with nprobes = 1 -- tcgen05 GEMM scores +
// coarse_finish_kernel (exact re-score, lance arithmetic) where the shape allows it, the exact kernels otherwise
static void assign_nearest(const float *d_x, uint64_t n, uint32_t dim, const float *d_cent, uint32_t k, int num_sms,
                           uint64_t *d_ids, float *d_dist, cudaStream_t st)
{
    const uint64_t ld = (k + 3ull) & ~3ull;
    const uint64_t CH = std::max<uint64_t>(256, std::min<uint64_t>(65536, ((uint64_t)1 << 28) / (ld * 4)));
    DevBuf D, cnt, flags, gate, xb, xn2, cb, cn2;
    D.ensure((size_t)CH * ld * 4); cnt.ensure((size_t)CH * 4);
    const bool tc = gemm_shape_supported(dim) && k >= 256 && tc_enabled();
    float cmax = 0.f;
    if (tc) {
        flags.ensure((size_t)CH * 4); gate.ensure(16);
        xb.ensure((size_t)CH * dim * 2); xn2.ensure((size_t)CH * 4);
        prepare_tc_operand(d_cent, k, dim, cb, cn2, cmax, st);
    }
    for (uint64_t r0 = 0; r0 < n; r0 += CH) {
        const uint32_t b = (uint32_t)std::min<uint64_t>(CH, n - r0);
        const float *q = d_x + r0 * dim;
        SelectArgs sa{};
        sa.mode = 1; sa.dense = D.as<float>(); sa.ncols = k; sa.row_stride = ld; sa.B = b; sa.k = 1;
        sa.out_ids = d_ids + r0; sa.out_dist = d_dist + r0; sa.out_count = cnt.as<uint32_t>();
        if (tc) {
            launch_to_bf16(q, b, dim, xb.p, xn2.as<float>(), st);
            launch_gemm_dist(xb.p, cb.p, cn2.as<float>(), b, k, dim, D.as<float>(), ld, num_sms, st);
            launch_coarse_finish(D.as<float>(), ld, b, k, q, d_cent, xn2.as<float>(), cmax, dim, 1, d_ids + r0, d_dist + r0,
                                 cnt.as<uint32_t>(), flags.as<uint32_t>(), gate.as<uint32_t>(), st);
            launch_dist_matrix(q, d_cent, b, k, dim, 0, nullptr, nullptr, D.as<float>(), ld, st, flags.as<uint32_t>(),
                               gate.as<uint32_t>());
            sa.only = flags.as<uint32_t>(); sa.gate = gate.as<uint32_t>();
            launch_select(sa, st);
        } else {
            launch_dist_matrix(q, d_cent, b, k, dim, 0, nullptr, nullptr, D.as<float>(), ld, st);
            launch_select(sa, st);
        }
    }
    LGPU_CUDA(cudaStreamSynchronize(st));                            // the scratch buffers die with this scope
}

int lgpu_kmeans_train(const float *x, uint64_t n, uint32_t dim, float *centroids, uint32_t k, uint32_t iters, int device,
                      double *inertia_out)
{
    return guarded([&] {
        LGPU_REQUIRE(x && centroids && n > 0 && dim > 0 && k > 0, "null argument / empty shape");
        LGPU_REQUIRE(n < (1ull << 32), "too many training rows (sample them: sample_rate * num_partitions)");
        require_device(device);
        cudaDeviceProp prop;
        LGPU_CUDA(cudaGetDeviceProperties(&prop, device));
        cudaStream_t st = nullptr;
        DevBuf dx, dc, ids, dist, counts, offsets, cursor, rows, inert;
        dx.ensure((size_t)n * dim * 4); dc.ensure((size_t)k * dim * 4);
        ids.ensure((size_t)n * 8); dist.ensure((size_t)n * 4);
        counts.ensure((size_t)k * 4); offsets.ensure((size_t)(k + 1) * 4); cursor.ensure((size_t)k * 4);
        rows.ensure((size_t)n * 4); inert.ensure(16);
        LGPU_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st));
        LGPU_CUDA(cudaMemcpyAsync(dc.p, centroids, (size_t)k * dim * 4, cudaMemcpyHostToDevice, st));
        for (uint32_t it = 0; it < std::max<uint32_t>(iters, 1); it++) {
            assign_nearest(dx.as<float>(), n, dim, dc.as<float>(), k, prop.multiProcessorCount, ids.as<uint64_t>(),
                           dist.as<float>(), st);
            launch_kmeans_update(ids.as<uint64_t>(), dx.as<float>(), n, dim, k, counts.as<uint32_t>(), offsets.as<uint32_t>(),
                                 cursor.as<uint32_t>(), rows.as<uint32_t>(), dc.as<float>(), st);
        }
        if (inertia_out) {                                           // of the trained centres
            assign_nearest(dx.as<float>(), n, dim, dc.as<float>(), k, prop.multiProcessorCount, ids.as<uint64_t>(),
                           dist.as<float>(), st);
            launch_kmeans_inertia(dist.as<float>(), n, inert.as<double>(), st);
            LGPU_CUDA(cudaMemcpyAsync(inertia_out, inert.p, 8, cudaMemcpyDeviceToHost, st));
        }
        LGPU_CUDA(cudaMemcpyAsync(centroids, dc.p, (size_t)k * dim * 4, cudaMemcpyDeviceToHost, st));
        LGPU_CUDA(cudaStreamSynchronize(st));
    });
}

int lgpu_pq_train(const float *x, uint64_t n, uint32_t dim, uint32_t m, float *codebook, uint32_t iters, int device)
{
    return guarded([&] {
        LGPU_REQUIRE(x && codebook && n > 0 && dim > 0 && m > 0, "null argument / empty shape");
        LGPU_REQUIRE(dim % m == 0 && scan_dsub_supported(dim / m),
                     "unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        require_device(device);
        cudaStream_t st = nullptr;
        const uint32_t dsub = dim / m;
        DevBuf dx, cb, zero, parts, codes, sums, counts;
        dx.ensure((size_t)n * dim * 4); cb.ensure((size_t)m * 256 * dsub * 4); zero.ensure((size_t)dim * 4);
        parts.ensure((size_t)n * 4); codes.ensure((size_t)n * m); sums.ensure((size_t)m * 256 * dsub * 8);
        counts.ensure((size_t)m * 256 * 4);
        LGPU_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)n * dim * 4, cudaMemcpyHostToDevice, st));
        LGPU_CUDA(cudaMemcpyAsync(cb.p, codebook, (size_t)m * 256 * dsub * 4, cudaMemcpyHostToDevice, st));
        LGPU_CUDA(cudaMemsetAsync(zero.p, 0, (size_t)dim * 4, st));          // one all-zero "centroid": residual = row
        LGPU_CUDA(cudaMemsetAsync(parts.p, 0, (size_t)n * 4, st));
        for (uint32_t it = 0; it < std::max<uint32_t>(iters, 1); it++) {
            launch_pq_encode(dx.as<float>(), parts.as<uint32_t>(), zero.as<float>(), cb.as<float>(), n, dim, m, LGPU_L2,
                             codes.as<unsigned char>(), st);
            launch_pq_update(dx.as<float>(), codes.as<unsigned char>(), n, dim, m, sums.as<double>(), counts.as<uint32_t>(),
                             cb.as<float>(), st);
        }
        LGPU_CUDA(cudaMemcpyAsync(codebook, cb.p, (size_t)m * 256 * dsub * 4, cudaMemcpyDeviceToHost, st));
        LGPU_CUDA(cudaStreamSynchronize(st));
    });
}

int lgpu_pq_encode(const float *centroids, const float *codebook, uint32_t nlist, uint32_t dim, uint32_t m,
                   int metric, const float *vectors, const uint32_t *parts, uint64_t n, int device,
                   unsigned char *out_codes)
{
    return guarded([&] {
        LGPU_REQUIRE(centroids && codebook && nlist > 0 && dim > 0 && m > 0, "null index array / empty shape");
        LGPU_REQUIRE(dim % m == 0 && scan_dsub_supported(dim / m),
                     "unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        LGPU_REQUIRE(metric == LGPU_L2 || metric == LGPU_COSINE || metric == LGPU_DOT, "unknown distance type");
        LGPU_REQUIRE(n == 0 || (vectors && parts && out_codes), "null vectors / partitions / output");
        if (n == 0) return;
        for (uint64_t r = 0; r < n; r++) LGPU_REQUIRE(parts[r] < nlist, "partition id out of range");
        require_device(device);
        cudaStream_t st = nullptr;
        const uint64_t CH = 65536;
        DevBuf cent, cb, x, xn, p, codes;
        cent.ensure((size_t)nlist * dim * 4); cb.ensure((size_t)m * 256 * (dim / m) * 4);
        LGPU_CUDA(cudaMemcpyAsync(cent.p, centroids, (size_t)nlist * dim * 4, cudaMemcpyHostToDevice, st));
        LGPU_CUDA(cudaMemcpyAsync(cb.p, codebook, (size_t)m * 256 * (dim / m) * 4, cudaMemcpyHostToDevice, st));
        x.ensure((size_t)CH * dim * 4); p.ensure((size_t)CH * 4); codes.ensure((size_t)CH * m);
        if (metric == LGPU_COSINE) xn.ensure((size_t)CH * dim * 4);
        for (uint64_t r0 = 0; r0 < n; r0 += CH) {
            const uint32_t b = (uint32_t)std::min<uint64_t>(CH, n - r0);
            LGPU_CUDA(cudaMemcpyAsync(x.p, vectors + r0 * dim, (size_t)b * dim * 4, cudaMemcpyHostToDevice, st));
            LGPU_CUDA(cudaMemcpyAsync(p.p, parts + r0, (size_t)b * 4, cudaMemcpyHostToDevice, st));
            const float *q = x.as<float>();
            if (metric == LGPU_COSINE) { launch_normalize(q, b, dim, xn.as<float>(), st); q = xn.as<float>(); }
            launch_pq_encode(q, p.as<uint32_t>(), cent.as<float>(), cb.as<float>(), b, dim, m, metric,
                             codes.as<unsigned char>(), st);
            LGPU_CUDA(cudaMemcpyAsync(out_codes + r0 * m, codes.p, (size_t)b * m, cudaMemcpyDeviceToHost, st));
            LGPU_CUDA(cudaStreamSynchronize(st));
        }
    });
}

}  // extern "C"
