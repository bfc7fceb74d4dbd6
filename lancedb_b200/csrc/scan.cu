// scan.cu -- K2+K3: fused residual / PQ distance-table build / PQ code scan.
//
// Replaces, for a whole batch at once, what lance runs per (query, probed partition)
// inside ANNIvfSubIndexExec [lance, recalled; SURVEY.md 8a rows a4-a7]:
//     r   = q - centroid[p]                                  (residual, L2/cosine)
//     LUT = build_distance_table_l2(codebook, r)             (m x 256 f32)
//     d_j = sum_i LUT[i][code[i][j]]  sequentially in i      (compute_pq_distance)
// Results are bit-identical to oracle.c: every f32 op is an explicit round-to-nearest
// op in the reference's order (the LUT entry uses the f32x8 reduce tree, the row sum is
// sequential over sub-vectors).
//
// Work decomposition (B200-first, not the reference's per-query loop):
//   tile = (partition p, up to 8 of the queries that probe p, up to rows_tile of its rows).
//   A persistent grid (one 512-thread CTA per SM) pulls tiles from an atomic counter;
//   tiles are ordered by partition so a partition's codes are read from HBM once and
//   then hit in L2 for the other query groups.
//   The distance table is never materialised whole: it is built 8 sub-spaces at a time
//   into a ring of three 64 KB shared-memory buffers laid out [h][c][s][4 queries]
//   (h = query half, c = code, s = sub-space within the chunk), so one LDS.128 returns
//   the entries of 4 queries.  The codebook chunk is read (coalesced, L2-resident) once
//   per tile and amortised over the 8 queries.
//   Warp specialisation: the table build is FP32-pipe work (23 flops per entry, done
//   with packed FADD2/FFMA2), the scan is shared-memory-gather work; PW producer warps
//   build chunk ch+1/ch+2 while CW consumer warps scan chunk ch, handing buffers over
//   with named barriers (bar.arrive / bar.sync), so the two pipes overlap instead of
//   alternating behind a CTA-wide barrier.
//   Bank conflicts: a straightforward "lane = row" scan makes 8 lanes of a quarter-warp
//   gather at random codes => ~2.6-way conflicts.  Here lane l runs `l % 8` sub-space
//   slots behind lane 0 (the code stream in HBM is pre-skewed by row % 8 bytes, see
//   retile.cu), so at any instant the 8 lanes of a quarter-warp read 8 *different*
//   sub-spaces = 8 different 16-byte bank groups: conflict-free by construction, while
//   each row still accumulates its sub-vectors strictly in order 0..m-1 in its own
//   register.
//
// Algorithmic bytes per tile row and query: m code bytes.
#include "kernels.cuh"
#include "scan_common.cuh"

namespace lgpu {

namespace {

// Optional in-kernel stall accounting (LGPU_SCAN_TIMING=1): lane 0 of every warp accumulates clock64()
// deltas per region into a.timing[]: 0 producer total, 1 producer wait-EMPTY, 2 producer barrier,
// 3 consumer total, 4 consumer wait-FULL, 5 tile fetch (thread 0), 6 tiles, 7 producer warps, 8 consumer warps
// Compiled in only with -DLGPU_SCAN_TIMING_BUILD=1 (it costs registers); see profiles/r01_scan_stalls.txt.
#ifndef LGPU_SCAN_TIMING_BUILD
#define LGPU_SCAN_TIMING_BUILD 0
#endif
struct Tm {
    unsigned long long t0;
    __device__ __forceinline__ void start(const ScanArgs &a)
    {
        if constexpr (LGPU_SCAN_TIMING_BUILD) { if (a.timing) t0 = clock64(); }
    }
    __device__ __forceinline__ void stop(const ScanArgs &a, int slot)
    {
        if constexpr (LGPU_SCAN_TIMING_BUILD) {
            if (a.timing && (threadIdx.x & 31) == 0) atomicAdd(a.timing + slot, (unsigned long long)(clock64() - t0));
        }
    }
};

// scalar FADD/FMUL variant of the same four entries (selected with LGPU_SCALAR_TABLE=1 for A/B timing)
__device__ __forceinline__ float4 l2_tree8_scalar_x4(const uint64_t (&r)[4][4], const uint64_t (&c)[4])
{
    float cv[8], o[4];
#pragma unroll
    for (int e = 0; e < 4; e++) upk2(c[e], cv[2 * e], cv[2 * e + 1]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float rv[8], sq[8];
#pragma unroll
        for (int e = 0; e < 4; e++) upk2(r[j][e], rv[2 * e], rv[2 * e + 1]);
#pragma unroll
        for (int e = 0; e < 8; e++) { float d = __fsub_rn(rv[e], cv[e]); sq[e] = __fmul_rn(d, d); }
        o[j] = reduce_sum_x8(sq);
    }
    return make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------ producer side

// One (c, s, h) task = one 8-float codebook entry against the residuals of the 4 queries of
// half h: 4 table entries (one STS.128).  Lane -> s = lane&7, h = (lane>>3)&1, cc = lane>>4;
// warp pw takes the code pairs {2*(pw + PW*k) + cc}.  Codebook loads are double-buffered
// two tasks deep in registers so an L2 round trip is covered by the previous batch's math.
template <int DSUB, int PW, int NT>
__device__ __forceinline__ void produce_tile(const ScanArgs &a, uint32_t p, int ng, const uint32_t *s_q, int tid)
{
    extern __shared__ __align__(1024) unsigned char smem[];      // declared here so every access is a
    unsigned char *const lut = smem;                             // plain shared-space LDS/STS
    float *const rbuf = reinterpret_cast<float *>(smem + 3 * SCAN_LUT_BYTES);
    constexpr int PT = PW * 32;
    constexpr int RB = SCAN_G * 8 * DSUB;             // floats per residual chunk [g][s][e]
    const int lane = tid & 31, pw = tid >> 5;
    const int s = lane & 7, h = (lane >> 3) & 1, cc = lane >> 4;
    const bool active = 4 * h < ng;
    const uint32_t nch = a.nch;
    const float *cenp = a.centroids + (size_t)p * a.dim;

    // residual chunk ch -> registers (global loads issued now, consumed by store_resid later)
    constexpr int RPT = (RB + PT - 1) / PT;
    float rq[RPT], rc[RPT];
    auto load_resid = [&](uint32_t ch) {
#pragma unroll
        for (int u = 0; u < RPT; u++) {
            int idx = tid + u * PT;
            rq[u] = 0.f; rc[u] = 0.f;
            if (idx < RB) {
                int g = idx / (8 * DSUB), rem = idx - g * (8 * DSUB);
                int ss = rem / DSUB, e = rem - ss * DSUB;
                uint32_t i = ch * 8 + ss;
                if (g < ng && i < a.m) {
                    uint32_t dimi = i * DSUB + e;
                    rq[u] = __ldg(a.queries + (size_t)s_q[g] * a.dim + dimi);
                    if (a.metric != LGPU_DOT) rc[u] = __ldg(cenp + dimi);
                }
            }
        }
    };
    auto store_resid = [&](uint32_t ch) {
        float *dst = rbuf + (ch & 1) * RB;
#pragma unroll
        for (int u = 0; u < RPT; u++) {
            int idx = tid + u * PT;
            if (idx < RB) dst[idx] = __fsub_rn(rq[u], rc[u]);      // q - 0 == q for dot
        }
    };

    load_resid(0);
    store_resid(0);
    bar_sync(BAR_PROD, PT);

    for (uint32_t ch = 0; ch <= nch; ch++) {
        const int b = ch % 3;
        if (ch + 1 < nch) load_resid(ch + 1);
        if (ch >= 2) { Tm tm; tm.start(a); bar_sync(BAR_EMPTY + b, NT); tm.stop(a, 1); }   // consumers done with iteration ch-2
        if (ch == 0 && tid < 64)       // "chunk -1": lagging lanes read code 0 of buffer 2 in iteration 0
            reinterpret_cast<float *>(lut + 2 * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
        if (ch == nch) {               // "chunk nch": zero row for the lanes that ran out of sub-vectors
            if (tid < 64)
                reinterpret_cast<float *>(lut + b * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
        } else if (active) {
            const bool sub_ok = (ch * 8 + s) < a.m;
            const float *rsrc = rbuf + (ch & 1) * RB + ((4 * h) * 8 + s) * DSUB;   // + j*8*DSUB per query
            unsigned char *dst = lut + b * SCAN_LUT_BYTES + h * SCAN_LUT_HALF + s * 16;
            const float *cbp = a.cb_tiled + ((size_t)ch * 256 * 8 + s) * DSUB;
            constexpr int NPAIR = 128;                          // code pairs per chunk
            if constexpr (DSUB == 8) {
                uint64_t pr[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 lo = *reinterpret_cast<const float4 *>(rsrc + j * 8 * 8);
                    const float4 hi = *reinterpret_cast<const float4 *>(rsrc + j * 8 * 8 + 4);
                    pr[j][0] = pk2(lo.x, lo.y); pr[j][1] = pk2(lo.z, lo.w);
                    pr[j][2] = pk2(hi.x, hi.y); pr[j][3] = pk2(hi.z, hi.w);
                }
                const bool l2 = a.metric != LGPU_DOT;
                // software pipeline over this warp's code pairs, 2 tasks per stage
                float4 bufA[2][2], bufB[2][2];
                auto fetch = [&](float4 (&buf)[2][2], int k) {
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        int pair = pw + PW * (k + u);
                        if (pair < NPAIR) {
                            const float4 *src = reinterpret_cast<const float4 *>(cbp + (size_t)(2 * pair + cc) * 8 * 8);
                            buf[u][0] = __ldg(src); buf[u][1] = __ldg(src + 1);
                        }
                    }
                };
                auto compute = [&](const float4 (&cur)[2][2], int k) {
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        int pair = pw + PW * (k + u);
                        if (pair < NPAIR) {
                            const int c = 2 * pair + cc;
                            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (sub_ok) {
                                if (l2) {
                                    uint64_t pc[4] = {pk2(cur[u][0].x, cur[u][0].y), pk2(cur[u][0].z, cur[u][0].w),
                                                      pk2(cur[u][1].x, cur[u][1].y), pk2(cur[u][1].z, cur[u][1].w)};
                                    out = a.scalar_table ? l2_tree8_scalar_x4(pr, pc) : l2_tree8_packed_x4(pr, pc, a.fzero2);
                                } else {
                                    float cv[8] = {cur[u][0].x, cur[u][0].y, cur[u][0].z, cur[u][0].w,
                                                   cur[u][1].x, cur[u][1].y, cur[u][1].z, cur[u][1].w};
                                    float rr[8];
                                    float o[4];
#pragma unroll
                                    for (int j = 0; j < 4; j++) {
#pragma unroll
                                        for (int e = 0; e < 4; e++) upk2(pr[j][e], rr[2 * e], rr[2 * e + 1]);
                                        o[j] = subvec_dot_dist<8>(rr, cv);
                                    }
                                    out = make_float4(o[0], o[1], o[2], o[3]);
                                }
                            }
                            *reinterpret_cast<float4 *>(dst + c * 128) = out;
                        }
                    }
                };
                // ping-pong: the loads of stage k+2 are in flight while stage k is computed
                fetch(bufA, 0);
                for (int k = 0;; k += 4) {
                    fetch(bufB, k + 2);
                    compute(bufA, k);
                    if (pw + PW * (k + 2) >= NPAIR) break;
                    fetch(bufA, k + 4);
                    compute(bufB, k + 2);
                    if (pw + PW * (k + 4) >= NPAIR) break;
                }
            } else {
                for (int k = 0; pw + PW * k < NPAIR; k++) {
                    const int c = 2 * (pw + PW * k) + cc;
                    float cbv[DSUB], rr[DSUB], o[4] = {0.f, 0.f, 0.f, 0.f};
                    load_vec<DSUB>(cbv, cbp + (size_t)c * 8 * DSUB);
                    if (sub_ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
#pragma unroll
                            for (int e = 0; e < DSUB; e++) rr[e] = rsrc[j * 8 * DSUB + e];
                            o[j] = (a.metric == LGPU_DOT) ? subvec_dot_dist<DSUB>(rr, cbv) : subvec_l2<DSUB>(rr, cbv);
                        }
                    }
                    *reinterpret_cast<float4 *>(dst + c * 128) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        bar_arrive(BAR_FULL + b, NT);
        if (ch + 1 < nch) store_resid(ch + 1);
        if (ch < nch) { Tm tm; tm.start(a); bar_sync(BAR_PROD, PT); tm.stop(a, 2); }   // residual chunk ch+1 visible
    }
}


// ------------------------------------------------------------------ approximate pass
// Producer of the approximate pass: nothing is computed per (query, partition).  The per-QUERY
// tables T_q[ch][c][s] = |q_i - codebook_i[c]|^2 (tables.cu) are copied into the same
// [h][c][s][4 queries] ring, 4 scalar loads + one STS.128 per task.
template <int PW, int NT>
__device__ __forceinline__ void produce_tile_copy(const ScanArgs &a, int ng, const uint32_t *s_q, int tid)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *const lut = smem;
    const int lane = tid & 31, pw = tid >> 5;
    const int s = lane & 7, h = (lane >> 3) & 1, cc = lane >> 4;
    const bool active = 4 * h < ng;
    const uint32_t nch = a.nch;
    const size_t tq_stride = (size_t)nch * 256 * 8;          // floats per query table
    const float *tp[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int g = 4 * h + j;
        tp[j] = a.tq + (size_t)s_q[g < ng ? g : 0] * tq_stride + s;     // idle slots alias query 0 (ignored later)
    }
    constexpr int NPAIR = 128, UNR = 8;
    for (uint32_t ch = 0; ch <= nch; ch++) {
        const int b = ch % 3;
        if (ch >= 2) bar_sync(BAR_EMPTY + b, NT);
        if (ch == 0 && tid < 64)
            reinterpret_cast<float *>(lut + 2 * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
        if (ch == nch) {
            if (tid < 64)
                reinterpret_cast<float *>(lut + b * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
        } else if (active) {
            unsigned char *dst = lut + b * SCAN_LUT_BYTES + h * SCAN_LUT_HALF + s * 16;
            const size_t choff = (size_t)ch * 256 * 8;
            for (int k0 = 0; pw + PW * k0 < NPAIR; k0 += UNR) {
                float4 v[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int pair = pw + PW * (k0 + u);
                    if (pair < NPAIR) {
                        const size_t o = choff + (size_t)(2 * pair + cc) * 8;
                        v[u] = make_float4(__ldg(tp[0] + o), __ldg(tp[1] + o), __ldg(tp[2] + o), __ldg(tp[3] + o));
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int pair = pw + PW * (k0 + u);
                    if (pair < NPAIR) *reinterpret_cast<float4 *>(dst + (2 * pair + cc) * 128) = v[u];
                }
            }
        }
        bar_arrive(BAR_FULL + b, NT);
    }
}

// ------------------------------------------------------------------ consumer side
template <int R, int CT, int NT>
__device__ __forceinline__ void consume_tile(const ScanArgs &a, uint32_t p, int ng, uint32_t row0, uint32_t nrows,
                                             const uint64_t *s_out, const float *s_A, int ct)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned char *const lut = smem;
    const int sig = ct & 7;                       // this lane's skew (== row % 8)
    const uint32_t nch = a.nch;
    const uint32_t n_p = a.part_n[p], npad = a.part_npad[p];
    const uint2 *cs = reinterpret_cast<const uint2 *>(a.codes + a.code_base[p]);   // [nch+1][npad]
    const bool two_halves = ng > 4;

    float acc[R][SCAN_G];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int g = 0; g < SCAN_G; g++) acc[r][g] = 0.f;

    bool valid[R];
    uint2 wn[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t row = row0 + ct + r * CT;
        valid[r] = row < row0 + nrows && row < n_p;
        wn[r] = valid[r] ? __ldg(cs + row) : make_uint2(0u, 0u);
    }

    for (uint32_t it = 0; it <= nch; it++) {
        uint2 w[R];
#pragma unroll
        for (int r = 0; r < R; r++) w[r] = wn[r];
        if (it < nch) {                            // prefetch the next block of code bytes
#pragma unroll
            for (int r = 0; r < R; r++) {
                uint32_t row = row0 + ct + r * CT;
                wn[r] = valid[r] ? __ldg(cs + (size_t)(it + 1) * npad + row) : make_uint2(0u, 0u);
            }
        }
        { Tm tm; tm.start(a); bar_sync(BAR_FULL + (int)(it % 3), NT); tm.stop(a, 4); }
        const uint32_t base_cur = (it % 3) * SCAN_LUT_BYTES;
        const uint32_t base_prev = ((it + 2) % 3) * SCAN_LUT_BYTES;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const uint32_t off = ((e < sig) ? base_prev : base_cur) + (((e - sig) & 7) << 4);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t word = (e < 4) ? w[r].x : w[r].y;
                const uint32_t c = (word >> (8 * (e & 3))) & 0xffu;
                const unsigned char *ent = lut + off + (c << 7);
                const float4 v0 = *reinterpret_cast<const float4 *>(ent);
                acc[r][0] = __fadd_rn(acc[r][0], v0.x);
                acc[r][1] = __fadd_rn(acc[r][1], v0.y);
                acc[r][2] = __fadd_rn(acc[r][2], v0.z);
                acc[r][3] = __fadd_rn(acc[r][3], v0.w);
                if (two_halves) {
                    const float4 v1 = *reinterpret_cast<const float4 *>(ent + SCAN_LUT_HALF);
                    acc[r][4] = __fadd_rn(acc[r][4], v1.x);
                    acc[r][5] = __fadd_rn(acc[r][5], v1.y);
                    acc[r][6] = __fadd_rn(acc[r][6], v1.z);
                    acc[r][7] = __fadd_rn(acc[r][7], v1.w);
                }
            }
        }
        // buffer (it+2)%3 == (it-1)%3 (chunk it-1, or the zero row before chunk 0) is no longer
        // read by anyone; the producers wait for it iff they still have chunk it+2 (or the
        // final zero row) to put there
        if (it + 2 <= nch) bar_arrive(BAR_EMPTY + (int)((it + 2) % 3), NT);
    }

    // ---- epilogue: metric post-processing, one f32 per (row, query) to HBM ----
    const float mcorr = (float)(a.m - 1);
#pragma unroll
    for (int g = 0; g < SCAN_G; g++) {
        if (g < ng) {
            float *out = a.dist_out + s_out[g];
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (valid[r]) {
                    float v = acc[r][g];
                    if (a.tq) v = (v + s_A[g]) + a.row_R[a.part_off[p] + row0 + ct + r * CT];   // approximate pass
                    if (a.metric == LGPU_COSINE) v = __fmul_rn(v, 0.5f);
                    else if (a.metric == LGPU_DOT) v = __fsub_rn(v, mcorr);
                    out[row0 + ct + r * CT] = v;
                }
            }
        }
    }
}

// Per-tile bookkeeping shared by both roles (every thread of the CTA runs it, so the two
// __syncthreads line up across the role-specific loops).
struct TileInfo {
    uint32_t p, row0, nrows;
    int ng;
    bool done;
};

template <int NT>
__device__ __forceinline__ TileInfo next_tile(const ScanArgs &a, uint32_t total, uint32_t *s_tile, uint32_t *s_p,
                                              uint32_t *s_q, uint64_t *s_out, float *s_A, int tid)
{
    TileInfo ti;
    __syncthreads();                            // previous tile fully drained
    if (tid == 0) {
        uint32_t t = atomicAdd(a.tile_counter, 1u);
        *s_tile = t;
        if (t < total) {                        // smallest p with tile_off[p+1] > t
            uint32_t lo = 0, hi = a.nlist - 1;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (a.tile_off[mid + 1] > t) hi = mid; else lo = mid + 1;
            }
            *s_p = lo;
        }
    }
    __syncthreads();
    const uint32_t t = *s_tile;
    ti.done = t >= total;
    if (ti.done) { ti.p = 0; ti.row0 = 0; ti.nrows = 0; ti.ng = 0; return ti; }
    const uint32_t p = *s_p;
    const uint32_t n_p = a.part_n[p];
    const uint32_t nrb = scan_nrb(n_p, a.rows_tile), rbr = scan_rb_rows(n_p, nrb);
    const uint32_t local = t - a.tile_off[p];
    const uint32_t grp = local / nrb, rb = local - grp * nrb;
    ti.p = p;
    ti.ng = (int)min((uint32_t)SCAN_G, a.part_cnt[p] - grp * SCAN_G);
    ti.row0 = rb * rbr;
    ti.nrows = ti.row0 < n_p ? min(rbr, n_p - ti.row0) : 0;
    if (tid < SCAN_G) {
        if (tid < ti.ng) {
            uint32_t e = a.qlist[a.qlist_off[p] + grp * SCAN_G + tid];
            s_q[tid] = e / a.nprobes;
            s_out[tid] = a.seg_off[e];
            if (a.tq) s_A[tid] = a.probe_A[e];
        } else {
            s_q[tid] = 0xffffffffu;
        }
    }
    __syncthreads();
    return ti;
}

// PW producer warps + CW consumer warps (both multiples of 4: setmaxnreg works on
// warpgroups).  The CTA is launched with 128 registers per thread; producers give
// registers back (PREG) and consumers take them (CREG) so that a consumer thread can hold
// RMAX rows x 8 queries of accumulators plus two blocks of code bytes.
template <int DSUB, int PW, int CW, int RMAX, int PREG, int CREG>
__global__ void __launch_bounds__((PW + CW) * 32, 1) scan_kernel(ScanArgs a)
{
    constexpr int NT = (PW + CW) * 32, CT = CW * 32;
    static_assert(PW % 4 == 0 && CW % 4 == 0, "roles must be whole warpgroups");
    static_assert(PW * 32 * PREG + CW * 32 * CREG <= 65536, "register budget");
    // dynamic smem: 3 x 64 KB table ring, then 2 x [8 g][8 s][DSUB] residual chunks
    __shared__ uint32_t s_tile, s_p;
    __shared__ uint32_t s_q[SCAN_G];
    __shared__ uint64_t s_out[SCAN_G];
    __shared__ float s_A[SCAN_G];

    const int tid = threadIdx.x;
    const uint32_t total = *a.total_tiles;

    if (tid < PW * 32) {
        if constexpr (PREG != CREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PREG));
        for (;;) {
            Tm tf; tf.start(a);
            TileInfo ti = next_tile<NT>(a, total, &s_tile, &s_p, s_q, s_out, s_A, tid);
            if (tid == 0) { tf.stop(a, 5); if constexpr (LGPU_SCAN_TIMING_BUILD) { if (a.timing && !ti.done) atomicAdd(a.timing + 6, 1ull); } }
            if (ti.done) break;
            Tm tt; tt.start(a);
            if constexpr (DSUB == 0) produce_tile_copy<PW, NT>(a, ti.ng, s_q, tid);
            else produce_tile<DSUB, PW, NT>(a, ti.p, ti.ng, s_q, tid);
            tt.stop(a, 0);
        }
    } else {
        if constexpr (PREG != CREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CREG));
        const int ct = tid - PW * 32;
        for (;;) {
            TileInfo ti = next_tile<NT>(a, total, &s_tile, &s_p, s_q, s_out, s_A, tid);
            if (ti.done) break;
            const int R = (int)((ti.nrows + CT - 1) / CT);     // uniform per tile; rounded up to even
            Tm tt; tt.start(a);
#define LGPU_CONSUME(RR) consume_tile<RR, CT, NT>(a, ti.p, ti.ng, ti.row0, ti.nrows, s_out, s_A, ct)
            if (R <= 2) LGPU_CONSUME(2);
            else if (R <= 4) LGPU_CONSUME(4);
            else if (R <= 6 || RMAX <= 6) LGPU_CONSUME(6);
            else if constexpr (RMAX >= 8) {
                if (R <= 8 || RMAX <= 8) LGPU_CONSUME(8);
                else if constexpr (RMAX > 8) { if (R <= 10) LGPU_CONSUME(10); else LGPU_CONSUME(12); }
            }
#undef LGPU_CONSUME
            tt.stop(a, 3);
        }
    }
}

template <int DSUB, int PW, int CW, int RMAX, int PREG, int CREG>
void launch_one(const ScanArgs &a, int grid, cudaStream_t st)
{
    size_t smem = 3 * (size_t)SCAN_LUT_BYTES + 2 * (size_t)SCAN_G * 8 * DSUB * sizeof(float);
    auto kern = scan_kernel<DSUB, PW, CW, RMAX, PREG, CREG>;
    LGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.rows_tile != (uint32_t)(CW * 32 * RMAX)) {
        set_error("internal: rows_tile does not match the scan kernel variant");
        throw Failure{LGPU_RUNTIME};
    }
    kern<<<grid, (PW + CW) * 32, smem, st>>>(a); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

template <>
void launch_one<0, 8, 8, 6, 128, 128>(const ScanArgs &a, int grid, cudaStream_t st)
{
    size_t smem = 3 * (size_t)SCAN_LUT_BYTES;
    auto kern = scan_kernel<0, 8, 8, 6, 128, 128>;
    LGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.rows_tile > 8u * 32u * 6u) {
        set_error("internal: rows_tile too large for the approximate scan variant");
        throw Failure{LGPU_RUNTIME};
    }
    kern<<<grid, 512, smem, st>>>(a); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

template <int DSUB>
void launch_variant(const ScanArgs &a, int grid, cudaStream_t st)
{
    if (a.rows_tile == SCAN_ROWS_TILE_MID) launch_one<DSUB, 12, 4, 12, 104, 200>(a, grid, st);
    else launch_one<DSUB, 8, 8, 8, 96, 160>(a, grid, st);
}

}  // namespace

bool scan_dsub_supported(uint32_t dsub)
{
    return dsub == 1 || dsub == 2 || dsub == 4 || dsub == 8 || dsub == 16 || dsub == 32;
}

void launch_scan(const ScanArgs &a, uint32_t dsub, int grid, cudaStream_t st)
{
    if (a.tq) { launch_one<0, 8, 8, 6, 128, 128>(a, grid, st); return; }   // approximate pass (any dsub)
    switch (dsub) {
    case 1: launch_variant<1>(a, grid, st); break;
    case 2: launch_variant<2>(a, grid, st); break;
    case 4: launch_variant<4>(a, grid, st); break;
    case 8: launch_variant<8>(a, grid, st); break;
    case 16: launch_variant<16>(a, grid, st); break;
    case 32: launch_variant<32>(a, grid, st); break;
    default:
        set_error("unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        throw Failure{LGPU_INVALID_INPUT};
    }
}

}  // namespace lgpu
