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
//   tile = (partition p, up to 8 of the queries that probe p, up to 2048 of its rows).
//   A persistent grid (one 512-thread CTA per SM) pulls tiles from an atomic counter;
//   tiles are ordered by partition so a partition's codes are read from HBM once and
//   then hit in L2 for the other query groups.
//   The distance table is never materialised whole: it is built 8 sub-spaces at a time
//   into a ring of three 64 KB shared-memory buffers laid out [h][c][s][4 queries]
//   (h = query half, c = code, s = sub-space within the chunk), so one LDS.128 returns
//   the entries of 4 queries.  The codebook chunk is read (coalesced, L2-resident) once
//   per tile and amortised over the 8 queries.
//   Bank conflicts: a straightforward "lane = row" scan makes 8 lanes of a quarter-warp
//   gather at random codes => ~2.6-way conflicts.  Here lane l runs `l % 8` sub-space
//   slots behind lane 0 (the code stream in HBM is pre-skewed by row % 8 bytes, see
//   index.cu), so at any instant the 8 lanes of a quarter-warp read 8 *different*
//   sub-spaces = 8 different 16-byte bank groups: conflict-free by construction, while
//   each row still accumulates its sub-vectors strictly in order 0..m-1 in its own
//   register.
//
// Algorithmic bytes per tile row and query: m code bytes.
#include "kernels.cuh"

namespace lgpu {

namespace {

template <int DSUB>
__device__ __forceinline__ void load_vec(float *dst, const float *src)
{
    if constexpr (DSUB % 4 == 0) {
#pragma unroll
        for (int i = 0; i < DSUB / 4; i++) {
            float4 v = reinterpret_cast<const float4 *>(src)[i];
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < DSUB; i++) dst[i] = src[i];
    }
}

template <int DSUB>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(ScanArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *lut = smem;                                              // 3 x 64 KB ring
    float *rbuf = reinterpret_cast<float *>(smem + 3 * SCAN_LUT_BYTES);      // 2 x [8 g][8 s][DSUB]
    __shared__ uint32_t s_tile, s_p;
    __shared__ uint32_t s_q[SCAN_G];
    __shared__ uint64_t s_out[SCAN_G];

    const int tid = threadIdx.x;
    const int sig = tid & 7;                       // this lane's skew (== row % 8)
    const uint32_t total = *a.total_tiles;
    const uint32_t nch = a.nch;
    constexpr int RB_FLOATS = SCAN_G * 8 * DSUB;

    for (;;) {
        __syncthreads();                            // previous tile fully drained
        if (tid == 0) {
            uint32_t t = atomicAdd(a.tile_counter, 1u);
            s_tile = t;
            if (t < total) {                        // smallest p with tile_off[p+1] > t
                uint32_t lo = 0, hi = a.nlist - 1;
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (a.tile_off[mid + 1] > t) hi = mid; else lo = mid + 1;
                }
                s_p = lo;
            }
        }
        __syncthreads();
        const uint32_t t = s_tile;
        if (t >= total) break;
        const uint32_t p = s_p;
        const uint32_t n_p = a.part_n[p], npad = a.part_npad[p];
        const uint32_t nrb = scan_nrb(n_p), rbr = scan_rb_rows(n_p, nrb);
        const uint32_t local = t - a.tile_off[p];
        const uint32_t grp = local / nrb, rb = local - grp * nrb;
        const int ng = (int)min((uint32_t)SCAN_G, a.part_cnt[p] - grp * SCAN_G);
        const uint32_t row0 = rb * rbr;
        const uint32_t nrows = row0 < n_p ? min(rbr, n_p - row0) : 0;
        const int R = (int)((nrows + SCAN_THREADS - 1) / SCAN_THREADS);

        if (tid < SCAN_G) {
            if (tid < ng) {
                uint32_t e = a.qlist[a.qlist_off[p] + grp * SCAN_G + tid];
                s_q[tid] = e / a.nprobes;
                s_out[tid] = a.seg_off[e];
            } else {
                s_q[tid] = 0xffffffffu;
            }
        }
        // "chunk -1" is read by lagging lanes during iteration 0 at code 0 only: zero it
        if (tid < 64)
            reinterpret_cast<float *>(lut + 2 * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
        __syncthreads();

        // ---- residual chunk: rbuf[g][s][e] = q_g[i*DSUB+e] - centroid_p[i*DSUB+e] ----
        auto compute_rbuf = [&](uint32_t ch) {
            float *dst = rbuf + (ch & 1) * RB_FLOATS;
            for (int idx = tid; idx < RB_FLOATS; idx += SCAN_THREADS) {
                int g = idx / (8 * DSUB), rem = idx - g * (8 * DSUB);
                int s = rem / DSUB, e = rem - s * DSUB;
                uint32_t i = ch * 8 + s;
                float val = 0.f;
                if (g < ng && i < a.m) {
                    uint32_t dimi = i * DSUB + e;
                    float qv = a.queries[(size_t)s_q[g] * a.dim + dimi];
                    val = (a.metric == LGPU_DOT) ? qv
                                                 : __fsub_rn(qv, a.centroids[(size_t)p * a.dim + dimi]);
                }
                dst[idx] = val;
            }
        };
        // ---- distance-table chunk: lut[ch%3][h][c][s][g&3] ----
        auto build = [&](uint32_t ch) {
            const int s = tid & 7, gp = (tid >> 3) & 3, cg = tid >> 5;
            const int g0 = 2 * gp;
            if (g0 >= ng) return;
            const bool sub_ok = (ch * 8 + s) < a.m;
            const float *rb0 = rbuf + (ch & 1) * RB_FLOATS + (g0 * 8 + s) * DSUB;
            float r0[DSUB], r1[DSUB];
            load_vec<DSUB>(r0, rb0);
            load_vec<DSUB>(r1, rb0 + 8 * DSUB);
            unsigned char *dst = lut + (ch % 3) * SCAN_LUT_BYTES + (gp >> 1) * SCAN_LUT_HALF + s * 16 +
                                 (gp & 1) * 8 + (cg * 16) * 128;
            const float *cbp = a.cb_tiled + ((size_t)(ch * 256 + cg * 16) * 8 + s) * DSUB;
            constexpr int UNR = (DSUB <= 8) ? 4 : (DSUB <= 16 ? 2 : 1);
            for (int jj = 0; jj < 16; jj += UNR) {
                float cbv[UNR][DSUB];
#pragma unroll
                for (int u = 0; u < UNR; u++) load_vec<DSUB>(cbv[u], cbp + (size_t)(jj + u) * 8 * DSUB);
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    float e0 = 0.f, e1 = 0.f;
                    if (sub_ok) {
                        if (a.metric == LGPU_DOT) {
                            e0 = subvec_dot_dist<DSUB>(r0, cbv[u]);
                            e1 = subvec_dot_dist<DSUB>(r1, cbv[u]);
                        } else {
                            e0 = subvec_l2<DSUB>(r0, cbv[u]);
                            e1 = subvec_l2<DSUB>(r1, cbv[u]);
                        }
                    }
                    *reinterpret_cast<float2 *>(dst + (jj + u) * 128) = make_float2(e0, e1);
                }
            }
        };

        compute_rbuf(0);
        if (nch > 1) compute_rbuf(1);
        __syncthreads();
        build(0);
        __syncthreads();

        float acc[SCAN_RMAX][SCAN_G];
#pragma unroll
        for (int r = 0; r < SCAN_RMAX; r++)
#pragma unroll
            for (int g = 0; g < SCAN_G; g++) acc[r][g] = 0.f;

        const uint2 *cs = reinterpret_cast<const uint2 *>(a.codes + a.code_base[p]);   // [nch+1][npad]
        const bool two_halves = ng > 4;

        for (uint32_t it = 0; it <= nch; it++) {
            if (it + 1 < nch) build(it + 1);
            else if (it + 1 == nch && tid < 64)     // "chunk nch": zero row for the drained lanes
                reinterpret_cast<float *>(lut + (nch % 3) * SCAN_LUT_BYTES + (tid >> 5) * SCAN_LUT_HALF)[tid & 31] = 0.f;
            if (it + 2 < nch) compute_rbuf(it + 2);

            const uint32_t base_cur = (it % 3) * SCAN_LUT_BYTES;
            const uint32_t base_prev = ((it + 2) % 3) * SCAN_LUT_BYTES;
            uint2 w[SCAN_RMAX];
#pragma unroll
            for (int r = 0; r < SCAN_RMAX; r++) {
                uint32_t row = row0 + tid + r * SCAN_THREADS;
                w[r] = (r < R && row < n_p) ? __ldg(cs + (size_t)it * npad + row) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t off = ((e < sig) ? base_prev : base_cur) + (((e - sig) & 7) << 4);
#pragma unroll
                for (int r = 0; r < SCAN_RMAX; r++) {
                    if (r < R) {
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
            }
            __syncthreads();
        }

        // ---- epilogue: metric post-processing, one f32 per (row, query) to HBM ----
        const float mcorr = (float)(a.m - 1);
#pragma unroll
        for (int g = 0; g < SCAN_G; g++) {
            if (g < ng) {
                float *out = a.dist_out + s_out[g];
#pragma unroll
                for (int r = 0; r < SCAN_RMAX; r++) {
                    uint32_t row = row0 + tid + r * SCAN_THREADS;
                    if (r < R && row < row0 + nrows) {
                        float v = acc[r][g];
                        if (a.metric == LGPU_COSINE) v = __fmul_rn(v, 0.5f);
                        else if (a.metric == LGPU_DOT) v = __fsub_rn(v, mcorr);
                        out[row] = v;
                    }
                }
            }
        }
    }
}

template <int DSUB>
void launch_one(const ScanArgs &a, int grid, cudaStream_t st)
{
    size_t smem = 3 * (size_t)SCAN_LUT_BYTES + 2 * (size_t)SCAN_G * 8 * DSUB * sizeof(float);
    static bool configured = false;   // per template instance
    if (!configured) {
        LGPU_CUDA(cudaFuncSetAttribute(scan_kernel<DSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    scan_kernel<DSUB><<<grid, SCAN_THREADS, smem, st>>>(a);
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace

bool scan_dsub_supported(uint32_t dsub)
{
    return dsub == 1 || dsub == 2 || dsub == 4 || dsub == 8 || dsub == 16 || dsub == 32;
}

void launch_scan(const ScanArgs &a, uint32_t dsub, int grid, cudaStream_t st)
{
    switch (dsub) {
    case 1: launch_one<1>(a, grid, st); break;
    case 2: launch_one<2>(a, grid, st); break;
    case 4: launch_one<4>(a, grid, st); break;
    case 8: launch_one<8>(a, grid, st); break;
    case 16: launch_one<16>(a, grid, st); break;
    case 32: launch_one<32>(a, grid, st); break;
    default:
        set_error("unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        throw Failure{LGPU_INVALID_INPUT};
    }
}

}  // namespace lgpu
