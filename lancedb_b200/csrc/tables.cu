// tables.cu -- the filter + verify form of the PQ scan: per-query tables, per-row constants, proof and re-score.
//
// The exact path (scan2.cu) builds one 96 KB distance table per (query, probed partition), because lance's table
// is on the *residual* q - c_p: 20 480 tables per 1024-query batch, 23 f32 ops per
// entry, and that table build -- not the scan -- bounds the kernel.  Algebraically
//     |(q_i - c_i) - b_i|^2 = |q_i - b_i|^2  +  (|c_i|^2 - 2 q_i.c_i)  +  2 b_i.c_i
// so the distance of row r of partition p is  S_q(r) + A(q,p) + R(r)  with
//     S_q(r) = sum_i T_q[i][code_i(r)],  T_q[i][c] = |q_i - codebook_i[c]|^2    one table per QUERY
//     A(q,p) = |q - c_p|^2 - |q|^2                                               from the coarse step
//     R(r)   = 2 * sum_i codebook_i[code_i(r)] . c_p,i                           one f32 per row, at open
// (dot: T_q[i][c] = 1 - q_i.codebook_i[c], A = R = 0, distance = S - (m - 1): lance's own table, no residual).
// The rounding differs from lance's, so this is only used as a FILTER.  T_q is quantised to 16 bits against a
// per-query step (launch_query_tables_q16): the scan kernel (scan3.cu) adds the m integers of a row exactly and
// turns the sum into a LOWER bound L(r) of the row's distance; with W = m step the real distance lies in
// [L - E, L + W + E] (E = floating-point slack, launch_band_check3).  The caller keeps the kp rows with the
// smallest L; if L[kp-1] > L[k-1] + W + 2E no other row can be among the exact top-k, `pq_rescore_kernel`
// recomputes the kp rows exactly as oracle.c does (residual, f32x8 tree entries, sequential sum) and the final
// top-k is selected on those; queries that cannot be proven are redone by the exact kernels.  Final ids and
// distances are bit-identical to the exact path.
#include "kernels.cuh"

#include <math_constants.h>

namespace lgpu {

namespace {

// ---- table entries for the FILTER.  Not lance's arithmetic (that is pq_rescore_kernel's job): the expansion
//     |q_i - b|^2 = |q_i|^2 + |b|^2 - 2 q_i.b      (1 - q_i.b for dot)
// with |b|^2 precomputed at open and the dot product as packed FFMA2 over (even, odd) dimension pairs: ~10
// instructions per entry instead of ~25.  Its rounding error, <= ~12 u (|q_i| + |b|)^2 per entry, is part of the band
// (band_check3: the 2 (|q|^2 + CB2) term).  Both table passes call this one function, so they see identical values.
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t pack2(float a, float b)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float sum2(uint64_t v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}
template <int DSUB>
struct SubVec {
    static constexpr int NP = (DSUB + 1) / 2;
    uint64_t p[NP];             // (even, odd) dimension pairs; an odd DSUB pads with 0
    __device__ __forceinline__ void load(const float *src)
    {
        if constexpr (DSUB % 4 == 0) {
#pragma unroll
            for (int i = 0; i < DSUB / 4; i++) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
                p[2 * i] = pack2(v.x, v.y); p[2 * i + 1] = pack2(v.z, v.w);
            }
        } else if constexpr (DSUB == 2) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(src));
            p[0] = pack2(v.x, v.y);
        } else {
            p[0] = pack2(__ldg(src), 0.f);
        }
    }
    __device__ __forceinline__ float dot(const SubVec &o) const
    {
        uint64_t acc = 0ull;
#pragma unroll
        for (int i = 0; i < NP; i++) acc = fma2(p[i], o.p[i], acc);
        return sum2(acc);
    }
};
template <int DSUB, bool DOT>
__device__ __forceinline__ float filter_entry(const SubVec<DSUB> &q, float qi2, const SubVec<DSUB> &cb, float cbn2)
{
    const float d = q.dot(cb);
    return DOT ? 1.0f - d : fmaf(-2.0f, d, qi2 + cbn2);
}

// thread mapping of both table passes: grid (ceil(B/8), nch), 256 threads; warp g = query 8 bx + g; lane =
// (s = lane & 7, cq = lane >> 3); step t handles code c = 4 t + cq, so a warp reads 4 codes x 8 sub-spaces of the
// tiled codebook (1 KB contiguous) and writes 4 codes x 16 bytes of the output (64 bytes contiguous).

// ---- pass 1: min_c / max_c of T_q[i][.] for every (query, sub-space)
template <int DSUB, bool DOT>
__global__ void __launch_bounds__(256) qtable_minmax_kernel(const float *__restrict__ Q, const float *__restrict__ cb_tiled,
                                                            const float *__restrict__ cb_n2, uint32_t B, uint32_t dim,
                                                            uint32_t m, uint32_t nch, float *__restrict__ mm)
{
    const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5), ch = blockIdx.y;
    const int lane = threadIdx.x & 31, s = lane & 7, cq = lane >> 3;
    if (q >= B) return;
    const uint32_t i = ch * 8 + s;
    float mn = 0.f, mx = 0.f;
    if (i < m) {
        SubVec<DSUB> qv;
        qv.load(Q + (size_t)q * dim + i * DSUB);
        const float qi2 = qv.dot(qv);
        mn = CUDART_INF_F; mx = -CUDART_INF_F;
        bool nan = false;
        const float *cb = cb_tiled + (((size_t)ch * 256 + cq) * 8 + s) * DSUB;
        const float *n2 = cb_n2 + ((size_t)ch * 256 + cq) * 8 + s;
#pragma unroll 4
        for (int t = 0; t < 64; t++) {
            SubVec<DSUB> cv;
            cv.load(cb + (size_t)t * 32 * DSUB);
            const float v = filter_entry<DSUB, DOT>(qv, qi2, cv, __ldg(n2 + t * 32));
            mn = fminf(mn, v); mx = fmaxf(mx, v);
            nan |= v != v;                              // fminf / fmaxf drop NaN
        }
        if (nan) { mn = CUDART_NAN_F; mx = CUDART_NAN_F; }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const float on = __shfl_xor_sync(0xffffffffu, mn, o), ox = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = (on != on || mn != mn) ? CUDART_NAN_F : fminf(mn, on);
        mx = (ox != ox || mx != mx) ? CUDART_NAN_F : fmaxf(mx, ox);
    }
    if (cq == 0) reinterpret_cast<float2 *>(mm)[(size_t)q * nch * 8 + i] = make_float2(mn, mx);
}

// ---- pass 2: quantise.  Position j of code c's 16-byte output holds sub-space (j + c) & 7 (the rotation scan3.cu's
// stagers rely on), i.e. sub-space s goes to position (s - c) & 7.
template <int DSUB, bool DOT>
__global__ void __launch_bounds__(256) qtable_quant_kernel(const float *__restrict__ Q, const float *__restrict__ cb_tiled,
                                                           const float *__restrict__ cb_n2, uint32_t B, uint32_t dim,
                                                           uint32_t m, uint32_t nch, const float *__restrict__ mm,
                                                           unsigned short *__restrict__ qt, float *__restrict__ step_out,
                                                           float *__restrict__ base_out, float *__restrict__ sbound_out,
                                                           uint32_t *__restrict__ bad_out)
{
    const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5), ch = blockIdx.y;
    const int lane = threadIdx.x & 31, s = lane & 7, cq = lane >> 3;
    if (q >= B) return;
    const uint32_t m8 = nch * 8;
    const float qmax = (float)(65535u / m);
    const float2 *row = reinterpret_cast<const float2 *>(mm) + (size_t)q * m8;
    // per-query step / base / bound from the min-max table (every warp of the query's 12 CTAs recomputes it)
    float rng = 0.f, base = 0.f, sb = 0.f;
    bool bad = false;
    for (uint32_t i = lane; i < m; i += 32) {
        const float2 v = row[i];
        rng = fmaxf(rng, v.y - v.x);
        base += v.x;
        sb += fmaxf(fabsf(v.x), fabsf(v.y));
        bad |= !(fabsf(v.x) < CUDART_INF_F) || !(fabsf(v.y) < CUDART_INF_F);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rng = fmaxf(rng, __shfl_xor_sync(0xffffffffu, rng, o));
        base += __shfl_xor_sync(0xffffffffu, base, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    bad = __any_sync(0xffffffffu, bad) || !(rng < CUDART_INF_F) || !(fabsf(base) < CUDART_INF_F) || !(sb < CUDART_INF_F);
    float step = (!bad && rng > 0.f) ? rng / qmax : 0.f;
    float inv = (step > 0.f && qmax / rng < CUDART_INF_F) ? qmax / rng : 0.f;
    if (rng > 0.f && inv == 0.f) bad = true;            // a range too small to invert: take the exact path
    if (bad) { step = 0.f; inv = 0.f; }
    if (ch == 0 && lane == 0) {
        step_out[q] = step;
        bad_out[q] = bad ? 1u : 0u;
        base_out[q] = DOT ? base - (float)(m - 1) : base;
        sbound_out[q] = sb;
    }
    const uint32_t i = ch * 8 + s;
    unsigned short *out = qt + ((size_t)q * nch + ch) * 2048;          // [256 codes][8 positions]
    if (i >= m) {                                                       // padding sub-space: every entry is 0
        for (int t = 0; t < 64; t++) { const uint32_t c = 4 * t + cq; out[c * 8 + ((s - c) & 7)] = 0; }
        return;
    }
    SubVec<DSUB> qv;
    qv.load(Q + (size_t)q * dim + i * DSUB);
    const float qi2 = qv.dot(qv);
    const float mn = row[i].x;
    const float *cb = cb_tiled + (((size_t)ch * 256 + cq) * 8 + s) * DSUB;
    const float *n2 = cb_n2 + ((size_t)ch * 256 + cq) * 8 + s;
#pragma unroll 4
    for (int t = 0; t < 64; t++) {
        SubVec<DSUB> cv;
        cv.load(cb + (size_t)t * 32 * DSUB);
        const float v = filter_entry<DSUB, DOT>(qv, qi2, cv, __ldg(n2 + t * 32));
        float x = (v - mn) * inv;
        x = x >= 0.f ? fminf(floorf(x), qmax) : 0.f;                    // NaN -> 0 (the query is flagged bad)
        const uint32_t c = 4 * t + cq;
        out[c * 8 + ((s - c) & 7)] = (unsigned short)x;
    }
}

// |b|^2 of every tiled codebook entry (open time), same pairing as SubVec::dot
template <int DSUB>
__global__ void cb_norms_kernel(const float *__restrict__ cb_tiled, uint64_t n, float *__restrict__ out)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    SubVec<DSUB> v;
    v.load(cb_tiled + e * DSUB);
    out[e] = v.dot(v);
}

__device__ __forceinline__ uint32_t find_partition(const uint64_t *__restrict__ part_off, uint32_t nlist, uint64_t pos)
{
    uint32_t lo = 0, hi = nlist - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (part_off[mid + 1] > pos) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ uint32_t stream_code(const unsigned char *__restrict__ codes, uint64_t base, uint32_t npad,
                                                uint32_t row, uint32_t i)
{
    const uint32_t M = i + (row & 7);                   // position in the row's skewed byte stream (retile.cu)
    return codes[base + ((uint64_t)(M >> 3) * npad + row) * 8 + (M & 7)];
}

// R[pos] = 2 * sum_i codebook_i[code_i] . c_p,i ; rmax = max |R| (as int bits)
template <int DSUB>
__global__ void row_const_kernel(const unsigned char *__restrict__ codes, const uint64_t *__restrict__ code_base,
                                 const uint32_t *__restrict__ part_npad, const uint64_t *__restrict__ part_off,
                                 uint32_t nlist, uint64_t nrows, const float *__restrict__ centroids,
                                 const float *__restrict__ cb_tiled, uint32_t dim, uint32_t m, float *__restrict__ R,
                                 int *__restrict__ rmax_bits)
{
    const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= nrows) return;
    const uint32_t p = find_partition(part_off, nlist, pos);
    const uint32_t row = (uint32_t)(pos - part_off[p]);
    const float *cen = centroids + (size_t)p * dim;
    double acc = 0.0;                                   // f64: keeps R within 1 ulp(f32) of the real value
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t c = stream_code(codes, code_base[p], part_npad[p], row, i);
        const float *cb = cb_tiled + (((size_t)(i >> 3) * 256 + c) * 8 + (i & 7)) * DSUB;
#pragma unroll
        for (int t = 0; t < DSUB; t++) acc = fma((double)cb[t], (double)cen[i * DSUB + t], acc);
    }
    const float r = (float)(2.0 * acc);
    R[pos] = r;
    atomicMax(rmax_bits, __float_as_int(fabsf(r)));
}

// exact PQ distance of one (query, stored row) pair, exactly as oracle.c::partition_distances:
// residual -> sub-vector table entry (l2_once tree for dsub 8/16) -> sequential f32 sum -> metric scale.
// Warp-cooperative: lane l evaluates the entries of sub-spaces l, l + 32, ... (independent loads of the code byte,
// the codeword and the query / centroid sub-vectors), then the entries are summed in order i = 0..m-1 (the oracle's
// order) by walking them through a shuffle.  m <= 512.  Every lane returns the distance.
template <int DSUB>
__device__ __forceinline__ float exact_pq_distance_warp(const float *__restrict__ qv, const float *__restrict__ cen,
                                                        const unsigned char *__restrict__ codes, uint64_t cbase,
                                                        uint32_t npad, uint32_t row, const float *__restrict__ cb_tiled,
                                                        uint32_t m, int metric, int lane)
{
    float tv[16];
#pragma unroll
    for (int it = 0; it < 16; it++) {
        tv[it] = 0.f;
        const uint32_t i = (uint32_t)it * 32 + lane;
        if ((uint32_t)it * 32 < m && i < m) {
            const uint32_t c = stream_code(codes, cbase, npad, row, i);
            const float *cb = cb_tiled + (((size_t)(i >> 3) * 256 + c) * 8 + (i & 7)) * DSUB;
            float r[DSUB], cv[DSUB];
#pragma unroll
            for (int t = 0; t < DSUB; t++) {
                cv[t] = cb[t];
                r[t] = (metric == LGPU_DOT) ? qv[i * DSUB + t] : __fsub_rn(qv[i * DSUB + t], cen[i * DSUB + t]);
            }
            tv[it] = (metric == LGPU_DOT) ? subvec_dot_dist<DSUB>(r, cv) : subvec_l2<DSUB>(r, cv);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int it = 0; it < 16; it++) {
        if ((uint32_t)it * 32 < m) {
            const uint32_t lim = min(32u, m - (uint32_t)it * 32);
            for (uint32_t l = 0; l < lim; l++) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, tv[it], (int)l));
        }
    }
    if (metric == LGPU_COSINE) acc = __fmul_rn(acc, 0.5f);
    else if (metric == LGPU_DOT) acc = __fsub_rn(acc, (float)(m - 1));
    return acc;
}

// one warp per (query, candidate position) pair
template <int DSUB>
__global__ void __launch_bounds__(256) pq_rescore_kernel(const float *__restrict__ Q, const uint64_t *__restrict__ pos,
                                                         uint32_t B, uint32_t nc, const unsigned char *__restrict__ codes,
                                                         const uint64_t *__restrict__ code_base,
                                                         const uint32_t *__restrict__ part_npad,
                                                         const uint64_t *__restrict__ part_off, uint32_t nlist,
                                                         const float *__restrict__ centroids,
                                                         const float *__restrict__ cb_tiled, uint32_t dim, uint32_t m,
                                                         int metric, float *__restrict__ out)
{
    const uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pair >= (uint64_t)B * nc) return;
    const uint64_t ps = pos[pair];
    if (ps == UINT64_MAX) { if (lane == 0) out[pair] = CUDART_INF_F; return; }
    const uint32_t q = (uint32_t)(pair / nc);
    const uint32_t p = find_partition(part_off, nlist, ps);
    const uint32_t row = (uint32_t)(ps - part_off[p]);
    const float d = exact_pq_distance_warp<DSUB>(Q + (size_t)q * dim, centroids + (size_t)p * dim, codes, code_base[p],
                                                 part_npad[p], row, cb_tiled, m, metric, lane);
    if (lane == 0) out[pair] = d;
}

// ---- candidate mode of the filter scan: per-query band, then the final exact top-k over the survivors ----
__global__ void cand_prepare_kernel(const float *__restrict__ step, const float *__restrict__ sbound,
                                    const float *__restrict__ amax, const int *__restrict__ rmax_bits,
                                    const float *__restrict__ qn2, float cb2, float scale, uint32_t m, uint32_t B,
                                    float *__restrict__ slack, uint32_t *__restrict__ thr, uint32_t *__restrict__ cand_cnt)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    const float mag = sbound[q] + (amax ? amax[q] : 0.f) + (rmax_bits ? __int_as_float(*rmax_bits) : 0.f) + (float)m +
                      2.0f * (qn2[q] + cb2);
    const float E = 3.0517578125e-5f * (float)((m + 95u) / 96u) * mag;       // same band as band_check3_kernel
    const float W = (float)m * step[q] * 1.0009765625f;
    slack[q] = scale * (W + 2.0f * E);
    thr[q] = CAND_NO_THR;
    cand_cnt[q] = 0u;
}

constexpr int FIN_THREADS = 256;
template <int DSUB>
__global__ void __launch_bounds__(FIN_THREADS) cand_finalize_kernel(FinalizeArgs a)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    uint64_t *s_id = reinterpret_cast<uint64_t *>(fsm);
    uint64_t *s_pos = s_id + a.cand_cap;
    uint32_t *s_key = reinterpret_cast<uint32_t *>(s_pos + a.cand_cap);
    __shared__ uint32_t s_n;
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t total = a.cand_cnt[q];
    const uint32_t n = min(total, a.cand_cap);
    if (tid == 0) {
        s_n = 0;
        a.flags[q] = (total > a.cand_cap || a.bad[q]) ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t tkey = a.thr[q];
    const float lim = tkey == CAND_NO_THR ? CUDART_INF_F : key_f32(tkey) + a.slack[q];
    const CandRec *cand = a.cand + (size_t)q * a.cand_cap;
    const float *qv = a.Q + (size_t)q * a.dim;
    for (uint32_t c = w; c < n; c += FIN_THREADS / 32) {
        const CandRec rec = cand[c];
        if (!(rec.lb <= lim)) continue;                     // the threshold tightened after this row was appended
        float d = exact_pq_distance_warp<DSUB>(qv, a.centroids + (size_t)rec.p * a.dim, a.codes, a.code_base[rec.p],
                                               a.part_npad[rec.p], rec.row, a.cb_tiled, a.m, a.metric, lane);
        if (d != d) continue;                               // FilterExec: _distance IS NOT NULL
        if (d == 0.f) d = 0.f;                              // -0 and +0 tie
        if (lane == 0) {
            const uint32_t at = atomicAdd(&s_n, 1u);
            const uint64_t ps = a.part_off[rec.p] + rec.row;
            s_key[at] = f32_key(d); s_id[at] = a.row_ids[ps]; s_pos[at] = ps;
        }
    }
    __syncthreads();
    const uint32_t cnt = s_n;
    uint32_t n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (uint32_t i = cnt + tid; i < n2; i += FIN_THREADS) { s_key[i] = 0xffffffffu; s_id[i] = UINT64_MAX; s_pos[i] = UINT64_MAX; }
    __syncthreads();
    for (uint32_t size = 2; size <= n2; size <<= 1) {       // bitonic sort by (distance key, row id), ascending
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < (n2 >> 1); i += FIN_THREADS) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint32_t ka = s_key[lo], kb = s_key[hi];
                const uint64_t ia = s_id[lo], ib = s_id[hi];
                const bool gt = kb < ka || (kb == ka && ib < ia);
                if (gt == asc) {
                    s_key[lo] = kb; s_key[hi] = ka; s_id[lo] = ib; s_id[hi] = ia;
                    const uint64_t pa = s_pos[lo]; s_pos[lo] = s_pos[hi]; s_pos[hi] = pa;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < a.k; i += FIN_THREADS) {
        const bool have = i < cnt;
        a.out_ids[(size_t)q * a.k + i] = have ? s_id[i] : UINT64_MAX;
        a.out_dist[(size_t)q * a.k + i] = have ? key_f32(s_key[i]) : CUDART_INF_F;
        if (a.out_pos) a.out_pos[(size_t)q * a.k + i] = have ? s_pos[i] : UINT64_MAX;
    }
    if (tid == 0) a.out_count[q] = min(cnt, a.k);
}

// probe_A[slot] = coarse_dist - |q|^2 ; amax[q] = max_j coarse + |q|^2
__global__ void probe_terms_kernel(const float *__restrict__ probe_dist, const float *__restrict__ Q, uint32_t B,
                                   uint32_t nprobes, uint32_t dim, float *__restrict__ probe_A,
                                   float *__restrict__ amax, float *__restrict__ qn2)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;    // one warp per query
    const int lane = threadIdx.x & 31;
    if (q >= B) return;
    double n2d = 0.0;                                   // f64: |q|^2 within 1 ulp(f32)
    for (uint32_t t = lane; t < dim; t += 32) { double v = Q[(size_t)q * dim + t]; n2d = fma(v, v, n2d); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2d += __shfl_xor_sync(0xffffffffu, n2d, o);
    const float n2 = (float)n2d;
    float mx = 0.f;
    if (probe_A) {                                       // (dot: no residual, A = 0)
        for (uint32_t j = lane; j < nprobes; j += 32) {
            const float cd = probe_dist[(size_t)q * nprobes + j];
            probe_A[(size_t)q * nprobes + j] = cd - n2;
            mx = fmaxf(mx, fabsf(cd));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
        qn2[q] = n2;
        if (amax) amax[q] = mx + n2;                    // >= |A| and also covers the coarse distance's own rounding
    }
}

// flags[q] = 1 when the shortlist cannot be proven to contain the exact top-k (see the header and kernels.cuh).
// Error budget, u = 2^-24: the exact (oracle-order) distance d* differs from the real-arithmetic distance D of the
// same f32 inputs by <= (m + 16) u (D + |q - c_p|^2); the table entries carry <= (dsub + 2) u T each, the floor of
// the quantiser can be off by one step when (T - min) / step lands within 2^-11 of an integer (absorbed by the
// factor 1 + 2^-10 on W), A carries <= 70 u (coarse + |q|^2), R one ulp, the epilogue of scan3 three more roundings
// of values bounded by sbound + amax + rmax, and the expansion form of the entries (filter_entry) <= 12 u (|q_i| + |b|)^2
// <= 24 u (|q_i|^2 + max_c |b|^2) each.  For m <= 96 all of it is < 2^9 u = 2^-15 of
// (sbound + amax + rmax + m + 2 (|q|^2 + CB2)), CB2 = sum_i max_c |codebook_i[c]|^2; larger m widens E proportionally.
__global__ void band_check3_kernel(const float *__restrict__ lb, const uint32_t *__restrict__ cnt,
                                   const float *__restrict__ step, const float *__restrict__ sbound,
                                   const float *__restrict__ amax, const int *__restrict__ rmax_bits,
                                   const uint32_t *__restrict__ bad, const float *__restrict__ qn2, float cb2, float scale,
                                   uint32_t m, uint32_t B, uint32_t k, uint32_t kp, uint32_t *__restrict__ flags)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    uint32_t f = bad[q] ? 1u : 0u;
    if (!f && cnt[q] >= kp && kp > 0) {
        const float mag = sbound[q] + (amax ? amax[q] : 0.f) + (rmax_bits ? __int_as_float(*rmax_bits) : 0.f) + (float)m +
                          2.0f * (qn2[q] + cb2);
        const float E = 3.0517578125e-5f * (float)((m + 95u) / 96u) * mag;
        const float W = (float)m * step[q] * 1.0009765625f;
        const float kth = lb[(size_t)q * kp + (k - 1 < kp ? k - 1 : kp - 1)];
        const float last = lb[(size_t)q * kp + kp - 1];
        f = (k >= kp || !(last > kth + scale * (W + 2.0f * E))) ? 1u : 0u;
    }
    flags[q] = f;
}

template <class F> void dispatch_dsub(uint32_t dsub, F &&f)
{
    switch (dsub) {
    case 1: f(std::integral_constant<int, 1>{}); break;
    case 2: f(std::integral_constant<int, 2>{}); break;
    case 4: f(std::integral_constant<int, 4>{}); break;
    case 8: f(std::integral_constant<int, 8>{}); break;
    case 16: f(std::integral_constant<int, 16>{}); break;
    case 32: f(std::integral_constant<int, 32>{}); break;
    default: set_error("unsupported PQ sub-vector length"); throw Failure{LGPU_INVALID_INPUT};
    }
}

}  // namespace

void launch_query_tables_q16(const float *Q, const float *cb_tiled, const float *cb_n2, uint32_t B, uint32_t dim,
                             uint32_t m, uint32_t nch, uint32_t dsub, int metric, float *mm, uint4 *qt, float *step,
                             float *base, float *sbound, uint32_t *bad, cudaStream_t st)
{
    if (B == 0) return;
    const dim3 grid((B + 7) / 8, nch);
    unsigned short *qt16 = reinterpret_cast<unsigned short *>(qt);
    dispatch_dsub(dsub, [&](auto D) {
        constexpr int DS = decltype(D)::value;
        if (metric == LGPU_DOT) {
            qtable_minmax_kernel<DS, true><<<grid, 256, 0, st>>>(Q, cb_tiled, cb_n2, B, dim, m, nch, mm); LGPU_COUNT_LAUNCH();
            qtable_quant_kernel<DS, true><<<grid, 256, 0, st>>>(Q, cb_tiled, cb_n2, B, dim, m, nch, mm, qt16, step, base, sbound, bad); LGPU_COUNT_LAUNCH();
        } else {
            qtable_minmax_kernel<DS, false><<<grid, 256, 0, st>>>(Q, cb_tiled, cb_n2, B, dim, m, nch, mm); LGPU_COUNT_LAUNCH();
            qtable_quant_kernel<DS, false><<<grid, 256, 0, st>>>(Q, cb_tiled, cb_n2, B, dim, m, nch, mm, qt16, step, base, sbound, bad); LGPU_COUNT_LAUNCH();
        }
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_cb_norms(const float *cb_tiled, uint32_t nch, uint32_t dsub, float *out, cudaStream_t st)
{
    const uint64_t n = (uint64_t)nch * 256 * 8;
    dispatch_dsub(dsub, [&](auto D) {
        cb_norms_kernel<decltype(D)::value><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cb_tiled, n, out); LGPU_COUNT_LAUNCH();
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_row_const(const unsigned char *codes, const uint64_t *code_base, const uint32_t *part_npad,
                      const uint64_t *part_off, uint32_t nlist, uint64_t nrows, const float *centroids,
                      const float *cb_tiled, uint32_t dim, uint32_t m, uint32_t dsub, float *R, int *rmax_bits,
                      cudaStream_t st)
{
    LGPU_CUDA(cudaMemsetAsync(rmax_bits, 0, sizeof(int), st));
    if (nrows == 0) return;
    dispatch_dsub(dsub, [&](auto D) {
        row_const_kernel<decltype(D)::value><<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(
            codes, code_base, part_npad, part_off, nlist, nrows, centroids, cb_tiled, dim, m, R, rmax_bits); LGPU_COUNT_LAUNCH();
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_pq_rescore(const float *Q, const uint64_t *pos, uint32_t B, uint32_t nc, const unsigned char *codes,
                       const uint64_t *code_base, const uint32_t *part_npad, const uint64_t *part_off, uint32_t nlist,
                       const float *centroids, const float *cb_tiled, uint32_t dim, uint32_t m, uint32_t dsub, int metric,
                       float *out, cudaStream_t st)
{
    if (B == 0 || nc == 0) return;
    if (m > 512) { set_error("internal: pq_rescore supports m <= 512"); throw Failure{LGPU_RUNTIME}; }
    const uint64_t total = (uint64_t)B * nc * 32;              // one warp per pair
    dispatch_dsub(dsub, [&](auto D) {
        pq_rescore_kernel<decltype(D)::value><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            Q, pos, B, nc, codes, code_base, part_npad, part_off, nlist, centroids, cb_tiled, dim, m, metric, out); LGPU_COUNT_LAUNCH();
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_probe_terms(const float *probe_dist, const float *Q, uint32_t B, uint32_t nprobes, uint32_t dim,
                        float *probe_A, float *amax, float *qn2, cudaStream_t st)
{
    if (B == 0) return;
    probe_terms_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(probe_dist, Q, B, nprobes, dim, probe_A, amax, qn2); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_cand_prepare(const float *step, const float *sbound, const float *amax, const int *rmax_bits, const float *qn2,
                         float cb2, float scale, uint32_t m, uint32_t B, float *slack, uint32_t *thr, uint32_t *cand_cnt,
                         cudaStream_t st)
{
    if (B == 0) return;
    cand_prepare_kernel<<<(B + 127) / 128, 128, 0, st>>>(step, sbound, amax, rmax_bits, qn2, cb2, scale, m, B, slack, thr, cand_cnt); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_cand_finalize(const FinalizeArgs &a, cudaStream_t st)
{
    if (a.B == 0) return;
    if (a.m > 512 || a.cand_cap < 2 || (a.cand_cap & (a.cand_cap - 1)) || a.k > a.cand_cap) {
        set_error("internal: cand_finalize needs m <= 512 and a power-of-two candidate capacity >= k");
        throw Failure{LGPU_RUNTIME};
    }
    const size_t smem = (size_t)a.cand_cap * 20;
    dispatch_dsub(a.dsub, [&](auto D) {
        auto kern = cand_finalize_kernel<decltype(D)::value>;
        LGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<a.B, FIN_THREADS, smem, st>>>(a); LGPU_COUNT_LAUNCH();
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_band_check3(const float *lb, const uint32_t *cnt, const float *step, const float *sbound, const float *amax,
                        const int *rmax_bits, const uint32_t *bad, const float *qn2, float cb2, float scale, uint32_t m,
                        uint32_t B, uint32_t k, uint32_t kp, uint32_t *flags, cudaStream_t st)
{
    if (B == 0) return;
    band_check3_kernel<<<(B + 127) / 128, 128, 0, st>>>(lb, cnt, step, sbound, amax, rmax_bits, bad, qn2, cb2, scale, m, B, k, kp, flags); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
