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

template <int DSUB>
__device__ __forceinline__ void load_cb(float *dst, const float *src)
{
    if constexpr (DSUB % 4 == 0) {
#pragma unroll
        for (int i = 0; i < DSUB / 4; i++) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < DSUB; i++) dst[i] = __ldg(src + i);
    }
}

template <int DSUB, bool DOT>
__device__ __forceinline__ float table_entry(const float *qv, const float *cv)
{
    return DOT ? subvec_dot_dist<DSUB>(qv, cv) : subvec_l2<DSUB>(qv, cv);
}

// ---- pass 1: min_c / max_c of T_q[i][.] for every (query, sub-space).  grid (ceil(B/8), nch), 256 threads:
// warp g = query 8 bx + g; lane = (s = lane & 7, code quarter = lane >> 3).
template <int DSUB, bool DOT>
__global__ void __launch_bounds__(256) qtable_minmax_kernel(const float *__restrict__ Q, const float *__restrict__ cb_tiled,
                                                            uint32_t B, uint32_t dim, uint32_t m, uint32_t nch,
                                                            float *__restrict__ mm)
{
    const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5), ch = blockIdx.y;
    const int lane = threadIdx.x & 31, s = lane & 7, cq = lane >> 3;
    if (q >= B) return;
    const uint32_t i = ch * 8 + s;
    float mn = 0.f, mx = 0.f;
    if (i < m) {
        float qv[DSUB];
#pragma unroll
        for (int t = 0; t < DSUB; t++) qv[t] = Q[(size_t)q * dim + i * DSUB + t];
        mn = CUDART_INF_F; mx = -CUDART_INF_F;
        const float *cb = cb_tiled + (((size_t)ch * 256 + cq * 64) * 8 + s) * DSUB;
        for (int t = 0; t < 64; t++) {
            float cv[DSUB];
            load_cb<DSUB>(cv, cb + (size_t)t * 8 * DSUB);
            const float v = table_entry<DSUB, DOT>(qv, cv);
            mn = fminf(mn, v); mx = fmaxf(mx, v);
            if (v != v) { mn = v; mx = v; }             // NaN sticks (fminf/fmaxf would drop it)
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const float on = __shfl_xor_sync(0xffffffffu, mn, o), ox = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = (on != on || mn != mn) ? CUDART_NAN_F : fminf(mn, on);
        mx = (ox != ox || mx != mx) ? CUDART_NAN_F : fmaxf(mx, ox);
    }
    if (cq == 0) {
        float2 *o = reinterpret_cast<float2 *>(mm) + (size_t)q * nch * 8 + i;
        *o = make_float2(mn, mx);
    }
}

// ---- pass 2: quantise.  grid (ceil(B/8), nch), 256 threads: thread c = code c of chunk ch for the CTA's 8
// queries.  Position j of the 16-byte output holds sub-space (j + c) & 7 (the rotation scan3.cu's stagers rely on).
template <int DSUB, bool DOT>
__global__ void __launch_bounds__(256) qtable_quant_kernel(const float *__restrict__ Q, const float *__restrict__ cb_tiled,
                                                           uint32_t B, uint32_t dim, uint32_t m, uint32_t nch,
                                                           const float *__restrict__ mm, uint4 *__restrict__ qt,
                                                           float *__restrict__ step_out, float *__restrict__ base_out,
                                                           float *__restrict__ sbound_out, uint32_t *__restrict__ bad_out)
{
    __shared__ float s_q[8][8][DSUB];
    __shared__ float s_min[8][8];
    __shared__ float s_inv[8];
    const uint32_t q0 = blockIdx.x * 8, ch = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const uint32_t m8 = nch * 8;
    const float qmax = (float)(65535u / m);
    {   // per-query step / base / bound from the min-max table: warp g handles query q0 + g
        const uint32_t q = q0 + g;
        float rng = 0.f, base = 0.f, sb = 0.f;
        bool bad = false;
        if (q < B) {
            const float2 *row = reinterpret_cast<const float2 *>(mm) + (size_t)q * m8;
            for (uint32_t i = lane; i < m; i += 32) {
                const float2 v = row[i];
                rng = fmaxf(rng, v.y - v.x);
                base += v.x;
                sb += fmaxf(fabsf(v.x), fabsf(v.y));
                bad |= !(fabsf(v.x) < CUDART_INF_F) || !(fabsf(v.y) < CUDART_INF_F);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rng = fmaxf(rng, __shfl_xor_sync(0xffffffffu, rng, o));
            base += __shfl_xor_sync(0xffffffffu, base, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        bad = __any_sync(0xffffffffu, bad) || !(rng < CUDART_INF_F) || !(fabsf(base) < CUDART_INF_F) || !(sb < CUDART_INF_F);
        const float step = (!bad && rng > 0.f) ? rng / qmax : 0.f;
        const float inv = (step > 0.f && qmax / rng < CUDART_INF_F) ? qmax / rng : 0.f;
        if (lane == 0) {
            s_inv[g] = bad ? 0.f : inv;
            if (ch == 0 && q < B) {
                step_out[q] = (inv > 0.f) ? step : 0.f;      // inv == 0: every entry quantises to 0 and W = 0 would be
                bad_out[q] = (bad || (rng > 0.f && inv == 0.f)) ? 1u : 0u;   // wrong unless the range is 0 too
                base_out[q] = DOT ? base - (float)(m - 1) : base;
                sbound_out[q] = sb;
            }
        }
        if (lane < 8) {
            const uint32_t i = ch * 8 + lane;
            s_min[g][lane] = (q < B && i < m) ? reinterpret_cast<const float2 *>(mm)[(size_t)q * m8 + i].x : 0.f;
        }
        for (int t = lane; t < 8 * DSUB; t += 32) {
            const uint32_t i = ch * 8 + t / DSUB;
            s_q[g][t / DSUB][t % DSUB] = (q < B && i < m) ? Q[(size_t)q * dim + i * DSUB + t % DSUB] : 0.f;
        }
    }
    __syncthreads();
    const uint32_t c = (uint32_t)tid;
    uint32_t out[8][4];
#pragma unroll
    for (int gg = 0; gg < 8; gg++)
#pragma unroll
        for (int w = 0; w < 4; w++) out[gg][w] = 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t s = ((uint32_t)j + c) & 7u;
        const uint32_t i = ch * 8 + s;
        if (i < m) {
            float cv[DSUB];
            load_cb<DSUB>(cv, cb_tiled + (((size_t)ch * 256 + c) * 8 + s) * DSUB);
#pragma unroll
            for (int gg = 0; gg < 8; gg++) {
                float qv[DSUB];
#pragma unroll
                for (int t = 0; t < DSUB; t++) qv[t] = s_q[gg][s][t];
                const float v = table_entry<DSUB, DOT>(qv, cv);
                float x = (v - s_min[gg][s]) * s_inv[gg];
                x = x >= 0.f ? fminf(floorf(x), qmax) : 0.f;           // NaN -> 0 (the query is flagged bad)
                out[gg][j >> 1] |= (uint32_t)x << (16 * (j & 1));
            }
        }
    }
#pragma unroll
    for (int gg = 0; gg < 8; gg++)
        if (q0 + gg < B)
            qt[((size_t)(q0 + gg) * nch + ch) * 256 + c] = make_uint4(out[gg][0], out[gg][1], out[gg][2], out[gg][3]);
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

// exact PQ distance of (query, stored row) pairs, exactly as oracle.c::partition_distances:
// residual -> sub-vector table entry (l2_once tree for dsub 8/16) -> sequential f32 sum -> metric scale.
// One WARP per pair: lane l evaluates the entries of sub-spaces l, l + 32, ... (independent loads of the code
// byte, the codeword and the query / centroid sub-vectors), then the entries are summed in order i = 0..m-1
// (the oracle's order) by walking them through a shuffle.  m <= 512.
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
    const float *qv = Q + (size_t)q * dim, *cen = centroids + (size_t)p * dim;
    const uint64_t cbase = code_base[p];
    const uint32_t npad = part_npad[p];
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
    if (lane == 0) out[pair] = acc;
}

// probe_A[slot] = coarse_dist - |q|^2 ; amax[q] = max_j coarse + |q|^2
__global__ void probe_terms_kernel(const float *__restrict__ probe_dist, const float *__restrict__ Q, uint32_t B,
                                   uint32_t nprobes, uint32_t dim, float *__restrict__ probe_A,
                                   float *__restrict__ amax)
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
    for (uint32_t j = lane; j < nprobes; j += 32) {
        const float cd = probe_dist[(size_t)q * nprobes + j];
        probe_A[(size_t)q * nprobes + j] = cd - n2;
        mx = fmaxf(mx, fabsf(cd));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) amax[q] = mx + n2;                   // >= |A| and also covers the coarse distance's own rounding
}

// flags[q] = 1 when the shortlist cannot be proven to contain the exact top-k (see the header and kernels.cuh).
// Error budget, u = 2^-24: the exact (oracle-order) distance d* differs from the real-arithmetic distance D of the
// same f32 inputs by <= (m + 16) u (D + |q - c_p|^2); the table entries carry <= (dsub + 2) u T each, the floor of
// the quantiser can be off by one step when (T - min) / step lands within 2^-11 of an integer (absorbed by the
// factor 1 + 2^-10 on W), A carries <= 70 u (coarse + |q|^2), R one ulp, the epilogue of scan3 three more roundings
// of values bounded by sbound + amax + rmax.  For m <= 96 all of it is < 2^9 u = 2^-15 of (sbound + amax + rmax);
// larger m widens E proportionally.
__global__ void band_check3_kernel(const float *__restrict__ lb, const uint32_t *__restrict__ cnt,
                                   const float *__restrict__ step, const float *__restrict__ sbound,
                                   const float *__restrict__ amax, const int *__restrict__ rmax_bits,
                                   const uint32_t *__restrict__ bad, float scale, uint32_t m, uint32_t B, uint32_t k,
                                   uint32_t kp, uint32_t *__restrict__ flags)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    uint32_t f = bad[q] ? 1u : 0u;
    if (!f && cnt[q] >= kp && kp > 0) {
        const float mag = sbound[q] + (amax ? amax[q] : 0.f) + (rmax_bits ? __int_as_float(*rmax_bits) : 0.f) + (float)m;
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

void launch_query_tables_q16(const float *Q, const float *cb_tiled, uint32_t B, uint32_t dim, uint32_t m, uint32_t nch,
                             uint32_t dsub, int metric, float *mm, uint4 *qt, float *step, float *base, float *sbound,
                             uint32_t *bad, cudaStream_t st)
{
    if (B == 0) return;
    const dim3 grid((B + 7) / 8, nch);
    dispatch_dsub(dsub, [&](auto D) {
        constexpr int DS = decltype(D)::value;
        if (metric == LGPU_DOT) {
            qtable_minmax_kernel<DS, true><<<grid, 256, 0, st>>>(Q, cb_tiled, B, dim, m, nch, mm); LGPU_COUNT_LAUNCH();
            qtable_quant_kernel<DS, true><<<grid, 256, 0, st>>>(Q, cb_tiled, B, dim, m, nch, mm, qt, step, base, sbound, bad); LGPU_COUNT_LAUNCH();
        } else {
            qtable_minmax_kernel<DS, false><<<grid, 256, 0, st>>>(Q, cb_tiled, B, dim, m, nch, mm); LGPU_COUNT_LAUNCH();
            qtable_quant_kernel<DS, false><<<grid, 256, 0, st>>>(Q, cb_tiled, B, dim, m, nch, mm, qt, step, base, sbound, bad); LGPU_COUNT_LAUNCH();
        }
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
                        float *probe_A, float *amax, cudaStream_t st)
{
    if (B == 0) return;
    probe_terms_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(probe_dist, Q, B, nprobes, dim, probe_A, amax); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_band_check3(const float *lb, const uint32_t *cnt, const float *step, const float *sbound, const float *amax,
                        const int *rmax_bits, const uint32_t *bad, float scale, uint32_t m, uint32_t B, uint32_t k,
                        uint32_t kp, uint32_t *flags, cudaStream_t st)
{
    if (B == 0) return;
    band_check3_kernel<<<(B + 127) / 128, 128, 0, st>>>(lb, cnt, step, sbound, amax, rmax_bits, bad, scale, m, B, k, kp, flags); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
