// tables.cu -- the two-pass form of the PQ scan (filter + verify).
//
// The exact path (scan.cu) builds one 96 KB distance table per (query, probed partition),
// because lance's table is on the *residual* q - c_p: 20 480
// tables per 1024-query batch, 23 f32 ops per entry, and that table build -- not the scan --
// bounds the kernel.  Algebraically
//     |(q_i - c_i) - b_i|^2 = |q_i - b_i|^2  +  (|c_i|^2 - 2 q_i.c_i)  +  2 b_i.c_i
// so the distance of row r of partition p is  S_q(r) + A(q,p) + R(r)  with
//     S_q(r) = sum_i T_q[i][code_i(r)],  T_q[i][c] = |q_i - codebook_i[c]|^2    one table per QUERY
//     A(q,p) = |q - c_p|^2 - |q|^2                                               from the coarse step
//     R(r)   = 2 * sum_i codebook_i[code_i(r)] . c_p,i                           one f32 per row, at open
// The rounding differs from lance's, so this is only used as a FILTER: the scan kernel's
// approximate pass ranks rows by S+A+R, the caller keeps every row within a rigorous error
// band of the k-th best (`launch_band_check2`), `pq_rescore_kernel` recomputes those few rows
// exactly as oracle.c does (residual, f32x8 tree entries, sequential sum), and queries whose
// band overflowed the shortlist are redone by the exact kernels.  Final ids/distances are
// bit-identical to the exact path; the table work drops ~20x.
#include "kernels.cuh"

#include <math_constants.h>

namespace lgpu {

namespace {

// T_q[ch][c][s] for every query: one thread per entry (coalesced on the tiled codebook)
template <int DSUB>
__global__ void query_tables_kernel(const float *__restrict__ Q, const float *__restrict__ cb_tiled, uint32_t B,
                                    uint32_t dim, uint32_t m, uint32_t nch, int metric, float *__restrict__ T)
{
    const uint64_t per_q = (uint64_t)nch * 256 * 8;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)B * per_q) return;
    const uint32_t q = (uint32_t)(idx / per_q);
    const uint64_t e = idx - (uint64_t)q * per_q;
    const uint32_t s = e & 7, c = (e >> 3) & 255, ch = (uint32_t)(e >> 11);
    const uint32_t i = ch * 8 + s;
    float v = 0.f;
    if (i < m) {
        float qv[DSUB], cv[DSUB];
        const float *qp = Q + (size_t)q * dim + i * DSUB, *cp = cb_tiled + e * DSUB;
        if constexpr (DSUB % 4 == 0) {
#pragma unroll
            for (int t = 0; t < DSUB; t += 4) {
                const float4 a4 = __ldg(reinterpret_cast<const float4 *>(qp + t));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(cp + t));
                qv[t] = a4.x; qv[t + 1] = a4.y; qv[t + 2] = a4.z; qv[t + 3] = a4.w;
                cv[t] = b4.x; cv[t + 1] = b4.y; cv[t + 2] = b4.z; cv[t + 3] = b4.w;
            }
        } else {
#pragma unroll
            for (int t = 0; t < DSUB; t++) { qv[t] = qp[t]; cv[t] = cp[t]; }
        }
        v = subvec_l2<DSUB>(qv, cv);
    }
    T[idx] = v;
    (void)metric;
}

// bound[q] = sum_i max_c T_q[i][c]  (largest possible S), one warp per (query)
__global__ void table_bound_kernel(const float *__restrict__ T, uint32_t B, uint32_t nch, float *__restrict__ bound)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= B) return;
    const float *t = T + (size_t)q * nch * 256 * 8;
    float total = 0.f;
    for (uint32_t ch = 0; ch < nch; ch++) {
        // lane handles sub-space s = lane & 7 for codes c = lane>>3, +4, ...
        float mx = 0.f;
        for (uint32_t c = lane >> 3; c < 256; c += 4) mx = fmaxf(mx, t[((size_t)ch * 256 + c) * 8 + (lane & 7)]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        // lanes 0..7 now hold the max of sub-space s; add the 8 of them
        float sm = mx;
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        sm += __shfl_xor_sync(0xffffffffu, sm, 4);
        total += sm;
    }
    if (lane == 0) bound[q] = total;
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
// residual -> sub-vector table entry (l2_once tree for dsub 8/16) -> sequential f32 sum -> metric scale
template <int DSUB>
__global__ void pq_rescore_kernel(const float *__restrict__ Q, const uint64_t *__restrict__ pos, uint32_t B, uint32_t nc,
                                  const unsigned char *__restrict__ codes, const uint64_t *__restrict__ code_base,
                                  const uint32_t *__restrict__ part_npad, const uint64_t *__restrict__ part_off,
                                  uint32_t nlist, const float *__restrict__ centroids,
                                  const float *__restrict__ cb_tiled, uint32_t dim, uint32_t m, int metric,
                                  float *__restrict__ out)
{
    const uint64_t pair = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= (uint64_t)B * nc) return;
    const uint64_t ps = pos[pair];
    if (ps == UINT64_MAX) { out[pair] = CUDART_INF_F; return; }
    const uint32_t q = (uint32_t)(pair / nc);
    const uint32_t p = find_partition(part_off, nlist, ps);
    const uint32_t row = (uint32_t)(ps - part_off[p]);
    const float *qv = Q + (size_t)q * dim, *cen = centroids + (size_t)p * dim;
    float acc = 0.f;
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t c = stream_code(codes, code_base[p], part_npad[p], row, i);
        const float *cb = cb_tiled + (((size_t)(i >> 3) * 256 + c) * 8 + (i & 7)) * DSUB;
        float r[DSUB], cv[DSUB];
#pragma unroll
        for (int t = 0; t < DSUB; t++) {
            cv[t] = cb[t];
            r[t] = (metric == LGPU_DOT) ? qv[i * DSUB + t] : __fsub_rn(qv[i * DSUB + t], cen[i * DSUB + t]);
        }
        const float e = (metric == LGPU_DOT) ? subvec_dot_dist<DSUB>(r, cv) : subvec_l2<DSUB>(r, cv);
        acc = __fadd_rn(acc, e);
    }
    if (metric == LGPU_COSINE) acc = __fmul_rn(acc, 0.5f);
    else if (metric == LGPU_DOT) acc = __fsub_rn(acc, (float)(m - 1));
    out[pair] = acc;
}

// probe_A[slot] = coarse_dist - |q|^2 ; amax[q] = max_j |probe_A|
__global__ void probe_terms_kernel(const float *__restrict__ probe_dist, const float *__restrict__ Q, uint32_t B,
                                   uint32_t nprobes, uint32_t dim, float *__restrict__ probe_A, float *__restrict__ qn2,
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
    if (lane == 0) {
        qn2[q] = n2;
        amax[q] = mx + n2;                              // >= |A| and also covers the coarse distance's own rounding
    }
}

// flags[q] = 1 when the approximate shortlist cannot be proven to contain the exact top-k.
// |approx - exact| <= E_q = 2^-16 (Smax_q + Amax_q + Rmax) * scale: with u = 2^-24, the approximate sum
// carries <= (m + 4) u S (sequential adds + table entries), the coarse distance <= 66 u, A and R a few u
// each, and the exact value itself <= (m + 5) u d with d <= S + |A| + |R|; for m <= 96 that is < 256 u of
// the bound (m > 96 widens the band proportionally, see launch_band_check2).
__global__ void band_check2_kernel(const float *__restrict__ approx, const uint32_t *__restrict__ cnt,
                                   const float *__restrict__ sbound, const float *__restrict__ amax,
                                   const int *__restrict__ rmax_bits, float scale, uint32_t B, uint32_t k, uint32_t kp,
                                   uint32_t *__restrict__ flags)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    uint32_t f = 0;
    if (cnt[q] >= kp && kp > 0) {
        const float E = 1.52587890625e-5f * (sbound[q] + amax[q] + __int_as_float(*rmax_bits)) * scale;
        const float kth = approx[(size_t)q * kp + (k - 1 < kp ? k - 1 : kp - 1)];
        const float last = approx[(size_t)q * kp + kp - 1];
        f = (k >= kp || !(last > kth + 2.0f * E)) ? 1u : 0u;
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

void launch_query_tables(const float *Q, const float *cb_tiled, uint32_t B, uint32_t dim, uint32_t m, uint32_t nch,
                         uint32_t dsub, int metric, float *T, float *sbound, cudaStream_t st)
{
    if (B == 0) return;
    const uint64_t total = (uint64_t)B * nch * 256 * 8;
    dispatch_dsub(dsub, [&](auto D) {
        query_tables_kernel<decltype(D)::value><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Q, cb_tiled, B, dim, m, nch, metric, T); LGPU_COUNT_LAUNCH();
    });
    table_bound_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(T, B, nch, sbound); LGPU_COUNT_LAUNCH();
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
    const uint64_t total = (uint64_t)B * nc;
    dispatch_dsub(dsub, [&](auto D) {
        pq_rescore_kernel<decltype(D)::value><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
            Q, pos, B, nc, codes, code_base, part_npad, part_off, nlist, centroids, cb_tiled, dim, m, metric, out); LGPU_COUNT_LAUNCH();
    });
    LGPU_CUDA(cudaGetLastError());
}

void launch_probe_terms(const float *probe_dist, const float *Q, uint32_t B, uint32_t nprobes, uint32_t dim,
                        float *probe_A, float *qn2, float *amax, cudaStream_t st)
{
    if (B == 0) return;
    probe_terms_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(probe_dist, Q, B, nprobes, dim, probe_A, qn2, amax); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_band_check2(const float *approx, const uint32_t *cnt, const float *sbound, const float *amax,
                        const int *rmax_bits, float scale, uint32_t B, uint32_t k, uint32_t kp, uint32_t *flags,
                        cudaStream_t st)
{
    if (B == 0) return;
    band_check2_kernel<<<(B + 127) / 128, 128, 0, st>>>(approx, cnt, sbound, amax, rmax_bits, scale, B, k, kp, flags); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
