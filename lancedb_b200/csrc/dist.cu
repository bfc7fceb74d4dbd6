// dist.cu -- exact f32 distance kernels in lance's rounding order.
//
// K1 (coarse): IvfModel::find_partitions computes dist(q, every centroid) with
// lance-linalg's l2 / dot [lance, recalled; SURVEY.md 8a row a3].  l2_scalar::<f32,16>
// keeps 16 lane accumulators (lane l sums dims l, l+16, ...) and then adds the 16
// lane sums sequentially; to be bit-identical the kernel keeps exactly that structure:
// the 16 K-lanes are 16 threads, every thread register-tiles 4 queries x 8 centroids
// for its lane, and one thread per (query, centroid) pair adds the 16 lane sums in order.
// The same kernel is the exact flat KNNVectorDistance and
// `launch_pair_distance` is the exact re-rank used by refine_factor (row a10).
#include "kernels.cuh"

#include <math_constants.h>

#include <algorithm>

namespace lgpu {

namespace {

constexpr int DM_Q = 16, DM_C = 32, DM_KT = 32, DM_STR = 36, DM_THREADS = 256;

__global__ void __launch_bounds__(DM_THREADS) dist_matrix_kernel(
    const float *__restrict__ Q, const float *__restrict__ C, uint32_t B, uint64_t N, uint32_t d, int mode,
    const float *__restrict__ xnorm, const float *__restrict__ ysqrt, float *__restrict__ D, uint64_t ldD,
    const uint32_t *__restrict__ only)
{
    if (only) {                                  // fix-up pass: skip query tiles with no flagged query
        bool any = false;
        for (uint32_t i = 0; i < DM_Q; i++) {
            uint32_t gq = blockIdx.y * DM_Q + i;
            if (gq < B && only[gq]) any = true;
        }
        if (!any) return;
    }
    __shared__ float qs[DM_Q][DM_STR];
    __shared__ float cs[DM_C][DM_STR];
    __shared__ float red[DM_Q * DM_C][17];

    const int tid = threadIdx.x;
    const int klane = tid & 15, pt = tid >> 4;
    const int tq = pt & 3, tc = pt >> 2;
    const uint32_t q0 = blockIdx.y * DM_Q;
    const uint32_t d16 = d & ~15u;
    const uint64_t ncol_tiles = (N + DM_C - 1) / DM_C;
  for (uint64_t ctile = blockIdx.x; ctile < ncol_tiles; ctile += gridDim.x) {     // column tiles of this CTA
    const uint64_t c0 = ctile * DM_C;
    __syncthreads();

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    for (uint32_t k0 = 0; k0 < d16; k0 += DM_KT) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int idx = tid + DM_THREADS * u, row = idx >> 5, col = idx & 31;
            uint32_t gq = q0 + row, gk = k0 + col;
            qs[row][col] = (gq < B && gk < d16) ? Q[(size_t)gq * d + gk] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int idx = tid + DM_THREADS * u, row = idx >> 5, col = idx & 31;
            uint64_t gc = c0 + row; uint32_t gk = k0 + col;
            cs[row][col] = (gc < N && gk < d16) ? C[(size_t)gc * d + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int step = 0; step < 2; step++) {
            if (k0 + 16 * step < d16) {
                const int kk = 16 * step + klane;
                float qv[4], cv[8];
#pragma unroll
                for (int i = 0; i < 4; i++) qv[i] = qs[4 * tq + i][kk];
#pragma unroll
                for (int j = 0; j < 8; j++) cv[j] = cs[8 * tc + j][kk];
                if (mode == 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float df = __fsub_rn(qv[i], cv[j]);
                            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(df, df));
                        }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 8; j++) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(qv[i], cv[j]));
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) red[(4 * tq + i) * DM_C + 8 * tc + j][klane] = acc[i][j];
    __syncthreads();
    for (int pair = tid; pair < DM_Q * DM_C; pair += DM_THREADS) {
        const int qrow = pair / DM_C, ccol = pair % DM_C;
        const uint32_t gq = q0 + qrow; const uint64_t gc = c0 + ccol;
        if (gq >= B || gc >= N) continue;
        float t = 0.f;
#pragma unroll
        for (int l = 0; l < 16; l++) t = __fadd_rn(t, red[pair][l]);
        float s = 0.f;                                   // remainder dims, sequential
        const float *x = Q + (size_t)gq * d, *y = C + (size_t)gc * d;
        for (uint32_t i = d16; i < d; i++) {
            if (mode == 0) { float df = __fsub_rn(x[i], y[i]); s = __fadd_rn(s, __fmul_rn(df, df)); }
            else s = __fadd_rn(s, __fmul_rn(x[i], y[i]));
        }
        float v = __fadd_rn(s, t);
        if (mode == 1) v = __fsub_rn(1.0f, v);
        else if (mode == 2) v = __fsub_rn(1.0f, __fdiv_rn(__fdiv_rn(v, xnorm[gq]), ysqrt[gc]));
        D[(size_t)gq * ldD + gc] = v;
    }
  }
}

// half-warp per row: lane l (< 16) is lance's accumulator lane l
__device__ __forceinline__ float halfwarp_dot(const float *__restrict__ x, const float *__restrict__ y,
                                               uint32_t d, int hl, unsigned hmask, int hbase)
{
    const uint32_t d16 = d & ~15u;
    float a = 0.f;
    for (uint32_t k = hl; k < d16; k += 16) a = __fadd_rn(a, __fmul_rn(x[k], y[k]));
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 16; l++) t = __fadd_rn(t, __shfl_sync(hmask, a, hbase + l));
    float s = 0.f;
    for (uint32_t i = d16; i < d; i++) s = __fadd_rn(s, __fmul_rn(x[i], y[i]));
    return __fadd_rn(s, t);
}
__device__ __forceinline__ float halfwarp_l2(const float *__restrict__ x, const float *__restrict__ y,
                                              uint32_t d, int hl, unsigned hmask, int hbase)
{
    const uint32_t d16 = d & ~15u;
    float a = 0.f;
    for (uint32_t k = hl; k < d16; k += 16) { float df = __fsub_rn(x[k], y[k]); a = __fadd_rn(a, __fmul_rn(df, df)); }
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 16; l++) t = __fadd_rn(t, __shfl_sync(hmask, a, hbase + l));
    float s = 0.f;
    for (uint32_t i = d16; i < d; i++) { float df = __fsub_rn(x[i], y[i]); s = __fadd_rn(s, __fmul_rn(df, df)); }
    return __fadd_rn(s, t);
}

__global__ void row_norms_kernel(const float *__restrict__ X, uint64_t n, uint32_t d, float *__restrict__ out)
{
    const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16;
    const unsigned hmask = 0xffffu << hbase;
    if (row >= n) return;                               // whole half-warp exits together
    const float *x = X + row * d;
    float v = sqrtf(halfwarp_dot(x, x, d, hl, hmask, hbase));
    if (hl == 0) out[row] = v;
}

__global__ void normalize_kernel(const float *__restrict__ X, uint32_t B, uint32_t d, float *__restrict__ out)
{
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16;
    const unsigned hmask = 0xffffu << hbase;
    if (row >= B) return;
    const float *x = X + (size_t)row * d;
    const float nrm = sqrtf(halfwarp_dot(x, x, d, hl, hmask, hbase));
    for (uint32_t k = hl; k < d; k += 16) out[(size_t)row * d + k] = __fdiv_rn(x[k], nrm);
}

__global__ void pair_distance_kernel(const float *__restrict__ Q, const float *__restrict__ V,
                                     const uint64_t *__restrict__ pos, uint32_t B, uint32_t nc, uint32_t d,
                                     int metric, float *__restrict__ out)
{
    const uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16;
    const unsigned hmask = 0xffffu << hbase;
    if (pair >= (uint64_t)B * nc) return;
    const uint32_t q = (uint32_t)(pair / nc);
    const uint64_t ps = pos[pair];
    if (ps == UINT64_MAX) { if (hl == 0) out[pair] = CUDART_INF_F; return; }
    const float *x = Q + (size_t)q * d, *y = V + ps * d;
    float v;
    if (metric == LGPU_L2) v = halfwarp_l2(x, y, d, hl, hmask, hbase);
    else if (metric == LGPU_DOT) v = __fsub_rn(1.0f, halfwarp_dot(x, y, d, hl, hmask, hbase));
    else {   // cosine_scalar: 1 - xy / |x| / sqrt(yy)
        float xn = sqrtf(halfwarp_dot(x, x, d, hl, hmask, hbase));
        float yy = halfwarp_dot(y, y, d, hl, hmask, hbase);
        float xy = halfwarp_dot(x, y, d, hl, hmask, hbase);
        v = __fsub_rn(1.0f, __fdiv_rn(__fdiv_rn(xy, xn), sqrtf(yy)));
    }
    if (hl == 0) out[pair] = v;
}

}  // namespace

void launch_dist_matrix(const float *Q, const float *C, uint32_t B, uint64_t N, uint32_t d, int mode,
                        const float *xnorm, const float *ysqrt, float *D, uint64_t ldD, cudaStream_t st,
                        const uint32_t *only)
{
    if (B == 0 || N == 0) return;
    // the fix-up pass (`only`) is almost always a no-op: keep its CTA count small
    const uint64_t ct = (N + DM_C - 1) / DM_C;
    dim3 grid((unsigned)std::min<uint64_t>(ct, only ? 256 : ((uint64_t)1 << 30)), (B + DM_Q - 1) / DM_Q);
    dist_matrix_kernel<<<grid, DM_THREADS, 0, st>>>(Q, C, B, N, d, mode, xnorm, ysqrt, D, ldD, only); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_row_norms(const float *X, uint64_t n, uint32_t d, float *out, cudaStream_t st)
{
    if (n == 0) return;
    uint64_t threads = n * 16;
    row_norms_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, n, d, out); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_normalize(const float *X, uint32_t B, uint32_t d, float *out, cudaStream_t st)
{
    if (B == 0) return;
    uint64_t threads = (uint64_t)B * 16;
    normalize_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, B, d, out); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_pair_distance(const float *Q, const float *V, const uint64_t *pos, uint32_t B, uint32_t nc,
                          uint32_t d, int metric, float *out, cudaStream_t st)
{
    if (B == 0 || nc == 0) return;
    uint64_t threads = (uint64_t)B * nc * 16;
    pair_distance_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(Q, V, pos, B, nc, d, metric, out); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
