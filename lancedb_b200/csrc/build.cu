// build.cu -- index-build steps that share arithmetic with the search path (the reference
// builds IVF_PQ through lance, params in rust/lancedb/src/table/create_index.rs:283-303 and
// rust/lancedb/src/index/vector.rs:246-319; "GPU support in building vector index",
// python/python/lancedb/table.py:2883-2937).  k-means training stays in the host layer; what lives here are
// the two passes over *every* row, written so that they agree bit for bit with what the search kernels
// compute later:
//   * IVF assignment = find_partitions(row, nprobes = 1): the exact coarse kernels (dist.cu + select.cu);
//   * PQ encoding [lance, recalled: ProductQuantizer::transform]: per sub-vector, the codeword with the
//     smallest table entry (the same `subvec_l2` / `1 - dot` the distance-table build evaluates), ties to
//     the lowest code.  Residual (row - its centroid) for L2 / cosine, the row itself for dot.
#include "kernels.cuh"

namespace lgpu {

namespace {

constexpr int ENC_THREADS = 128;

// grid (ceil(n / 128), m): one thread per row, one sub-space per blockIdx.y; the sub-space's 256 codewords
// sit in shared memory ([256][DSUB] f32 <= 32 KB).
template <int DSUB>
__global__ void __launch_bounds__(ENC_THREADS) pq_encode_kernel(const float *__restrict__ X, const uint32_t *__restrict__ parts,
                                                               const float *__restrict__ centroids,
                                                               const float *__restrict__ codebook, uint64_t n,
                                                               uint32_t dim, uint32_t m, int metric,
                                                               unsigned char *__restrict__ codes)
{
    __shared__ float s_cb[256 * DSUB];
    const uint32_t i = blockIdx.y;
    const float *cb = codebook + (size_t)i * 256 * DSUB;
    for (int t = threadIdx.x; t < 256 * DSUB; t += ENC_THREADS) s_cb[t] = cb[t];
    __syncthreads();
    const uint64_t row = (uint64_t)blockIdx.x * ENC_THREADS + threadIdx.x;
    if (row >= n) return;
    float r[DSUB];
    const float *x = X + row * dim + (size_t)i * DSUB;
    if (metric == LGPU_DOT) {
#pragma unroll
        for (int e = 0; e < DSUB; e++) r[e] = x[e];
    } else {
        const float *c = centroids + (size_t)parts[row] * dim + (size_t)i * DSUB;
#pragma unroll
        for (int e = 0; e < DSUB; e++) r[e] = __fsub_rn(x[e], c[e]);
    }
    float best = 0.f;
    int best_c = 0;
    for (int c = 0; c < 256; c++) {
        float cv[DSUB];
#pragma unroll
        for (int e = 0; e < DSUB; e++) cv[e] = s_cb[c * DSUB + e];
        const float d = metric == LGPU_DOT ? subvec_dot_dist<DSUB>(r, cv) : subvec_l2<DSUB>(r, cv);
        if (c == 0 || d < best) { best = d; best_c = c; }          // strict <: lowest code wins a tie; NaN never wins
    }
    codes[row * m + i] = (unsigned char)best_c;
}

}  // namespace

void launch_pq_encode(const float *X, const uint32_t *parts, const float *centroids, const float *codebook,
                      uint64_t n, uint32_t dim, uint32_t m, int metric, unsigned char *codes, cudaStream_t st)
{
    if (n == 0) return;
    const uint32_t dsub = dim / m;
    dim3 grid((unsigned)((n + ENC_THREADS - 1) / ENC_THREADS), m);
#define LGPU_ENC(D) pq_encode_kernel<D><<<grid, ENC_THREADS, 0, st>>>(X, parts, centroids, codebook, n, dim, m, metric, codes), LGPU_COUNT_LAUNCH()
    switch (dsub) {
    case 1: LGPU_ENC(1); break;
    case 2: LGPU_ENC(2); break;
    case 4: LGPU_ENC(4); break;
    case 8: LGPU_ENC(8); break;
    case 16: LGPU_ENC(16); break;
    case 32: LGPU_ENC(32); break;
    default:
        set_error("unsupported PQ sub-vector length (dim/num_sub_vectors must be 1,2,4,8,16 or 32)");
        throw Failure{LGPU_INVALID_INPUT};
    }
#undef LGPU_ENC
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
