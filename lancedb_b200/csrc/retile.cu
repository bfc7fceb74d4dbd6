// retile.cu -- open-time re-layout of the index into the HBM layout the scan kernel
// streams (the analogue of `prewarm_index` loading an index into the Session cache,
// rust/lancedb/src/table.rs:3283-3286).  Pure byte shuffling; exact.
//
// PQ codes.  lance keeps a partition's codes transposed [m][n_p] [lance, recalled].
// Here each partition becomes (nch+1) blocks of 8 bytes per row, laid out
// [block][row][8]: a warp reading 32 consecutive rows of one block reads 256 contiguous
// bytes.  Row r's byte stream is its m codes (padded with zeros to 8*nch) delayed by
// r % 8 positions: stream position M holds code[r][M - r%8].  That skew is what makes
// the scan kernel's shared-memory gathers conflict-free (scan2.cu, scan3.cu).
//
// Codebook.  [m][256][dsub] becomes [nch][256 c][8 s][dsub] so the 8 sub-spaces of a
// chunk for one code are contiguous (a warp building 4 codes x 8 sub-spaces reads 1 KB).
#include "kernels.cuh"

namespace lgpu {

namespace {

__global__ void retile_codes_kernel(const unsigned char *__restrict__ codes, int layout,
                                    const uint64_t *__restrict__ part_off, uint32_t nlist, uint64_t nrows,
                                    uint32_t m, uint32_t nch, const uint64_t *__restrict__ code_base,
                                    const uint32_t *__restrict__ part_npad, unsigned char *__restrict__ out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nblk = nch + 1;
    if (idx >= nrows * nblk) return;
    const uint64_t grow = idx / nblk;              // global storage row
    const uint32_t blk = (uint32_t)(idx - grow * nblk);
    uint32_t lo = 0, hi = nlist - 1;               // partition of grow: part_off[p] <= grow < part_off[p+1]
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (part_off[mid + 1] > grow) hi = mid; else lo = mid + 1;
    }
    const uint32_t p = lo;
    const uint64_t pbase = part_off[p];
    const uint32_t row = (uint32_t)(grow - pbase);
    const uint32_t n_p = (uint32_t)(part_off[p + 1] - pbase);
    const uint32_t sig = row & 7;
    uint64_t word = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        int i = (int)(blk * 8 + e) - (int)sig;
        if (i >= 0 && i < (int)m) {
            unsigned char c = layout == LGPU_CODES_ROW_MAJOR
                                  ? codes[grow * m + i]
                                  : codes[pbase * m + (uint64_t)i * n_p + row];
            word |= (uint64_t)c << (8 * e);
        }
    }
    *reinterpret_cast<uint64_t *>(out + code_base[p] + ((uint64_t)blk * part_npad[p] + row) * 8) = word;
}

__global__ void retile_codebook_kernel(const float *__restrict__ cb, uint32_t m, uint32_t dsub, uint32_t nch,
                                       float *__restrict__ out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)nch * 256 * 8 * dsub;
    if (idx >= total) return;
    uint32_t e = idx % dsub;
    uint64_t t = idx / dsub;
    uint32_t s = t % 8; t /= 8;
    uint32_t c = t % 256;
    uint32_t ch = (uint32_t)(t / 256);
    uint32_t i = ch * 8 + s;
    out[idx] = i < m ? cb[((size_t)i * 256 + c) * dsub + e] : 0.f;
}

}  // namespace

void launch_retile_codes(const unsigned char *codes, int layout, const uint64_t *part_off, uint32_t nlist,
                         uint64_t nrows, uint32_t m, uint32_t nch, const uint64_t *code_base,
                         const uint32_t *part_npad, unsigned char *out, cudaStream_t st)
{
    if (nrows == 0) return;
    uint64_t total = nrows * (nch + 1);
    retile_codes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes, layout, part_off, nlist, nrows, m,
                                                                       nch, code_base, part_npad, out); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_retile_codebook(const float *codebook, uint32_t m, uint32_t dsub, uint32_t nch, float *out,
                            cudaStream_t st)
{
    uint64_t total = (uint64_t)nch * 256 * 8 * dsub;
    retile_codebook_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codebook, m, dsub, nch, out); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
