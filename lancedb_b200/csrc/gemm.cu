// gemm.cu -- tcgen05 / TMEM / TMA bf16 GEMM used as a *shortlist generator*.
//
// Where the path is a true dense Q x C contraction -- the IVF coarse step
// (IvfModel::find_partitions) at large nlist and the flat
// KNNVectorDistance (row a11, BASELINE config 4) -- the reference spends
// B*N*d fused multiply-adds of f32 SIMD.  Here the bulk of that work runs on the 5th-gen
// tensor cores in bf16 with f32 accumulation in TMEM:
//     S[q][x] = |x|^2 - 2 * sum_k bf16(q_k) * bf16(x_k)          (= |q-x|^2 - |q|^2, approx.)
// and is only used to pick candidates: the caller keeps every column whose S is within a
// rigorous error band of the k-th best and re-scores those exactly in f32 in lance's
// rounding order (dist.cu), so the final ids / distances are still bit-identical to the
// oracle while >99.9% of the arithmetic is tensor-core work.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads, cta_group::1):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of a 128x64 Q tile and a 256x64 X tile
//               (bf16, K-major, SWIZZLE_128B) into a 4-stage shared-memory ring, mbarrier
//               complete_tx signalling;
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues
//               tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) x4 per stage,
//               tcgen05.commit releases the stage / publishes the accumulator;
//   warps 2-5   epilogue: tcgen05.ld 32x32b.x32 of their TMEM lane quadrant, |x|^2 - 2*acc,
//               128-byte row stores; two 256-column accumulators double-buffer MMA against
//               the epilogue.
// Tiles are walked N-tile-major so the eight 128-query tiles of a batch reuse an X tile
// from L2.
#include "kernels.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

namespace lgpu {

namespace {

constexpr int GM = 128, GN = 256, GK = 64, GSTAGES = 4, G_THREADS = 192;
constexpr uint32_t A_STAGE_BYTES = GM * GK * 2;     // 16 KB
constexpr uint32_t B_STAGE_BYTES = GN * GK * 2;     // 32 KB
constexpr uint32_t G_SMEM_TILES = GSTAGES * (A_STAGE_BYTES + B_STAGE_BYTES);   // 192 KB
constexpr uint32_t G_SMEM_BYTES = G_SMEM_TILES + 256 + 1024;                   // + barriers + alignment slack

// ---- raw PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 | LBO(=1, unused for swizzled K-major)<<16 | SBO(=1024 B between 8-row groups)>>4 <<32 |
// version 1 <<46 | layout SWIZZLE_128B(2) <<61
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major A and B,
// N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t G_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);

// LIST: the filtering epilogue for DENSE hit rates (coarse step at many lists), see the epilogue; a separate
// instantiation so that the dense / sparse-filter kernel of the flat path and of the small coarse problems is
// unchanged (staging |x|^2 through shared memory and the per-tile barrier cost the flat C4 launch 15 %).
template <bool LIST>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_dist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 const float *__restrict__ xnorm2, float *__restrict__ out, uint64_t ld_out, uint32_t B, uint64_t N,
                 uint32_t num_kb, GemmFilter flt)
{
    pdl_entry();                                       // PDL: let the next grid in, wait for the previous one
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(16) float s_xn[LIST ? 2 : 1][LIST ? GN : 4];  // LIST: |x|^2 of the current tile's columns, per accumulator
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;         // SWIZZLE_128B needs 1024 B alignment
    const uint32_t smem_a = base, smem_b = base + GSTAGES * A_STAGE_BYTES;
    const uint32_t bars = base + G_SMEM_TILES;
    const uint32_t full_bar = bars, empty_bar = bars + 8 * GSTAGES;
    const uint32_t tfull_bar = bars + 16 * GSTAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    unsigned char *smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + G_SMEM_TILES + 16 * GSTAGES + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t num_m = (B + GM - 1) / GM;
    const uint64_t num_n = (N + GN - 1) / GN;
    const uint64_t num_tiles = num_m * num_n;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < GSTAGES; i++) { mbar_init(full_bar + 8 * i, 1); mbar_init(empty_bar + 8 * i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull_bar + 8 * i, 1); mbar_init(tempty_bar + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (uint64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const uint32_t m_tile = (uint32_t)(t % num_m);
                const uint64_t n_tile = t / num_m;
                for (uint32_t kb = 0; kb < num_kb; kb++) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    mbar_expect_tx(full_bar + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
                    tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_q, full_bar + 8 * stage, (int)(kb * GK), (int)(m_tile * GM));
                    tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_x, full_bar + 8 * stage, (int)(kb * GK), (int)(n_tile * GN));
                    if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t stage = 0, phase = 0, it = 0;
        for (uint64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, it++) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);              // epilogue drained this accumulator
            tc_fence_after();
            for (uint32_t kb = 0; kb < num_kb; kb++) {
                mbar_wait(full_bar + 8 * stage, phase);                  // TMA bytes landed
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_kmajor_sw128_desc(smem_a + stage * A_STAGE_BYTES);
                    const uint64_t db = make_kmajor_sw128_desc(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < GK / 16; k++)                    // +32 B (>>4 = 2) per K=16 slice
                        tc_mma_bf16(tmem_base + acc * GN, da + 2 * k, db + 2 * k, G_IDESC, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(empty_bar + 8 * stage);                    // frees the smem stage when the MMAs retire
                    if (kb + 1 == num_kb) tc_commit(tfull_bar + 8 * acc);   // accumulator complete
                }
                __syncwarp();
                if (++stage == GSTAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====
        const int quad = warp & 3;
        uint32_t it = 0;
        for (uint64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, it++) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const uint32_t m_tile = (uint32_t)(t % num_m);
            const uint64_t n_tile = t / num_m;
            const uint64_t x0 = n_tile * GN;
            if constexpr (LIST) {
                // |x|^2 of the tile's 256 columns -> shared memory while the MMAs of the tile still run (read from global
                // memory chunk by chunk, each chunk exposed an L2 round trip behind its TMEM load).  Columns past N read
                // as 0 and are masked where it matters.  The barrier also keeps a warp from overwriting the buffer of tile
                // i + 2 while another still reads tile i's.
                const int et = (int)threadIdx.x - 64;                    // 0..127 among the epilogue threads
                s_xn[acc][et] = x0 + et < N ? __ldg(xnorm2 + x0 + et) : 0.f;
                s_xn[acc][et + 128] = x0 + et + 128 < N ? __ldg(xnorm2 + x0 + et + 128) : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t q = m_tile * GM + quad * 32 + lane;
            float *orow = out + (size_t)q * ld_out;
#define LGPU_TMEM_LD32(r, taddr)                                                                                        \
    asm volatile(                                                                                                       \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                       \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                       \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),               \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),         \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),       \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])        \
        : "r"(taddr));                                                                                                  \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
            [[maybe_unused]] const float *xn_s = s_xn[LIST ? acc : 0];
            if constexpr (LIST) {
                // Filtering epilogue for DENSE hit rates (the coarse step's lists: ~1.5 % of the columns pass, so nearly
                // every 32-column chunk holds one).  One returning atomic per hit, as in the sparse form below, made this
                // launch 3.5x slower than writing the dense matrix.  Two passes over the accumulator, which stays in
                // TMEM until it is released: (1) hit masks of the tile's 8 chunks + their count, ONE atomicAdd per
                // (query, tile) reserves the slots; (2) the chunks that hold hits are loaded again and their (column,
                // score) pairs stored.  tcgen05.ld is warp-collective, so pass 2 re-loads a chunk when ANY lane has a hit.
                const bool live = q < B;
                const float thr = live ? flt.thr[q] : 0.f;
                uint32_t tot = 0, slot = 0;
                // ROLLED loops, one chunk body run 2 x 8 times: fully unrolled (masks kept in registers between the
                // passes) the kernel grew to 17 k instructions and stalled on instruction fetch (ncu: no_inst) -- 3x
                // slower than writing the dense matrix.  Pass 1 recomputes nothing it can keep: only the count.
#pragma unroll 1
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 1) {
                        if (!__any_sync(0xffffffffu, tot != 0u)) break;
                        if (tot) slot = atomicAdd(flt.count + q, tot);
                    }
#pragma unroll 1
                    for (int c = 0; c < GN / 32; c++) {
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + acc * GN + c * 32 + ((uint32_t)(quad * 32) << 16);
                        LGPU_TMEM_LD32(r, taddr);
                        const uint64_t xb = x0 + (uint64_t)c * 32;
                        unsigned hit = 0;
                        if (live && xb + 32 <= N) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 xn = *reinterpret_cast<const float4 *>(xn_s + c * 32 + j);
                                hit |= (xn.x - 2.0f * __uint_as_float(r[j]) <= thr ? 1u : 0u) << j;
                                hit |= (xn.y - 2.0f * __uint_as_float(r[j + 1]) <= thr ? 1u : 0u) << (j + 1);
                                hit |= (xn.z - 2.0f * __uint_as_float(r[j + 2]) <= thr ? 1u : 0u) << (j + 2);
                                hit |= (xn.w - 2.0f * __uint_as_float(r[j + 3]) <= thr ? 1u : 0u) << (j + 3);
                            }
                        } else if (live) {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (xb + j < N && xn_s[c * 32 + j] - 2.0f * __uint_as_float(r[j]) <= thr) hit |= 1u << j;
                        }
                        if (pass == 0) {
                            tot += __popc(hit);
                        } else if (hit) {
#pragma unroll
                            for (int j = 0; j < 32; j++) {
                                if ((hit >> j) & 1u) {
                                    if (slot < flt.cap) {
                                        flt.cand_pos[(size_t)q * flt.cap + slot] = xb + j;
                                        flt.cand_s[(size_t)q * flt.cap + slot] = xn_s[c * 32 + j] - 2.0f * __uint_as_float(r[j]);
                                    }
                                    slot++;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar + 8 * acc);
                continue;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < GN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + acc * GN + c0 + ((uint32_t)(quad * 32) << 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (q < B && flt.thr) {
                    // filtering epilogue: nothing dense is written; scores not above the per-query threshold
                    // are appended to the query's candidate list (rare: ~1e-4 of the columns)
                    const uint64_t xb = x0 + c0;
                    const float thr = flt.thr[q];
                    auto admit = [&](uint64_t x) {
                        const uint32_t slot = atomicAdd(flt.count + q, 1u);
                        if (slot < flt.cap) {
                            flt.cand_pos[(size_t)q * flt.cap + slot] = x;
                            if (flt.cand_ids) flt.cand_ids[(size_t)q * flt.cap + slot] = flt.col_ids ? flt.col_ids[x] : x;
                        }
                    };
                    if (xb + 32 <= N) {
                        unsigned hit = 0;                                  // bit j: column xb+j passes
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 xn = __ldg(reinterpret_cast<const float4 *>(xnorm2 + xb + j));
                            hit |= (xn.x - 2.0f * __uint_as_float(r[j]) <= thr ? 1u : 0u) << j;
                            hit |= (xn.y - 2.0f * __uint_as_float(r[j + 1]) <= thr ? 1u : 0u) << (j + 1);
                            hit |= (xn.z - 2.0f * __uint_as_float(r[j + 2]) <= thr ? 1u : 0u) << (j + 2);
                            hit |= (xn.w - 2.0f * __uint_as_float(r[j + 3]) <= thr ? 1u : 0u) << (j + 3);
                        }
                        while (hit) { const int j = __ffs(hit) - 1; hit &= hit - 1; admit(xb + j); }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint64_t x = xb + j;
                            if (x < N && __ldg(xnorm2 + x) - 2.0f * __uint_as_float(r[j]) <= thr) admit(x);
                        }
                    }
                } else if (q < B) {
                    const uint64_t xb = x0 + c0;
                    if (xb + 32 <= N && ((ld_out & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 xn = __ldg(reinterpret_cast<const float4 *>(xnorm2 + xb + j));
                            float4 o;
                            o.x = xn.x - 2.0f * __uint_as_float(r[j]);
                            o.y = xn.y - 2.0f * __uint_as_float(r[j + 1]);
                            o.z = xn.z - 2.0f * __uint_as_float(r[j + 2]);
                            o.w = xn.w - 2.0f * __uint_as_float(r[j + 3]);
                            *reinterpret_cast<float4 *>(orow + xb + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (xb + j < N) orow[xb + j] = __ldg(xnorm2 + xb + j) - 2.0f * __uint_as_float(r[j]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar + 8 * acc);                           // 128 arrivals free the accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// X f32 [n][d] -> bf16 [n][d] (round to nearest even) and |x|^2 in f32; one warp per row
__global__ void to_bf16_norm_kernel(const float *__restrict__ X, uint64_t n, uint32_t d, __nv_bfloat16 *__restrict__ Xb,
                                    float *__restrict__ norm2)
{
    pdl_entry();                                       // PDL: let the next grid in, wait for the previous one
    const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float *x = X + row * d;
    __nv_bfloat16 *xb = Xb + row * d;
    float s = 0.f;
    for (uint32_t k = lane; k < d; k += 32) {
        float v = x[k];
        xb[k] = __float2bfloat16_rn(v);
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && norm2) norm2[row] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LGPU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
            throw Failure{LGPU_RUNTIME};
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 row-major [rows][d] matrix, box = box_rows x 64 elements, 128-byte swizzle, OOB = zeros
CUtensorMap make_map(const void *ptr, uint64_t rows, uint32_t d, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {d, rows};
    cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)GK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        throw Failure{LGPU_RUNTIME};
    }
    return m;
}

__global__ void band_check_kernel(const float *__restrict__ approx, const uint32_t *__restrict__ cnt,
                                  const float *__restrict__ qnorm2, float xmax, uint32_t d, uint32_t B, uint32_t k,
                                  uint32_t kp, uint32_t *__restrict__ flags)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    uint32_t f = 0;
    if (cnt[q] >= kp && kp > 0) {              // list is full: columns beyond it may exist
        const float qn = sqrtf(qnorm2[q]);
        const float s = qn + xmax;
        const float E = 0.0078125f * 1.00390625f * qn * xmax + 4.0f * (float)d * 5.9604645e-8f * s * s;
        const float kth = approx[(size_t)q * kp + (k - 1 < kp ? k - 1 : kp - 1)];
        const float last = approx[(size_t)q * kp + kp - 1];
        f = (k >= kp || !(last > kth + 2.0f * E)) ? 1u : 0u;
    }
    flags[q] = f;
}

// thr[q] = (k-th smallest approximate score of the sample) + 2 E_q, or +inf when the sample holds < k rows;
// every true top-k row of the full set has a score <= thr[q] (E_q as in band_check_kernel)
__global__ void sample_threshold_kernel(const float *__restrict__ approx, const uint32_t *__restrict__ cnt,
                                        const float *__restrict__ qnorm2, float xmax, uint32_t d, uint32_t B, uint32_t k,
                                        float *__restrict__ thr)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B) return;
    float t = __int_as_float(0x7f800000);
    if (cnt[q] >= k) {
        const float qn = sqrtf(qnorm2[q]);
        const float s = qn + xmax;
        const float E = 0.0078125f * 1.00390625f * qn * xmax + 4.0f * (float)d * 5.9604645e-8f * s * s;
        t = approx[(size_t)q * k + k - 1] + 2.0f * E;
    }
    thr[q] = t;
}
// The same threshold straight from the dense sample scores D[B][ld] (ns columns), one warp per query: the k-th smallest
// by counting bisection over the lane's register copy of the row (no ids, no sorted output -- a top-k select over
// 8192 x 2048 scores took 0.21 ms), stopped when the bracket is a quarter of the band it feeds; any hi with
// count(S <= hi) >= k is a valid bound.  NaN scores never count.  ns <= 32 * VPL.
template <int VPL>
__global__ void __launch_bounds__(128) sample_kth_threshold_kernel(const float *__restrict__ D, uint64_t ld, uint32_t ns,
                                                                  const float *__restrict__ qnorm2, float xmax, uint32_t d,
                                                                  uint32_t B, uint32_t k, float *__restrict__ thr)
{
    pdl_entry();                                       // PDL: let the next grid in, wait for the previous one
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= B) return;
    const float *row = D + (size_t)q * ld;
    // global -> shared with cp.async (all of a lane's requests in flight), then shared -> registers: plain register
    // loads come out of ptxas as load -> use -> load, one DRAM round trip after the other (see coarse_finish_kernel)
    __shared__ __align__(16) float s_rows[4][VPL * 32];
    float *srow = s_rows[threadIdx.x >> 5];
    {
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(srow);
        for (uint32_t i = (uint32_t)lane * 4; i < ld && i < (uint32_t)VPL * 32; i += 128)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + i * 4), "l"(row + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    float v[VPL];
#pragma unroll
    for (int j = 0; j < VPL; j++) v[j] = srow[min((uint32_t)j * 32 + lane, ns - 1)];
    float lo = __int_as_float(0x7f800000), hi = -lo;
    uint32_t nv = 0;
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        if ((uint32_t)j * 32 + lane >= ns) v[j] = __int_as_float(0x7fc00000);
        if (v[j] == v[j]) { lo = fminf(lo, v[j]); hi = fmaxf(hi, v[j]); nv++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    float t = __int_as_float(0x7f800000);
    if (nv >= k && hi < __int_as_float(0x7f800000) && lo > -__int_as_float(0x7f800000)) {
        const float qn = sqrtf(qnorm2[q]);
        const float s = qn + xmax;
        const float E = 0.0078125f * 1.00390625f * qn * xmax + 4.0f * (float)d * 5.9604645e-8f * s * s;
        for (int it = 0; it < 24 && hi - lo > 0.25f * E; it++) {        // invariant: count(S <= hi) >= k
            const float mid = 0.5f * lo + 0.5f * hi;
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < VPL; j++) c += v[j] <= mid ? 1u : 0u;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= k) hi = mid; else lo = mid;
        }
        t = hi + 2.0f * E;
    }
    if (lane == 0) thr[q] = t;
}

__global__ void overflow_flags_kernel(const uint32_t *__restrict__ count, uint32_t cap, uint32_t B, uint32_t *__restrict__ flags)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < B) flags[q] = count[q] > cap ? 1u : 0u;
}

}  // namespace

void launch_sample_threshold(const float *approx, const uint32_t *cnt, const float *qnorm2, float xmax, uint32_t d,
                             uint32_t B, uint32_t k, float *thr, cudaStream_t st)
{
    if (B == 0) return;
    sample_threshold_kernel<<<(B + 127) / 128, 128, 0, st>>>(approx, cnt, qnorm2, xmax, d, B, k, thr); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

bool launch_sample_kth_threshold(const float *D, uint64_t ld, uint32_t ns, const float *qnorm2, float xmax, uint32_t d,
                                 uint32_t B, uint32_t k, float *thr, cudaStream_t st)
{
    if (B == 0) return true;
    if (ns == 0 || ns > 2048 || (ld & 3) || ld < ns) return false;       // caller falls back to select + threshold
    const unsigned grid = (B + 3) / 4;
    if (ns <= 1024) launch_k(sample_kth_threshold_kernel<32>, dim3(grid), dim3(128), 0, st, D, ld, ns, qnorm2, xmax, d, B, k, thr);
    else launch_k(sample_kth_threshold_kernel<64>, dim3(grid), dim3(128), 0, st, D, ld, ns, qnorm2, xmax, d, B, k, thr);
    LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
    return true;
}

void launch_overflow_flags(const uint32_t *count, uint32_t cap, uint32_t B, uint32_t *flags, cudaStream_t st)
{
    if (B == 0) return;
    overflow_flags_kernel<<<(B + 127) / 128, 128, 0, st>>>(count, cap, B, flags); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_band_check(const float *approx, const uint32_t *cnt, const float *qnorm2, float xmax, uint32_t d,
                       uint32_t B, uint32_t k, uint32_t kp, uint32_t *flags, cudaStream_t st)
{
    if (B == 0) return;
    band_check_kernel<<<(B + 127) / 128, 128, 0, st>>>(approx, cnt, qnorm2, xmax, d, B, k, kp, flags); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

bool gemm_shape_supported(uint32_t d) { return d >= 8 && d % 8 == 0; }

void launch_to_bf16(const float *X, uint64_t n, uint32_t d, void *Xb, float *norm2, cudaStream_t st)
{
    if (n == 0) return;
    uint64_t threads = n * 32;
    launch_k(to_bf16_norm_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, X, n, d, reinterpret_cast<__nv_bfloat16 *>(Xb), norm2); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

// The filtering epilogue applied to an already written dense score matrix D[B][ld] (same scores, same
// `<= thr` test, same candidate lists up to order): used when the threshold sample was the whole matrix, so
// a second tensor-core pass would only recompute what D already holds.
__global__ void filter_dense_kernel(const float *__restrict__ D, uint64_t ld, uint64_t N, GemmFilter flt)
{
    const uint32_t q = blockIdx.y;
    const float thr = flt.thr[q];
    const float *row = D + (size_t)q * ld;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < N; x += (uint64_t)gridDim.x * blockDim.x) {
        if (row[x] <= thr) {
            const uint32_t slot = atomicAdd(flt.count + q, 1u);
            if (slot < flt.cap) {
                flt.cand_pos[(size_t)q * flt.cap + slot] = x;
                flt.cand_ids[(size_t)q * flt.cap + slot] = flt.col_ids ? flt.col_ids[x] : x;
            }
        }
    }
}

void launch_filter_dense(const float *D, uint64_t ld, uint32_t B, uint64_t N, const GemmFilter &flt, cudaStream_t st)
{
    if (B == 0 || N == 0) return;
    dim3 grid((unsigned)std::min<uint64_t>((N + 1023) / 1024, 64), B);
    filter_dense_kernel<<<grid, 256, 0, st>>>(D, ld, N, flt); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

void launch_gemm_dist(const void *Qb, const void *Xb, const float *xnorm2, uint32_t B, uint64_t N, uint32_t d,
                      float *out, uint64_t ld_out, int num_sms, cudaStream_t st, const GemmFilter *filter)
{
    if (B == 0 || N == 0) return;
    LGPU_REQUIRE(gemm_shape_supported(d), "tensor-core path needs a dimension that is a multiple of 8");
    CUtensorMap mq = make_map(Qb, B, d, GM);
    CUtensorMap mx = make_map(Xb, N, d, GN);
    const uint64_t tiles = (uint64_t)((B + GM - 1) / GM) * ((N + GN - 1) / GN);
    const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)num_sms);
    GemmFilter flt{};
    if (filter) flt = *filter;
    const bool list = flt.thr && flt.cand_s;                  // dense hit rates: two-pass list epilogue
    auto kern = list ? gemm_dist_kernel<true> : gemm_dist_kernel<false>;
    LGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES));
    launch_k(kern, dim3(grid), dim3(G_THREADS), G_SMEM_BYTES, st, mq, mx, xnorm2, out, ld_out, B, N, (uint32_t)((d + GK - 1) / GK), flt); LGPU_COUNT_LAUNCH();
    LGPU_CUDA(cudaGetLastError());
}

}  // namespace lgpu
