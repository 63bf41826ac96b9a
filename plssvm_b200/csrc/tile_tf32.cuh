// fp32 tensor-core tile kernel of the implicit kernel matrix for sm_100a: tcgen05.mma kind::tf32 with the 3xTF32 split.
//
// One TF32 product loses 13 mantissa bits of each operand, which breaks the 1e-4 parity tolerance after a CG solve, so every
// operand is split once per data set into hi = tf32(x) and lo = tf32(x - hi) (split_tf32_kernel) and the contraction is
//     x_i . x_j  ~=  lo_i . hi_j  +  hi_i . lo_j  +  hi_i . hi_j        (fp32 accumulation in TMEM; lo.lo ~ 2^-22 dropped)
// Structure (the canonical Blackwell pipeline: TMA -> smem ring -> tcgen05.mma -> TMEM -> tcgen05.ld epilogue):
//   * warp 0: TMA producer — per 32-feature slab four 128x128-byte boxes (A_hi, A_lo, B_hi, B_lo), SWIZZLE_128B
//   * warp 1: allocates TMEM (2 x 128 columns: double-buffered 128x128 fp32 accumulators) and issues the MMAs: one elected
//     lane, 12 tcgen05.mma.cta_group::1.kind::tf32 (M = N = 128, K = 8) per slab, tcgen05.commit frees the smem stage and,
//     after the last slab, publishes the accumulator to the epilogue
//   * warps 2-5: epilogue — tcgen05.ld 32x32b.x32 (one accumulator row per thread), kernel function + QA_cost - q_i - q_j
//     (+ 1/C on the diagonal), v-weighted row sums in registers, mirrored column sums by a 31-shuffle butterfly per 32 columns;
//     the accumulator buffer is released right after the loads, so the epilogue of tile t overlaps the MMAs of tile t + 1
// Replaces device_kernel_{linear,polynomial,rbf}<float> (reference svm_kernel.cu:17-222) and device_kernel_predict_*<float>.
#pragma once

#include "common.cuh"
#include "tile_dmma.cuh"  // mbarrier / TMA helpers

namespace pb {

constexpr int TF32_BK = 32;                               // floats per slab = 128 bytes = one swizzle row
constexpr int TF32_STAGES = 3;
constexpr int TF32_BOX_BYTES = TILE * TF32_BK * 4;        // 16 KiB
constexpr int TF32_STAGE_BYTES = 4 * TF32_BOX_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int TF32_THREADS = 192;                         // producer warp, MMA warp, 4 epilogue warps
constexpr int TF32_VEC_BYTES = 6 * TILE * 4;
constexpr int TF32_COLSUM_BYTES = 4 * TILE * 4;
constexpr int TF32_SMEM_BYTES = 1024 + TF32_STAGES * TF32_STAGE_BYTES + TF32_VEC_BYTES + TF32_COLSUM_BYTES + (2 * TF32_STAGES + 4) * 8 + 16;
constexpr std::uint32_t TF32_TMEM_COLS = 256;             // 2 accumulator buffers x 128 fp32 columns

// hi / lo split (round to nearest TF32; the tensor core truncates, so pre-rounded operands are exact for it)
__global__ void split_tf32_kernel(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, const std::size_t count) {
    for (std::size_t i = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
        const float v = x[i];
        std::uint32_t h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hf));
        hi[i] = hf;
        lo[i] = __uint_as_float(l);
    }
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ std::uint64_t umma_desc_sw128(const std::uint32_t smem_addr) {
    return static_cast<std::uint64_t>((smem_addr & 0x3FFFFu) >> 4)  // start address, bits [0, 14)
           | (static_cast<std::uint64_t>(1) << 16)                  // leading byte offset (unused for swizzled K-major), bits [16, 30)
           | (static_cast<std::uint64_t>(1024 >> 4) << 32)          // stride byte offset: 8 rows x 128 B, bits [32, 46)
           | (static_cast<std::uint64_t>(1) << 46)                  // descriptor version (Blackwell)
           | (static_cast<std::uint64_t>(2) << 61);                 // layout type SWIZZLE_128B
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr std::uint32_t TF32_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<std::uint32_t>(TILE >> 3) << 17) | (static_cast<std::uint32_t>(TILE >> 4) << 24);

__device__ __forceinline__ void umma_tf32(const std::uint32_t tmem_d, const std::uint64_t adesc, const std::uint64_t bdesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TF32_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(const std::uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 accumulator columns of this thread's row (TMEM lane) -> registers
__device__ __forceinline__ void tmem_ld_32x32b_x32(const std::uint32_t taddr, float (&v)[32]) {
    std::uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r[i]); }
}

template <int KERNEL, int MODE>
__global__ void __launch_bounds__(TF32_THREADS, 1)
tile_kernel_tf32(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBhi,
                 const __grid_constant__ CUtensorMap tmBlo, const TileParams<float> p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    float *s_vec = reinterpret_cast<float *>(smem + TF32_STAGES * TF32_STAGE_BYTES);  // [6][TILE]
    float *s_colsum = s_vec + 6 * TILE;                                               // [4][TILE]
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_colsum + 4 * TILE);     // full[S], empty[S], tmem_full[2], tmem_empty[2]
    std::uint32_t *tmem_slot = reinterpret_cast<std::uint32_t *>(bars + 2 * TF32_STAGES + 4);
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TF32_STAGES);
    const std::uint32_t tfull0 = smem_u32(bars + 2 * TF32_STAGES), tempty0 = smem_u32(bars + 2 * TF32_STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t num_slabs = p.ld / TF32_BK;

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < TF32_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        #pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TF32_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0;
            for (std::uint64_t L = p.tile_lo + blockIdx.x; L < p.tile_hi; L += gridDim.x) {
                std::uint32_t I, J;
                if constexpr (MODE == MODE_SYM) {
                    tri_decode(p.T_rows, L, I, J);
                } else {
                    rect_decode(p.T_rows, p.T_cols, L, I, J);
                }
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const std::uint32_t dst = smem_u32(stages + stage * TF32_STAGE_BYTES);
                    const std::uint32_t bar = full0 + 8 * stage;
                    mbar_arrive_expect_tx(bar, TF32_STAGE_BYTES);
                    const int kc = static_cast<int>(ks * TF32_BK), ra = static_cast<int>(I * TILE), rb = static_cast<int>(J * TILE);
                    tma_load_2d(dst + 0 * TF32_BOX_BYTES, &tmAhi, kc, ra, bar);
                    tma_load_2d(dst + 1 * TF32_BOX_BYTES, &tmAlo, kc, ra, bar);
                    tma_load_2d(dst + 2 * TF32_BOX_BYTES, &tmBhi, kc, rb, bar);
                    tma_load_2d(dst + 3 * TF32_BOX_BYTES, &tmBlo, kc, rb, bar);
                    if (++stage == TF32_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0, tile_iter = 0;
            for (std::uint64_t L = p.tile_lo + blockIdx.x; L < p.tile_hi; L += gridDim.x, ++tile_iter) {
                const std::uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);  // epilogue has drained this accumulator buffer
                tcgen05_fence_after();
                const std::uint32_t tmem_d = tmem_base + acc * TILE;
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tcgen05_fence_after();
                    const std::uint32_t base = smem_u32(stages + stage * TF32_STAGE_BYTES);
                    const std::uint64_t d_ahi = umma_desc_sw128(base), d_alo = umma_desc_sw128(base + TF32_BOX_BYTES);
                    const std::uint64_t d_bhi = umma_desc_sw128(base + 2 * TF32_BOX_BYTES), d_blo = umma_desc_sw128(base + 3 * TF32_BOX_BYTES);
                    #pragma unroll
                    for (std::uint32_t k = 0; k < TF32_BK / 8; ++k) {
                        const std::uint64_t koff = static_cast<std::uint64_t>((k * 8 * 4) >> 4);  // advance the start address by 32 bytes per K = 8
                        umma_tf32(tmem_d, d_alo + koff, d_bhi + koff, (ks | k) != 0u ? 1u : 0u);
                        umma_tf32(tmem_d, d_ahi + koff, d_blo + koff, 1u);
                        umma_tf32(tmem_d, d_ahi + koff, d_bhi + koff, 1u);
                    }
                    umma_commit(empty0 + 8 * stage);  // smem stage reusable once these MMAs have read it
                    if (++stage == TF32_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull0 + 8 * acc);  // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32 * (warp % 4) .. + 31 =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;  // accumulator row of this thread
        const int et = tid - 64;              // 0..127 among the epilogue threads
        std::uint32_t tile_iter = 0;
        for (std::uint64_t L = p.tile_lo + blockIdx.x; L < p.tile_hi; L += gridDim.x, ++tile_iter) {
            std::uint32_t I, J;
            if constexpr (MODE == MODE_SYM) {
                tri_decode(p.T_rows, L, I, J);
            } else {
                rect_decode(p.T_rows, p.T_cols, L, I, J);
            }
            const std::uint32_t row0 = I * TILE, col0 = J * TILE;
            {
                const std::uint32_t gi = row0 + et, gj = col0 + et;
                const bool oki = gi < p.n_rows, okj = gj < p.n_cols;
                s_vec[0 * TILE + et] = (MODE == MODE_SYM && oki) ? p.q[gi] : 0.f;
                s_vec[1 * TILE + et] = (MODE == MODE_SYM && oki) ? p.v[gi] : 0.f;
                s_vec[2 * TILE + et] = (KERNEL == K_RBF && oki) ? p.row_sq[gi] : 0.f;
                s_vec[3 * TILE + et] = (MODE == MODE_SYM && okj) ? p.q[gj] : 0.f;
                s_vec[4 * TILE + et] = okj ? p.v[gj] : 0.f;
                s_vec[5 * TILE + et] = (KERNEL == K_RBF && okj) ? p.col_sq[gj] : 0.f;
            }
            named_bar_sync(1, 128);
            const float qa = (MODE == MODE_SYM) ? *p.QA_cost : 0.f;
            const bool diag = (MODE == MODE_SYM) && (I == J);
            const float qi = s_vec[0 * TILE + row], vi = s_vec[1 * TILE + row], sqi = s_vec[2 * TILE + row];

            const std::uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tcgen05_fence_after();
            const std::uint32_t taddr = tmem_base + acc * TILE + (static_cast<std::uint32_t>(quarter * 32) << 16);

            float rowacc = 0.f;
            #pragma unroll 1
            for (int chunk = 0; chunk < TILE / 32; ++chunk) {
                float a[32];
                tmem_ld_32x32b_x32(taddr + chunk * 32, a);
                if (chunk == TILE / 32 - 1) {
                    // all of this warp's accumulator reads are done: hand the buffer back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(tempty0 + 8 * acc); }
                }
                #pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int cl = chunk * 32 + j;
                    const float kv = kernel_from_dot<KERNEL>(a[j], sqi, s_vec[5 * TILE + cl], p.kp);
                    float t = kv;
                    if constexpr (MODE == MODE_SYM) {
                        t = kv + qa - qi - s_vec[3 * TILE + cl];
                        if (diag && row == cl) { t += p.cost_inv; }
                    }
                    rowacc = fmaf(t, s_vec[4 * TILE + cl], rowacc);
                    a[j] = t * vi;  // mirrored contribution of this row to column cl
                }
                if constexpr (MODE == MODE_SYM) {
                    if (!diag) {  // CTA-uniform
                        // butterfly: after 5 halving steps lane c holds the sum over the warp's 32 rows of column chunk * 32 + c
                        #pragma unroll
                        for (int step = 16; step >= 1; step >>= 1) {
                            const bool upper = (lane & step) != 0;
                            #pragma unroll
                            for (int k = 0; k < step; ++k) {
                                const float send = upper ? a[k] : a[k + step];
                                const float keep = upper ? a[k + step] : a[k];
                                a[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                            }
                        }
                        s_colsum[quarter * TILE + chunk * 32 + lane] = a[0];
                    }
                }
            }
            const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
            p.partial[slot * TILE + row] = (row0 + row < p.n_rows) ? rowacc : 0.f;
            named_bar_sync(1, 128);  // column sums of all four warps visible; s_vec free for the next tile afterwards
            if constexpr (MODE == MODE_SYM) {
                if (!diag) {
                    const float s = ((s_colsum[et] + s_colsum[TILE + et]) + s_colsum[2 * TILE + et]) + s_colsum[3 * TILE + et];
                    const std::size_t mslot = static_cast<std::size_t>(J) * p.T_cols + I;
                    p.partial[mslot * TILE + et] = (col0 + et < p.n_cols) ? s : 0.f;
                }
            }
            named_bar_sync(1, 128);  // s_colsum consumed before the next tile overwrites it
        }
    }

    // teardown: everyone done with TMEM before the allocating warp frees it
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TF32_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
