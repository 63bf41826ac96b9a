// fp64 tensor-core tile kernel of the implicit kernel matrix for sm_100a.
//
// tcgen05.mma has no f64 kind: the FP64 tensor path on Blackwell is the warp-level DMMA (mma.sync.m8n8k4.f64), so this
// kernel is "TMA -> 128B-swizzled shared memory ring -> LDS.128 fragments -> DMMA":
//   * one persistent CTA per SM walks this rank's share of the banded tile order (tile_order.hpp)
//   * warp 8 (one elected lane) is the TMA producer: per 16-feature slab it issues two cp.async.bulk.tensor boxes
//     (128 rows x 128 bytes of the row block and of the column block of X) into a STAGES-deep ring, completion on mbarriers
//   * warps 0-7 are DMMA consumers, 4 (M) x 2 (N), each owning a 32 x 64 sub-tile = 32 independent m8n8k4 accumulators
//   * fragments are fetched with LDS.128 (two k-steps per load): lane (g, t) reads the 16-byte chunk (4*kg + t) of row
//     perm(g); the row permutation makes every quarter-warp hit all 32 banks once under TMA's 128B swizzle, and summing
//     k in the order {0,2,4,6},{1,3,5,7} per 8-group is legal because A and B use the same order
//   * epilogue in registers: kernel function (poly / rbf via precomputed squared norms), QA_cost - q_i - q_j, the 1/C
//     diagonal, then v-weighted row sums and (training) mirrored column sums via warp shuffles; results go to the
//     per-tile partial buffer — no atomics, deterministic
// Replaces device_kernel_{linear,polynomial,rbf} (reference svm_kernel.cu:17-222) and device_kernel_predict_{polynomial,rbf}
// (predict_kernel.cu:32-74).
#pragma once

#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encode function is fetched through cudaGetDriverEntryPoint)

namespace pb {

constexpr int DMMA_BK = 16;                                  // doubles per slab = 128 bytes = one swizzle row
constexpr int DMMA_STAGES = 6;
constexpr int DMMA_STAGE_BYTES = 2 * TILE * DMMA_BK * 8;     // A box + B box = 32 KiB
constexpr int DMMA_THREADS = 288;                            // 8 consumer warps + 1 producer warp
constexpr int DMMA_VEC_BYTES = 6 * TILE * 8;                 // q_i v_i sq_i q_j v_j sq_j
constexpr int DMMA_ROWSUM_BYTES = 2 * TILE * 8;
constexpr int DMMA_COLSUM_BYTES = 4 * TILE * 8;
constexpr int DMMA_SMEM_BYTES = 1024 /* alignment slack */ + DMMA_STAGES * DMMA_STAGE_BYTES + DMMA_VEC_BYTES + DMMA_ROWSUM_BYTES + DMMA_COLSUM_BYTES + 2 * DMMA_STAGES * 8;

__device__ __forceinline__ std::uint32_t smem_u32(const void *p) { return static_cast<std::uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(const std::uint32_t bar, const std::uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const std::uint32_t bar, const std::uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(const std::uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(const std::uint32_t bar, const std::uint32_t parity) {
    std::uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a pipeline bug becomes a trap (-> CUDA error -> exception) instead of a hung GPU box
__device__ __forceinline__ void mbar_wait(const std::uint32_t bar, const std::uint32_t parity) {
    if (mbar_try_wait(bar, parity)) { return; }
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) { __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(const std::uint32_t dst, const CUtensorMap *tm, const int c0, const int c1, const std::uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(const int id, const int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// logical MMA row/column index g (0..7) -> physical row inside an 8-row swizzle atom
__device__ __forceinline__ int frag_perm(const int g) { return (g >> 1) | ((g & 1) << 2); }

template <int KERNEL, int MODE>
__global__ void __launch_bounds__(DMMA_THREADS, 1)
tile_kernel_dmma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TileParams<double> p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    double *s_vec = reinterpret_cast<double *>(smem + DMMA_STAGES * DMMA_STAGE_BYTES);  // [6][TILE]
    double *s_rowsum = s_vec + 6 * TILE;                                                // [2][TILE]
    double *s_colsum = s_rowsum + 2 * TILE;                                             // [4][TILE]
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_colsum + 4 * TILE);       // full[STAGES], empty[STAGES]
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + DMMA_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t num_slabs = p.ld / DMMA_BK;

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < DMMA_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
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
                    const std::uint32_t dstA = smem_u32(stages + stage * DMMA_STAGE_BYTES);
                    mbar_arrive_expect_tx(full0 + 8 * stage, DMMA_STAGE_BYTES);
                    tma_load_2d(dstA, &tmA, static_cast<int>(ks * DMMA_BK), static_cast<int>(I * TILE), full0 + 8 * stage);
                    tma_load_2d(dstA + DMMA_STAGE_BYTES / 2, &tmB, static_cast<int>(ks * DMMA_BK), static_cast<int>(J * TILE), full0 + 8 * stage);
                    if (++stage == DMMA_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ===== DMMA consumers: warps 0..7 =====
    const int warp_m = warp & 3, warp_n = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int pg = frag_perm(g);
    // byte offsets of this lane's 16-byte chunk for k-group 0 inside the A / B boxes (k-group 1 = offset ^ 64)
    const std::uint32_t offA = static_cast<std::uint32_t>((warp_m * 32 + pg) * 128 + ((t ^ pg) << 4));
    const std::uint32_t offB = static_cast<std::uint32_t>(DMMA_STAGE_BYTES / 2 + (warp_n * 64 + pg) * 128 + ((t ^ pg) << 4));

    std::uint32_t stage = 0, phase = 0;
    for (std::uint64_t L = p.tile_lo + blockIdx.x; L < p.tile_hi; L += gridDim.x) {
        std::uint32_t I, J;
        if constexpr (MODE == MODE_SYM) {
            tri_decode(p.T_rows, L, I, J);
        } else {
            rect_decode(p.T_rows, p.T_cols, L, I, J);
        }
        const std::uint32_t row0 = I * TILE, col0 = J * TILE;

        // per-tile epilogue vectors (read again only after the named barrier below)
        if (tid < TILE) {
            const std::uint32_t gi = row0 + tid;
            const bool ok = gi < p.n_rows;
            s_vec[0 * TILE + tid] = (MODE == MODE_SYM && ok) ? p.q[gi] : 0.0;
            s_vec[1 * TILE + tid] = (MODE == MODE_SYM && ok) ? p.v[gi] : 0.0;
            s_vec[2 * TILE + tid] = (KERNEL == K_RBF && ok) ? p.row_sq[gi] : 0.0;
        } else {
            const int c = tid - TILE;
            const std::uint32_t gj = col0 + c;
            const bool ok = gj < p.n_cols;
            s_vec[3 * TILE + c] = (MODE == MODE_SYM && ok) ? p.q[gj] : 0.0;
            s_vec[4 * TILE + c] = ok ? p.v[gj] : 0.0;
            s_vec[5 * TILE + c] = (KERNEL == K_RBF && ok) ? p.col_sq[gj] : 0.0;
        }

        double acc[4][8][2];
        #pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            #pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                acc[mi][ni][0] = 0.0;
                acc[mi][ni][1] = 0.0;
            }
        }

        for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
            mbar_wait(full0 + 8 * stage, phase);
            const unsigned char *sbase = stages + stage * DMMA_STAGE_BYTES;
            #pragma unroll
            for (int kg = 0; kg < 2; ++kg) {
                double2 a[4], b[8];
                #pragma unroll
                for (int mi = 0; mi < 4; ++mi) { a[mi] = *reinterpret_cast<const double2 *>(sbase + ((offA + mi * 1024) ^ (kg * 64))); }
                #pragma unroll
                for (int ni = 0; ni < 8; ++ni) { b[ni] = *reinterpret_cast<const double2 *>(sbase + ((offB + ni * 1024) ^ (kg * 64))); }
                #pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    #pragma unroll
                    for (int ni = 0; ni < 8; ++ni) { dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b[ni].x); }
                }
                #pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    #pragma unroll
                    for (int ni = 0; ni < 8; ++ni) { dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b[ni].y); }
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(empty0 + 8 * stage); }
            if (++stage == DMMA_STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }

        // ---- epilogue ---------------------------------------------------------------------------------------------------
        named_bar_sync(1, 256);  // s_vec of this tile visible to all consumer warps
        const double qa = (MODE == MODE_SYM) ? *p.QA_cost : 0.0;
        const bool diag = (MODE == MODE_SYM) && (I == J);
        double qi[4], vi[4], sqi[4], rowacc[4];
        #pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int rl = warp_m * 32 + mi * 8 + pg;
            qi[mi] = s_vec[0 * TILE + rl];
            vi[mi] = s_vec[1 * TILE + rl];
            sqi[mi] = s_vec[2 * TILE + rl];
            rowacc[mi] = 0.0;
        }
        #pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            #pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cl = warp_n * 64 + ni * 8 + frag_perm(2 * t + c);
                const double qj = s_vec[3 * TILE + cl], vj = s_vec[4 * TILE + cl], sqj = s_vec[5 * TILE + cl];
                double colacc = 0.0;
                #pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const double kv = kernel_from_dot<KERNEL>(acc[mi][ni][c], sqi[mi], sqj, p.kp);
                    double tt = kv;
                    if constexpr (MODE == MODE_SYM) {
                        tt = kv + qa - qi[mi] - qj;
                        if (diag && (warp_m * 32 + mi * 8 + pg) == cl) { tt += p.cost_inv; }
                        colacc += tt * vi[mi];
                    }
                    rowacc[mi] += tt * vj;
                }
                if constexpr (MODE == MODE_SYM) {
                    if (!diag) {  // CTA-uniform: mirrored contribution of an off-diagonal tile
                        colacc += __shfl_xor_sync(0xffffffffu, colacc, 4);
                        colacc += __shfl_xor_sync(0xffffffffu, colacc, 8);
                        colacc += __shfl_xor_sync(0xffffffffu, colacc, 16);
                        if (g == 0) { s_colsum[warp_m * TILE + cl] = colacc; }
                    }
                }
            }
        }
        #pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double s = rowacc[mi];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (t == 0) { s_rowsum[warp_n * TILE + warp_m * 32 + mi * 8 + pg] = s; }
        }
        named_bar_sync(1, 256);
        if (tid < TILE) {
            const double s = s_rowsum[tid] + s_rowsum[TILE + tid];
            const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
            p.partial[slot * TILE + tid] = (row0 + tid < p.n_rows) ? s : 0.0;
        } else if (MODE == MODE_SYM && !diag) {
            const int c = tid - TILE;
            const double s = ((s_colsum[c] + s_colsum[TILE + c]) + s_colsum[2 * TILE + c]) + s_colsum[3 * TILE + c];
            const std::size_t slot = static_cast<std::size_t>(J) * p.T_cols + I;
            p.partial[slot * TILE + c] = (col0 + c < p.n_cols) ? s : 0.0;
        }
    }
}

}  // namespace pb
