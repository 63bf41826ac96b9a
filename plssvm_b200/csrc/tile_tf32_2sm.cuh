// fp32 tensor-core tile kernel, CTA-pair version: tcgen05.mma.cta_group::2 (3xTF32), 256 x 256 super-tiles per cluster of two SMs.
//
// Same algorithm as tile_tf32.cuh; what changes is operand movement, which is what limits the single-CTA kernel:
// a CTA pair computes a 256 x 256 tile of the implicit matrix with M = 256 / N = 256 UMMA instructions.  Each CTA stages only
// ITS 128 rows of A and ITS half (128 rows) of B per slab — 64 KB per CTA for twice the MMA work of the single-CTA kernel, so
// L2 -> shared-memory traffic per FLOP halves and the shared-memory operand reads per FLOP drop by a quarter.
//   * both CTAs: warp 0 = TMA producer (cp.async.bulk.tensor ... cta_group::2, completing on the LEADER's mbarrier),
//     warps 2-5 = epilogue over their own 128 accumulator rows x 256 columns in TMEM
//   * leader CTA (cluster rank 0): warp 1 issues the MMAs for the pair; tcgen05.commit ... multicast::cluster releases the
//     shared-memory stage / publishes the accumulator in BOTH CTAs; the epilogue warps of both CTAs hand the accumulator
//     buffer back by arriving on the leader's mbarrier (mapa + mbarrier.arrive.shared::cluster)
//   * TMEM: 512 columns per SM = two 128 x 256 fp32 accumulator buffers (epilogue of tile t overlaps the MMAs of tile t + 1)
// Super-tiles follow the same banded lower-triangle order (tile_order.hpp) on ceil(T / 2) super rows; inside a diagonal
// super-tile the strictly-upper 128-block is skipped, the diagonal blocks are direct-only.
#pragma once

#include "tile_tf32.cuh"

namespace pb {

constexpr int TF2_STAGES = 3;
constexpr int TF2_STAGE_BYTES = 4 * TF32_BOX_BYTES;  // per CTA: A_hi, A_lo, B_hi (its 128 of the 256 columns), B_lo
constexpr int TF2_THREADS = 192;
constexpr int TF2_NCOL = 2 * TILE;                   // accumulator columns per buffer
constexpr int TF2_VEC_FLOATS = 3 * TILE + 3 * TF2_NCOL;
constexpr int TF2_COLSUM_FLOATS = 4 * TF2_NCOL;
constexpr int TF2_SMEM_BYTES = 1024 + TF2_STAGES * TF2_STAGE_BYTES + (TF2_VEC_FLOATS + TF2_COLSUM_FLOATS) * 4 + (2 * TF2_STAGES + 4) * 8 + 16;
constexpr std::uint32_t TF2_TMEM_COLS = 512;
constexpr std::uint32_t TF2_PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the even (leader) CTA of the pair
// instruction descriptor: D = F32, A = B = TF32, K-major, M = 256 (pair), N = 256
constexpr std::uint32_t TF2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<std::uint32_t>(TF2_NCOL >> 3) << 17) | (static_cast<std::uint32_t>(256 >> 4) << 24);

__device__ __forceinline__ std::uint32_t cluster_ctarank() {
    std::uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const std::uint32_t dst, const CUtensorMap *tm, const int c0, const int c1, const std::uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(leader_bar)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(const std::uint32_t tmem_d, const std::uint64_t adesc, const std::uint64_t bdesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TF2_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(const std::uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(static_cast<unsigned short>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(const std::uint32_t local_bar, const std::uint32_t cta_rank) {
    asm volatile(
        "{\n\t.reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}"
        ::"r"(local_bar), "r"(cta_rank)
        : "memory");
}

template <int KERNEL, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF2_THREADS, 1)
tile_kernel_tf32_2sm(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBhi,
                     const __grid_constant__ CUtensorMap tmBlo, const TileParams<float> p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    float *s_vec = reinterpret_cast<float *>(smem + TF2_STAGES * TF2_STAGE_BYTES);  // q_i v_i sq_i [TILE] | q_j v_j sq_j [2 TILE]
    float *s_colsum = s_vec + TF2_VEC_FLOATS;                                       // [4][2 TILE]
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_colsum + TF2_COLSUM_FLOATS);
    std::uint32_t *tmem_slot = reinterpret_cast<std::uint32_t *>(bars + 2 * TF2_STAGES + 4);
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TF2_STAGES);
    const std::uint32_t tfull0 = smem_u32(bars + 2 * TF2_STAGES), tempty0 = smem_u32(bars + 2 * TF2_STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const std::uint32_t num_slabs = p.ld / TF32_BK;
    const std::uint32_t S_rows = (p.T_rows + 1) >> 1, S_cols = (p.T_cols + 1) >> 1;
    const std::uint64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < TF2_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);   // leader's only: one arrive.expect_tx, bytes of both CTAs
            mbar_init(empty0 + 8 * s, 1);  // one multicast commit per phase
        }
        #pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 8);  // leader's only: 4 epilogue warps x 2 CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TF2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
    __syncthreads();     // redundant after the cluster barrier; keeps compute-sanitizer's racecheck (which tracks bar.sync only) quiet about tmem_slot
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0;
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
                std::uint32_t I2, J2;
                if constexpr (MODE == MODE_SYM) {
                    tri_decode(S_rows, L, I2, J2);
                } else {
                    rect_decode(S_rows, S_cols, L, I2, J2);
                }
                const int ra = static_cast<int>((2 * I2 + rank) * TILE), rb = static_cast<int>((2 * J2 + rank) * TILE);
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const std::uint32_t dst = smem_u32(stages + stage * TF2_STAGE_BYTES);
                    const std::uint32_t bar = full0 + 8 * stage;
                    if (leader) { mbar_arrive_expect_tx(bar, 2 * TF2_STAGE_BYTES); }
                    const std::uint32_t leader_bar = bar & TF2_PEER_BIT_MASK;
                    const int kc = static_cast<int>(ks * TF32_BK);
                    tma_load_2d_2sm(dst + 0 * TF32_BOX_BYTES, &tmAhi, kc, ra, leader_bar);
                    tma_load_2d_2sm(dst + 1 * TF32_BOX_BYTES, &tmAlo, kc, ra, leader_bar);
                    tma_load_2d_2sm(dst + 2 * TF32_BOX_BYTES, &tmBhi, kc, rb, leader_bar);
                    tma_load_2d_2sm(dst + 3 * TF32_BOX_BYTES, &tmBlo, kc, rb, leader_bar);
                    if (++stage == TF2_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            std::uint32_t stage = 0, phase = 0, tile_iter = 0;
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters, ++tile_iter) {
                const std::uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
                tcgen05_fence_after();
                const std::uint32_t tmem_d = tmem_base + acc * TF2_NCOL;
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tcgen05_fence_after();
                    const std::uint32_t base = smem_u32(stages + stage * TF2_STAGE_BYTES);
                    const std::uint64_t d_ahi = umma_desc_sw128(base), d_alo = umma_desc_sw128(base + TF32_BOX_BYTES);
                    const std::uint64_t d_bhi = umma_desc_sw128(base + 2 * TF32_BOX_BYTES), d_blo = umma_desc_sw128(base + 3 * TF32_BOX_BYTES);
                    #pragma unroll
                    for (std::uint32_t k = 0; k < TF32_BK / 8; ++k) {
                        const std::uint64_t koff = static_cast<std::uint64_t>((k * 8 * 4) >> 4);
                        umma_tf32_2sm(tmem_d, d_alo + koff, d_bhi + koff, (ks | k) != 0u ? 1u : 0u);
                        umma_tf32_2sm(tmem_d, d_ahi + koff, d_blo + koff, 1u);
                        umma_tf32_2sm(tmem_d, d_ahi + koff, d_bhi + koff, 1u);
                    }
                    umma_commit_2sm(empty0 + 8 * stage);
                    if (++stage == TF2_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2sm(tfull0 + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): 128 accumulator rows x 256 columns each =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int et = tid - 64;
        std::uint32_t tile_iter = 0;
        for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters, ++tile_iter) {
            std::uint32_t I2, J2;
            if constexpr (MODE == MODE_SYM) {
                tri_decode(S_rows, L, I2, J2);
            } else {
                rect_decode(S_rows, S_cols, L, I2, J2);
            }
            const std::uint32_t I = 2 * I2 + rank;
            const std::uint32_t row0 = I * TILE, col0 = 2 * J2 * TILE;
            {
                const std::uint32_t gi = row0 + et;
                const bool oki = gi < p.n_rows;
                s_vec[0 * TILE + et] = (MODE == MODE_SYM && oki) ? p.q[gi] : 0.f;
                s_vec[1 * TILE + et] = (MODE == MODE_SYM && oki) ? p.v[gi] : 0.f;
                s_vec[2 * TILE + et] = (KERNEL == K_RBF && oki) ? p.row_sq[gi] : 0.f;
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const std::uint32_t gj = col0 + h * TILE + et;
                    const bool okj = gj < p.n_cols;
                    s_vec[3 * TILE + 0 * TF2_NCOL + h * TILE + et] = (MODE == MODE_SYM && okj) ? p.q[gj] : 0.f;
                    s_vec[3 * TILE + 1 * TF2_NCOL + h * TILE + et] = okj ? p.v[gj] : 0.f;
                    s_vec[3 * TILE + 2 * TF2_NCOL + h * TILE + et] = (KERNEL == K_RBF && okj) ? p.col_sq[gj] : 0.f;
                }
            }
            named_bar_sync(1, 128);
            const float *s_qj = s_vec + 3 * TILE, *s_vj = s_qj + TF2_NCOL, *s_sqj = s_vj + TF2_NCOL;
            const float qa = (MODE == MODE_SYM) ? *p.QA_cost : 0.f;
            const float qi = s_vec[0 * TILE + row], vi = s_vec[1 * TILE + row], sqi = s_vec[2 * TILE + row];

            const std::uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            mbar_wait(tfull0 + 8 * acc, acc_phase);
            tcgen05_fence_after();
            const std::uint32_t taddr = tmem_base + acc * TF2_NCOL + (static_cast<std::uint32_t>(quarter * 32) << 16);

            float rowacc[2] = { 0.f, 0.f };
            #pragma unroll 1
            for (int chunk = 0; chunk < TF2_NCOL / 32; ++chunk) {
                const int h = chunk >> 2;  // 128-column block inside the super-tile
                const std::uint32_t J = 2 * J2 + h;
                // block status (CTA-uniform): 0 = skip (strictly upper, SYM only), 1 = diagonal (direct only), 2 = lower / rect (direct [+ mirror])
                const int status = (MODE == MODE_SYM) ? (J > I ? 0 : (J == I ? 1 : 2)) : 2;
                float a[32];
                tmem_ld_32x32b_x32(taddr + chunk * 32, a);
                if (chunk == TF2_NCOL / 32 - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive_cluster(tempty0 + 8 * acc, 0u); }  // hand the buffer back to the leader's MMA warp
                }
                if (status == 0) { continue; }
                float racc = 0.f;
                #pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int cl = chunk * 32 + j;
                    const float kv = kernel_from_dot<KERNEL>(a[j], sqi, s_sqj[cl], p.kp);
                    float t = kv;
                    if constexpr (MODE == MODE_SYM) {
                        t = kv + qa - qi - s_qj[cl];
                        if (status == 1 && row == (cl & (TILE - 1))) { t += p.cost_inv; }
                    }
                    racc = fmaf(t, s_vj[cl], racc);
                    a[j] = t * vi;
                }
                rowacc[h] += racc;
                if constexpr (MODE == MODE_SYM) {
                    if (status == 2) {
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
                        s_colsum[quarter * TF2_NCOL + chunk * 32 + lane] = a[0];
                    }
                }
            }
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const std::uint32_t J = 2 * J2 + h;
                const bool direct = (MODE == MODE_SYM) ? (J <= I) : true;
                if (direct && I < p.T_rows && J < p.T_cols) {
                    const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
                    p.partial[slot * TILE + row] = (row0 + row < p.n_rows) ? rowacc[h] : 0.f;
                }
            }
            named_bar_sync(1, 128);
            if constexpr (MODE == MODE_SYM) {
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const std::uint32_t J = 2 * J2 + h;
                    if (J < I && I < p.T_rows) {  // mirrored contribution of a strictly-lower block: output block J, source block I
                        const int c = h * TILE + et;
                        const float s = ((s_colsum[c] + s_colsum[TF2_NCOL + c]) + s_colsum[2 * TF2_NCOL + c]) + s_colsum[3 * TF2_NCOL + c];
                        const std::size_t mslot = static_cast<std::size_t>(J) * p.T_cols + I;
                        p.partial[mslot * TILE + et] = (col0 + c < p.n_cols) ? s : 0.f;
                    }
                }
            }
            named_bar_sync(1, 128);
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TF2_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
