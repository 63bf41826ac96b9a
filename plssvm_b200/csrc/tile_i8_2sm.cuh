// fp32 int8-slice tile kernel, CTA-pair version: tcgen05.mma.cta_group::2.kind::i8, M = 256 (two SMs x 128 rows), N = 128.
//
// Same arithmetic as tile_i8.cuh (S int8 digit planes per operand, exact int32 accumulators in TMEM, S (S + 1) / 2 products);
// what changes is operand movement.  The single-CTA kernel is bound by the shared-memory port of the SM: per
// 32-feature step its MMAs read 40 KB of operands and the producer writes 24 KB into the ring = 167 B/clk at the tensor pipe's
// pace against 128 B/clk.  A CTA pair computes a 256 x 128 block with M = 256 MMAs: each CTA stages ITS 128 rows of A and only
// ITS 64 of the 128 columns of B (the tensor cores of the pair exchange the B halves), i.e. 36 KB instead of 48 KB per slab
// and SM and 36 instead of 40 KB of operand reads per step: 141 B/clk.
//   * both CTAs: warp 0 = TMA producer — the pre-swizzled operand boxes of split_i8_kernel are contiguous, so each is fetched as
//     one 2-D box of 128-byte lines (cp.async.bulk.tensor ... cta_group::2, completing on the LEADER's mbarrier);
//     warps 2-9 = epilogue over their own 128 accumulator rows x 128 columns in their own TMEM
//   * leader CTA (cluster rank 0): warp 1 issues the MMAs for the pair; tcgen05.commit ... multicast::cluster releases the ring
//     stage / publishes the accumulators in BOTH CTAs; the epilogue warps of both CTAs hand TMEM back by arriving on the
//     leader's mbarrier (mapa + mbarrier.arrive.shared::cluster)
// Work items are the 256 x 256 super-tiles of the CTA-pair 3xTF32 kernel (same schedule, ownership and reduction, tile_shift = 1);
// a pair processes the two 128-column units of a super-tile one after the other.  In a diagonal super-tile the strictly-upper
// tile is computed but never stored.
#pragma once

#include "tile_i8.cuh"

namespace pb {

template <int S_>
struct I8PairLayout {
    static constexpr int S = S_, NH = TILE;                       // unit = 128 columns, 64 staged per CTA
    static constexpr int A_SLICE = TILE * I8_BK;                  // 8 KiB
    static constexpr int BH_SLICE = (NH / 2) * I8_BK;             // 4 KiB
    static constexpr int A_BYTES = S * A_SLICE, BH_BYTES = S * BH_SLICE;
    static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;        // 36 KiB for S = 3
    static constexpr int VEC_BYTES = (4 * TILE + 4 * NH + 4 * NH + TILE) * 4;
    static constexpr int STAGES_FIT = (227 * 1024 - 1024 - VEC_BYTES - 256) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 5 ? 5 : STAGES_FIT;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + VEC_BYTES + (2 * STAGES + 2) * 8 + 16;
    static constexpr int CPT = NH / 2;                            // columns per epilogue thread
    static_assert(S * NH <= 512 && STAGE_BYTES % 1024 == 0 && STAGES >= 2, "pair layout");
};

template <int S_, int KERNEL, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
tile_kernel_i8_2sm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TileParams<float> p) {
    using L8 = I8PairLayout<S_>;
    using T = float;
    constexpr int S = L8::S, NH = L8::NH, STAGES = L8::STAGES, CPT = L8::CPT;
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    T *s_row = reinterpret_cast<T *>(smem + STAGES * L8::STAGE_BYTES);            // [4][TILE]: q_i, v_i, sq_i, scale_i
    T *s_col = s_row + 4 * TILE;                                                   // [4][NH]: q_j, v_j, sq_j, scale_j
    T *s_colsum = s_col + 4 * NH;                                                  // [4][NH]
    T *s_rowsum = s_colsum + 4 * NH;                                               // [TILE]
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_rowsum + TILE);      // full[STAGES], empty[STAGES], tmem_full, tmem_empty
    std::uint32_t *tmem_slot = reinterpret_cast<std::uint32_t *>(bars + 2 * STAGES + 2);
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const std::uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const std::uint32_t num_slabs = p.ld8 / I8_BK;
    const std::uint32_t S_rows = (p.T_rows + 1) >> 1, S_cols = (p.T_cols + 1) >> 1;
    const std::uint64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);   // leader's only: one arrive.expect_tx, bytes of both CTAs
            mbar_init(empty0 + 8 * s, 1);  // one multicast commit per phase
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 2 * (I8_EPI_THREADS / 32));  // leader's only: 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(I8_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
    __syncthreads();
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    // work item L -> super-tile (I2, J2); this CTA's tile row I = 2 I2 + rank, the pair's two units are the tile columns 2 J2 and 2 J2 + 1
    auto decode2 = [&](const std::uint64_t L, std::uint32_t &I2, std::uint32_t &J2) {
        if constexpr (MODE == MODE_SYM) {
            tri_decode(S_rows, L, I2, J2);
        } else {
            rect_decode(S_rows, S_cols, L, I2, J2);
        }
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0;
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
                std::uint32_t I2, J2;
                decode2(L, I2, J2);
                const std::uint32_t I = 2 * I2 + rank, Il = I < p.T_rows ? I : p.T_rows - 1;  // a padding tile row re-reads the last valid block
                for (std::uint32_t c = 0; c < 2; ++c) {
                    const std::uint32_t J = 2 * J2 + c;
                    if (J >= p.T_cols) { continue; }
                    const std::size_t line_a = static_cast<std::size_t>(Il) * num_slabs * (L8::A_BYTES / 128);
                    const std::size_t line_b = (static_cast<std::size_t>(J) * 2 + rank) * num_slabs * (L8::BH_BYTES / 128);
                    for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                        const std::uint32_t dst = smem_u32(stages + stage * L8::STAGE_BYTES);
                        const std::uint32_t bar = full0 + 8 * stage;
                        if (leader) { mbar_arrive_expect_tx(bar, 2 * L8::STAGE_BYTES); }
                        const std::uint32_t leader_bar = bar & TF2_PEER_BIT_MASK;
                        tma_load_2d_2sm(dst, &tmA, 0, static_cast<int>(line_a + static_cast<std::size_t>(ks) * (L8::A_BYTES / 128)), leader_bar);
                        tma_load_2d_2sm(dst + L8::A_BYTES, &tmB, 0, static_cast<int>(line_b + static_cast<std::size_t>(ks) * (L8::BH_BYTES / 128)), leader_bar);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            std::uint32_t stage = 0, phase = 0, unit_iter = 0;
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
                std::uint32_t I2, J2;
                decode2(L, I2, J2);
                for (std::uint32_t c = 0; c < 2; ++c) {
                    if (2 * J2 + c >= p.T_cols) { continue; }
                    mbar_wait(tempty, (unit_iter & 1u) ^ 1u);  // the epilogues of both CTAs have drained the accumulators of the previous unit
                    tcgen05_fence_after();
                    for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                        mbar_wait(full0 + 8 * stage, phase);
                        tcgen05_fence_after();
                        const std::uint32_t base = smem_u32(stages + stage * L8::STAGE_BYTES);
                        const std::uint64_t d_a = umma_desc_sw64(base), d_b = umma_desc_sw64(base + L8::A_BYTES);
                        #pragma unroll
                        for (std::uint32_t k = 0; k < I8_BK / 32; ++k) {
                            const std::uint64_t koff = static_cast<std::uint64_t>((k * 32) >> 4);
                            #pragma unroll
                            for (int pp = S - 1; pp >= 0; --pp) {
                                #pragma unroll
                                for (int t = 0; t <= pp; ++t) {  // digit diagonal t': q = t' + S - 1 - p
                                    umma_i8_2sm(tmem_base + static_cast<std::uint32_t>(t * NH), d_a + koff + static_cast<std::uint64_t>((pp * L8::A_SLICE) >> 4),
                                                d_b + koff + static_cast<std::uint64_t>(((t + S - 1 - pp) * L8::BH_SLICE) >> 4), i8_idesc_pair(static_cast<std::uint32_t>(NH)),
                                                ((ks | k) == 0u && pp == S - 1) ? 0u : 1u);
                                }
                            }
                        }
                        umma_commit_2sm(empty0 + 8 * stage);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit_2sm(tfull);
                    ++unit_iter;
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): warps 2..9; warp w owns TMEM lanes 32 (w % 4) .. + 31 and columns CPT ch .. + CPT - 1 of the unit =====
        const int quarter = warp & 3;
        const int ch = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const int et = tid - 64;
        std::uint32_t unit_iter = 0;
        for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
            std::uint32_t I2, J2;
            decode2(L, I2, J2);
            const std::uint32_t I = 2 * I2 + rank, row0 = I * TILE;
            const T qa = (MODE == MODE_SYM) ? *p.QA_cost : T(0);
            for (std::uint32_t c = 0; c < 2; ++c) {
                const std::uint32_t J = 2 * J2 + c;
                if (J >= p.T_cols) { continue; }
                const bool valid = I < p.T_rows && (MODE == MODE_RECT || J <= I);
                const bool diag = (MODE == MODE_SYM) && (I == J);
                const std::uint32_t col0 = J * TILE;
                if (et < TILE) {
                    const std::uint32_t gi = row0 + et;
                    const bool oki = gi < p.n_rows;
                    s_row[0 * TILE + et] = (MODE == MODE_SYM && oki) ? p.q[gi] : T(0);
                    s_row[1 * TILE + et] = (MODE == MODE_SYM && oki) ? p.v[gi] : T(0);
                    s_row[2 * TILE + et] = (KERNEL == K_RBF && oki) ? p.row_sq[gi] : T(0);
                    s_row[3 * TILE + et] = oki ? p.A_scale[gi] : T(0);
                } else {
                    const int cidx = et - TILE;
                    const std::uint32_t gj = col0 + cidx;
                    const bool okj = gj < p.n_cols;
                    s_col[0 * NH + cidx] = (MODE == MODE_SYM && okj) ? p.q[gj] : T(0);
                    s_col[1 * NH + cidx] = okj ? p.v[gj] : T(0);
                    s_col[2 * NH + cidx] = (KERNEL == K_RBF && okj) ? p.col_sq[gj] : T(0);
                    s_col[3 * NH + cidx] = okj ? p.B_scale[gj] : T(0);
                }
                named_bar_sync(1, I8_EPI_THREADS);
                const T qi = s_row[0 * TILE + row], vi = s_row[1 * TILE + row], sqi = s_row[2 * TILE + row], sci = s_row[3 * TILE + row];

                mbar_wait(tfull, unit_iter & 1u);
                tcgen05_fence_after();
                const std::uint32_t taddr = tmem_base + (static_cast<std::uint32_t>(quarter * 32) << 16) + static_cast<std::uint32_t>(ch * CPT);

                // phase 1: S int32 diagonals -> one value per element (Horner in fp64 from the least significant diagonal)
                T a[CPT];
                #pragma unroll
                for (int g = 0; g < CPT / 8; ++g) {
                    std::uint32_t r[S][8];
                    #pragma unroll
                    for (int t = 0; t < S; ++t) { tmem_ld_32x32b_x8(taddr + static_cast<std::uint32_t>(t * NH + g * 8), r[t]); }
                    tmem_ld_wait();
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        double s = i32_to_f64(r[0][j]);
                        #pragma unroll
                        for (int t = 1; t < S; ++t) { s = fma(s, 0.00390625, i32_to_f64(r[t][j])); }
                        a[g * 8 + j] = static_cast<T>(s);
                    }
                }
                // all of this warp's accumulator reads are done: hand TMEM back to the leader's MMA warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive_cluster(tempty, 0u); }

                // phase 2: kernel function and the weighted sums
                T rowacc = T(0);
                #pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int cl = ch * CPT + j;
                    const T dot = a[j] * (sci * s_col[3 * NH + cl]);
                    const T kv = kernel_from_dot<KERNEL>(dot, sqi, s_col[2 * NH + cl], p.kp);
                    T t = kv;
                    if constexpr (MODE == MODE_SYM) {
                        t = kv + qa - qi - s_col[0 * NH + cl];
                        if (diag && row == cl) { t += p.cost_inv; }
                    }
                    rowacc = pb_fma(t, s_col[1 * NH + cl], rowacc);
                    a[j] = t * vi;  // mirrored contribution of this row to column cl
                }
                if constexpr (MODE == MODE_SYM) {
                    if (!diag) {  // CTA-uniform
                        #pragma unroll
                        for (int cc = 0; cc < CPT / 32; ++cc) {
                            #pragma unroll
                            for (int step = 16; step >= 1; step >>= 1) {
                                const bool upper = (lane & step) != 0;
                                #pragma unroll
                                for (int k = 0; k < step; ++k) {
                                    const T send = upper ? a[cc * 32 + k] : a[cc * 32 + k + step];
                                    const T keep = upper ? a[cc * 32 + k + step] : a[cc * 32 + k];
                                    a[cc * 32 + k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                                }
                            }
                            s_colsum[quarter * NH + ch * CPT + cc * 32 + lane] = a[cc * 32];
                        }
                    }
                }
                if (ch == 1) { s_rowsum[row] = rowacc; }
                named_bar_sync(1, I8_EPI_THREADS);
                if constexpr (MODE == MODE_SYM) {
                    if (!diag && valid && et < NH) {
                        const T s = ((s_colsum[et] + s_colsum[NH + et]) + s_colsum[2 * NH + et]) + s_colsum[3 * NH + et];
                        const std::size_t mslot = static_cast<std::size_t>(J) * p.T_cols + I;
                        p.partial[mslot * TILE + et] = (col0 + et < p.n_cols) ? s : T(0);
                    }
                }
                if (ch == 0 && valid) {
                    const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
                    p.partial[slot * TILE + row] = (row0 + row < p.n_rows) ? rowacc + s_rowsum[row] : T(0);
                }
                named_bar_sync(1, I8_EPI_THREADS);
                ++unit_iter;
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
