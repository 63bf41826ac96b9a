// int8-slice tile kernel on CTA pairs: tcgen05.mma.cta_group::2.kind::i8, M = 256 (two SMs x 128 rows), N <= 256.  The automatic choice for fp32.
//
// Same arithmetic, accumulator layout and unit epilogue (i8_epilogue_unit) as tile_i8.cuh — results are bit-identical; what changes is how the B
// operand reaches the tensor cores.  On a single CTA the MMAs of a 32-feature step read 96 KB (fp64) / 40 KB (fp32) of operands and the producer
// writes 42 / 24 KB into the ring: 154 / 171 B/clk at the tensor pipe's pace against the 128 B/clk of the shared-memory port.  With
// cta_group::2 the two tensor cores of a pair share the B operand: each CTA supplies only HALF of the N rows of every instruction out of its own
// shared memory (ncu at C3: tensor-core shared-memory wavefronts 61 % -> 57 % with the tensor pipe 74 % -> 76 % active).
//
// Two forms of the B area of a ring stage (compile-time per real type, I8PairConfig):
//   * NARROW (the fp32 default): one instruction per plane pair, N = NH; CTA r stages only ITS NH / 2 rows of every B plane — the fewest bytes
//     (fp32: 36 KB instead of 48 KB per slab and CTA, 6 stages), but S (S + 1) / 2 instructions that each re-read an A plane.
//   * WIDE: the wide-N instructions of tile_i8.cuh (one instruction multiplies A_p with several consecutive B planes that lie side by side in
//     shared memory).  One descriptor serves the pair, so the B rows of an instruction must be [first half | second half] = [CTA 0 | CTA 1] of the
//     CONCATENATION of its planes at the SAME shared-memory offset in both CTAs:
//       - region 1, plane-sized slots j = 0 .. S - HALF - 1 (HALF = half the planes of a full N = 256 instruction): CTA 0 holds plane j, CTA 1
//         plane j + HALF.  A full instruction over planes q .. q + 2 HALF - 1 points at slot q: CTA 0 supplies planes q .. q + HALF - 1, CTA 1 the rest.
//       - region 2, one piece per partial instruction (the nsl < 2 HALF planes S - nsl .. S - 1 that end a plane's range): CTA r holds the rows
//         [r nsl NH / 2, (r + 1) nsl NH / 2) of that concatenation.
//     fp64 (S = 7, NH = 64): 5 planes + 192 rows = 32 KB of B per 64-feature slab and CTA (single CTA: 28 KB), 123 B/clk through the port.
// Measured (profiles/r02/ab_pair_kernel.txt): fp32 narrow + 4.5 % at C3 (wide + 3.5 %); fp64 wide - 15 % at C2 (the fp64 epilogues of a pair run
// their exp chains in lock step and become the bottleneck, and the N = 64 / 128 / 192 / 256 mix of cta_group::2 instructions issues ~10 % slower
// than on single CTAs): the fp64 instantiation is only built with -DPLSSVM_B200_EXPERIMENTAL.
//   * both CTAs: warp 0 = TMA producer — the pre-swizzled boxes of split_i8_kernel are contiguous, so every piece is one 2-D box of 128-byte lines
//     (cp.async.bulk.tensor ... cta_group::2, completing on the LEADER's mbarrier); warps 2-9 = epilogue over their own 128 accumulator rows
//   * leader CTA (cluster rank 0): warp 1 issues the MMAs for the pair; tcgen05.commit ... multicast::cluster releases the ring stage /
//     publishes the accumulators in BOTH CTAs; the epilogue warps of both CTAs hand TMEM back by arriving on the leader's mbarrier
// Work items are 256 x 256 super-tiles (schedule, ownership and reduction as for the other CTA-pair kernels, tile_shift = 1): CTA r owns tile row
// 2 I2 + r, the pair walks the two tile columns (and the 128 / NH units of each) one after the other.  In a diagonal super-tile the strictly-upper
// tile is computed but never stored.
// Replaces device_kernel_{linear,polynomial,rbf} (reference svm_kernel.cu:17-222) and device_kernel_predict_* (predict_kernel.cu:32-74).
#pragma once

#include "tile_i8.cuh"

namespace pb {

// Compile-time configuration of the pair kernel per real type (A/B builds override them with -D...):
//   slab width BK (bytes = features per ring stage: 64 = SWIZZLE_64B rows / two K = 32 steps, 32 = SWIZZLE_32B rows / one step — twice the stages, finer refill
//   granularity: what a 2-stage ring of 88 KB needs) and WIDE (1: the wide-N layout above; 0: one instruction per plane pair, N = NH, every CTA stages
//   only its NH / 2 rows of each B plane — the least L2 -> shared-memory traffic, but S (S + 1) / 2 instructions that each re-read an A plane).
#ifndef PB_PAIR_BK_F64
    #define PB_PAIR_BK_F64 32
#endif
#ifndef PB_PAIR_BK_F32
    #define PB_PAIR_BK_F32 64
#endif
#ifndef PB_PAIR_WIDE_F64
    #define PB_PAIR_WIDE_F64 1
#endif
#ifndef PB_PAIR_WIDE_F32
    #define PB_PAIR_WIDE_F32 0
#endif
template <typename T>
struct I8PairConfig;
template <>
struct I8PairConfig<double> {
    static constexpr int BK = PB_PAIR_BK_F64;
    static constexpr bool WIDE = PB_PAIR_WIDE_F64 != 0;
};
template <>
struct I8PairConfig<float> {
    static constexpr int BK = PB_PAIR_BK_F32;
    static constexpr bool WIDE = PB_PAIR_WIDE_F32 != 0;
};

template <typename T, int S_>
struct I8PairLayout2 {
    static constexpr int S = S_, NH = I8<T>::NH, BK = I8PairConfig<T>::BK;
    static constexpr bool WIDE = I8PairConfig<T>::WIDE;
    static constexpr int UNITS = TILE / NH;
    static constexpr int SPM = WIDE ? 256 / NH : 1;              // B planes one instruction covers (wide: N = 256)
    static constexpr int HALF = SPM / 2;                         // wide: ... of which each CTA supplies this many
    static constexpr int A_SLICE = TILE * BK;                    // BK = 64: 8 KiB
    static constexpr int B_SLICE = NH * BK;                      // BK = 64: fp64 4 KiB, fp32 8 KiB
    static constexpr int HALF_SLICE = B_SLICE / 2;               // NH / 2 rows of one plane: the granule of region 2 / the slot of the narrow layout
    static constexpr int A_BYTES = S * A_SLICE;
    static constexpr int R1_SLOTS = WIDE ? S - HALF : S;         // narrow: slot q = this CTA's NH / 2 rows of plane q
    static constexpr int R1_BYTES = R1_SLOTS * (WIDE ? B_SLICE : HALF_SLICE);
    static constexpr int R2_BYTES = WIDE ? HALF_SLICE * (SPM * (SPM - 1) / 2) : 0;
    static constexpr int B_BYTES = R1_BYTES + R2_BYTES;          // wide, BK = 64: fp64 32 KiB, fp32 (S = 3) 20 KiB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // wide, BK = 64: fp64 88 KiB, fp32 44 KiB
    static constexpr int VEC_BYTES = (4 * TILE + 4 * NH + 4 * NH + TILE) * static_cast<int>(sizeof(T));
    static constexpr int STAGES_FIT = (227 * 1024 - 1024 - VEC_BYTES - 512) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 12 ? 12 : STAGES_FIT;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + VEC_BYTES + (2 * STAGES + 2) * 8 + 16;
    static constexpr int CPT = NH / 2;                           // columns per epilogue thread
    // 2-D boxes of 128-byte lines (at most 256 lines each)
    static constexpr int PLANE_LINES = A_SLICE / 128;            // lines per plane of a (row block, slab) box of the plane buffer
    static constexpr int A_LINES = A_BYTES / 128;
    static constexpr int A_BOXES = (A_LINES + 255) / 256;        // BK = 64: fp64 2 boxes of 224 lines, fp32 1 box of 192
    static constexpr int A_BOX_LINES = A_LINES / A_BOXES;
    static constexpr bool R1_ONE_BOX = WIDE && UNITS == 1;       // NH = 128: the planes of region 1 are contiguous in the source box
    static constexpr int R1_BOX_LINES = (R1_ONE_BOX ? R1_BYTES : B_SLICE) / 128;
    static constexpr int BIG_LINES = B_SLICE / 128, SMALL_LINES = HALF_SLICE / 128;
    // offset of the region-2 piece of the partial instruction over nsl planes
    __host__ __device__ static constexpr int r2_offset(const int nsl) { return R1_BYTES + HALF_SLICE * (nsl * (nsl - 1) / 2); }
    static_assert(BK == 64 || BK == 32, "slab width");
    static_assert(S * NH <= 512 && S >= SPM && (!WIDE || (SPM >= 2 && SPM % 2 == 0)), "accumulators must fit into TMEM; full instructions split evenly over the pair");
    static_assert(STAGE_BYTES % 1024 == 0 && STAGES >= 2 && HALF_SLICE % 512 == 0, "stage layout");
    static_assert(A_LINES % A_BOXES == 0 && A_BOX_LINES <= 256 && R1_BOX_LINES <= 256, "box sizes");
    static_assert(VEC_BYTES % 8 == 0 && CPT % 32 == 0, "epilogue layout");
};

template <typename T, int S_, int KERNEL, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1)
tile_kernel_i8_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmBig,
                    const __grid_constant__ CUtensorMap tmSmall, const TileParams<T> p) {
    using L8 = I8PairLayout2<T, S_>;
    constexpr int S = L8::S, NH = L8::NH, UNITS = L8::UNITS, STAGES = L8::STAGES, SPM = L8::SPM;
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
    const std::uint32_t num_slabs = p.ld8 / L8::BK;
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
    __syncthreads();     // (redundant after the cluster barrier; racecheck tracks bar.sync only)
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    // work item L -> super-tile (I2, J2); this CTA's tile row is I = 2 I2 + rank, the pair's tile columns are 2 J2 and 2 J2 + 1
    auto decode2 = [&](const std::uint64_t L, std::uint32_t &I2, std::uint32_t &J2) {
        if constexpr (MODE == MODE_SYM) {
            tri_decode(S_rows, L, I2, J2);
        } else {
            rect_decode(S_rows, S_cols, L, I2, J2);
        }
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (role_entered<PB_ELECT_PRODUCER != 0>(lane)) {
            std::uint32_t stage = 0, phase = 0;
            long long w_empty = 0;
            constexpr std::uint32_t BOX_LINES = static_cast<std::uint32_t>(L8::A_LINES);  // 128-byte lines per (row block, slab) box of the plane buffer
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
                std::uint32_t I2, J2;
                decode2(L, I2, J2);
                const std::uint32_t I = 2 * I2 + rank, Il = I < p.T_rows ? I : p.T_rows - 1;  // a padding tile row re-reads the last valid block
                for (std::uint32_t c = 0; c < 2; ++c) {
                    const std::uint32_t J = 2 * J2 + c;
                    if (J >= p.T_cols) { continue; }
                    for (int h = 0; h < UNITS; ++h) {
                        const std::uint32_t line_a0 = Il * num_slabs * BOX_LINES;
                        // first line of (plane 0, row h NH) of column block J's box
                        const std::uint32_t line_b0 = J * num_slabs * BOX_LINES + static_cast<std::uint32_t>(h * L8::BIG_LINES);
                        for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                            const long long c0 = p.stats != nullptr ? clock64() : 0;
                            mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                            if (p.stats != nullptr) { w_empty += clock64() - c0; }
                            const std::uint32_t dst = smem_u32(stages + stage * L8::STAGE_BYTES);
                            const std::uint32_t bar = full0 + 8 * stage;
                            if (leader) { mbar_arrive_expect_tx(bar, 2 * L8::STAGE_BYTES); }
                            const std::uint32_t leader_bar = bar & TF2_PEER_BIT_MASK;
                            const std::uint32_t la = line_a0 + ks * BOX_LINES, lb = line_b0 + ks * BOX_LINES;
                            #pragma unroll
                            for (int b = 0; b < L8::A_BOXES; ++b) {
                                tma_load_2d_2sm(dst + b * (L8::A_BYTES / L8::A_BOXES), &tmA, 0, static_cast<int>(la + b * L8::A_BOX_LINES), leader_bar);
                            }
                            const std::uint32_t dst_b = dst + L8::A_BYTES;
                            if constexpr (!L8::WIDE) {
                                // narrow layout: slot q <- this CTA's NH / 2 rows of plane q
                                #pragma unroll
                                for (int q = 0; q < S; ++q) {
                                    tma_load_2d_2sm(dst_b + q * L8::HALF_SLICE, &tmSmall, 0, static_cast<int>(lb + q * L8::PLANE_LINES + rank * L8::SMALL_LINES), leader_bar);
                                }
                            } else {
                                // region 1: slot j <- plane j + rank HALF (rows of unit h)
                                if constexpr (L8::R1_ONE_BOX) {
                                    tma_load_2d_2sm(dst_b, &tmR1, 0, static_cast<int>(lb + rank * L8::HALF * L8::PLANE_LINES), leader_bar);
                                } else {
                                    #pragma unroll
                                    for (int j = 0; j < L8::R1_SLOTS; ++j) {
                                        tma_load_2d_2sm(dst_b + j * L8::B_SLICE, &tmBig, 0, static_cast<int>(lb + (j + rank * L8::HALF) * L8::PLANE_LINES), leader_bar);
                                    }
                                }
                                // region 2: for nsl = 1 .. SPM - 1 the half-plane granules [rank nsl, (rank + 1) nsl) of the planes S - nsl .. S - 1
                                #pragma unroll
                                for (int nsl = 1; nsl < SPM; ++nsl) {
                                    const std::uint32_t dst_r2 = dst + L8::A_BYTES + L8::r2_offset(nsl);
                                    const int u0 = static_cast<int>(rank) * nsl;
                                    int i = 0;
                                    while (i < nsl) {
                                        const int u = u0 + i;  // granule u = (plane S - nsl + u / 2, row half u % 2)
                                        const std::uint32_t src_line = lb + static_cast<std::uint32_t>((S - nsl + (u >> 1)) * L8::PLANE_LINES + (u & 1) * L8::SMALL_LINES);
                                        if ((u & 1) == 0 && i + 1 < nsl) {  // a whole plane: one big box
                                            tma_load_2d_2sm(dst_r2 + i * L8::HALF_SLICE, &tmBig, 0, static_cast<int>(src_line), leader_bar);
                                            i += 2;
                                        } else {
                                            tma_load_2d_2sm(dst_r2 + i * L8::HALF_SLICE, &tmSmall, 0, static_cast<int>(src_line), leader_bar);
                                            i += 1;
                                        }
                                    }
                                }
                            }
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            }
            if (p.stats != nullptr) { p.stats[blockIdx.x * 8 + 3] = static_cast<unsigned long long>(w_empty); }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && role_entered<(PB_ELECT_MMA < 0 ? sizeof(T) == 4 : PB_ELECT_MMA != 0)>(lane)) {
            std::uint32_t stage = 0, phase = 0, unit_iter = 0;
            long long w_full = 0, w_tempty = 0;
            const long long c_begin = p.stats != nullptr ? clock64() : 0;
            for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
                std::uint32_t I2, J2;
                decode2(L, I2, J2);
                for (std::uint32_t c = 0; c < 2; ++c) {
                    if (2 * J2 + c >= p.T_cols) { continue; }
                    for (int h = 0; h < UNITS; ++h, ++unit_iter) {
                        const long long c0 = p.stats != nullptr ? clock64() : 0;
                        mbar_wait(tempty, (unit_iter & 1u) ^ 1u);  // the epilogues of both CTAs have drained the accumulators of the previous unit
                        if (p.stats != nullptr) { w_tempty += clock64() - c0; }
                        tcgen05_fence_after();
                        for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                            const long long c1 = p.stats != nullptr ? clock64() : 0;
                            mbar_wait(full0 + 8 * stage, phase);
                            if (p.stats != nullptr) { w_full += clock64() - c1; }
                            tcgen05_fence_after();
                            const std::uint32_t base = smem_u32(stages + stage * L8::STAGE_BYTES);
                            const std::uint64_t d_a = umma_desc_kmajor<L8::BK>(base), d_b = umma_desc_kmajor<L8::BK>(base + L8::A_BYTES);
                            #pragma unroll
                            for (std::uint32_t k = 0; k < L8::BK / 32; ++k) {
                                const std::uint64_t koff = static_cast<std::uint64_t>((k * 32) >> 4);
                                const bool first = (ks | k) == 0u;
                                // slice A_p times the slices B_q, q = S-1-p .. S-1, lands in the accumulators t' = 0 .. p (N <= 256 per instruction)
                                #pragma unroll
                                for (int pp = S - 1; pp >= 0; --pp) {
                                    const int q_lo = S - 1 - pp, cnt = pp + 1;
                                    #pragma unroll
                                    for (int cch = 0; SPM * cch < cnt; ++cch) {
                                        const int nsl = cnt - SPM * cch < SPM ? cnt - SPM * cch : SPM;
                                        const std::uint32_t d_acc = tmem_base + static_cast<std::uint32_t>(cch * SPM * NH);
                                        // wide: a full instruction reads region 1 at the slot of its first plane, a partial one its own piece of region 2;
                                        // narrow: the slot of its plane
                                        const int b_off = !L8::WIDE ? (q_lo + cch) * L8::HALF_SLICE : (nsl == SPM ? (q_lo + SPM * cch) * L8::B_SLICE : L8::r2_offset(nsl));
                                        umma_i8_2sm(d_acc, d_a + koff + static_cast<std::uint64_t>((pp * L8::A_SLICE) >> 4), d_b + koff + static_cast<std::uint64_t>(b_off >> 4),
                                                    i8_idesc_pair(static_cast<std::uint32_t>(nsl * NH)), (first && pp == S - 1) ? 0u : 1u);
                                    }
                                }
                            }
                            umma_commit_2sm(empty0 + 8 * stage);  // ring stage reusable in both CTAs once these MMAs have read it
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                        umma_commit_2sm(tfull);  // all S accumulators of this unit complete (both CTAs)
                    }
                }
            }
            if (p.stats != nullptr) {
                p.stats[blockIdx.x * 8 + 0] = static_cast<unsigned long long>(clock64() - c_begin);
                p.stats[blockIdx.x * 8 + 1] = static_cast<unsigned long long>(w_full);
                p.stats[blockIdx.x * 8 + 2] = static_cast<unsigned long long>(w_tempty);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): warps 2..9, the unit epilogue of tile_i8.cuh with the remote hand-back of TMEM =====
        std::uint32_t unit_iter = 0;
        for (std::uint64_t L = p.tile_lo + cluster_id; L < p.tile_hi; L += num_clusters) {
            std::uint32_t I2, J2;
            decode2(L, I2, J2);
            const std::uint32_t I = 2 * I2 + rank;
            const T qa = (MODE == MODE_SYM) ? *p.QA_cost : T(0);
            for (std::uint32_t c = 0; c < 2; ++c) {
                const std::uint32_t J = 2 * J2 + c;
                if (J >= p.T_cols) { continue; }
                const bool valid = I < p.T_rows && (MODE == MODE_RECT || J <= I);
                const bool diag = (MODE == MODE_SYM) && (I == J);
                T rowacc = T(0);
                for (int h = 0; h < UNITS; ++h, ++unit_iter) {
                    i8_epilogue_unit<T, S, KERNEL, MODE, true>(p, s_row, s_col, s_colsum, s_rowsum, tmem_base, tfull, tempty, unit_iter, I, J, h, valid, diag, qa, rowacc);
                }
            }
        }
    }

    // teardown: both CTAs are done with TMEM and with each other's shared memory and barriers
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
