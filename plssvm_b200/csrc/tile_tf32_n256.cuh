// fp32 tensor-core tile kernel, wide-tile version: tcgen05.mma.cta_group::1 (3xTF32) with M = 128, N = 256.
//
// Measured on B200 the 128 x 128 kernel (tile_tf32.cuh) and the CTA-pair kernel (tile_tf32_2sm.cuh) run at the same speed:
// both move ~13.3 KB through the shared-memory port per 128x128x8 MMA (TMA fill of hi AND lo operands + operand reads), i.e.
// ~104 clk at 128 B/clk against 110 clk measured — the 3xTF32 scheme is bound by shared-memory bandwidth, not by the tensor
// pipe.  A 128 x 256 tile per CTA cuts that traffic by a quarter: per 32-feature slab A (hi, lo) 32 KB + B (hi, lo) 64 KB feed
// 12 MMAs of 128 x 256 x 8 (twice the work of the 128 x 128 kernel's 64 KB).
// Scheduling unit: half of a 256 x 256 super-tile of the banded order (row block 2 I2 + r, column blocks 2 J2 and 2 J2 + 1), so
// tile ranges / rank ownership are identical to the CTA-pair kernel.  2-stage ring (2 x 96 KB), TMEM 2 x 256 columns.
#pragma once

#include "tile_tf32_2sm.cuh"

namespace pb {

constexpr int TN_STAGES = 2;
constexpr int TN_STAGE_BYTES = 2 * TF32_BOX_BYTES + 2 * 2 * TF32_BOX_BYTES;  // A_hi, A_lo (128 rows) + B_hi, B_lo (256 rows)
constexpr int TN_THREADS = 192;
constexpr int TN_SMEM_BYTES = 1024 + TN_STAGES * TN_STAGE_BYTES + (TF2_VEC_FLOATS + TF2_COLSUM_FLOATS) * 4 + (2 * TN_STAGES + 4) * 8 + 16;
// instruction descriptor: D = F32, A = B = TF32, K-major, M = 128, N = 256
constexpr std::uint32_t TN_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<std::uint32_t>(TF2_NCOL >> 3) << 17) | (static_cast<std::uint32_t>(TILE >> 4) << 24);

__device__ __forceinline__ void umma_tf32_n256(const std::uint32_t tmem_d, const std::uint64_t adesc, const std::uint64_t bdesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TN_IDESC), "r"(accumulate)
        : "memory");
}

template <int KERNEL, int MODE>
__global__ void __launch_bounds__(TN_THREADS, 1)
tile_kernel_tf32_n256(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBhi,
                      const __grid_constant__ CUtensorMap tmBlo, const TileParams<float> p) {
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    float *s_vec = reinterpret_cast<float *>(smem + TN_STAGES * TN_STAGE_BYTES);
    float *s_colsum = s_vec + TF2_VEC_FLOATS;
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_colsum + TF2_COLSUM_FLOATS);
    std::uint32_t *tmem_slot = reinterpret_cast<std::uint32_t *>(bars + 2 * TN_STAGES + 4);
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TN_STAGES);
    const std::uint32_t tfull0 = smem_u32(bars + 2 * TN_STAGES), tempty0 = smem_u32(bars + 2 * TN_STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t num_slabs = p.ld / TF32_BK;
    const std::uint32_t S_rows = (p.T_rows + 1) >> 1, S_cols = (p.T_cols + 1) >> 1;
    const std::uint64_t unit_lo = 2 * p.tile_lo, unit_hi = 2 * p.tile_hi;  // unit = (super-tile, half r)

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < TN_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        #pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TF2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    const auto decode = [&](const std::uint64_t U, std::uint32_t &I, std::uint32_t &J2) {
        std::uint32_t I2;
        if constexpr (MODE == MODE_SYM) {
            tri_decode(S_rows, U >> 1, I2, J2);
        } else {
            rect_decode(S_rows, S_cols, U >> 1, I2, J2);
        }
        I = 2 * I2 + static_cast<std::uint32_t>(U & 1u);
    };

    if (warp == 0) {
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0;
            for (std::uint64_t U = unit_lo + blockIdx.x; U < unit_hi; U += gridDim.x) {
                std::uint32_t I, J2;
                decode(U, I, J2);
                const int ra = static_cast<int>(I * TILE), rb = static_cast<int>(2 * J2 * TILE);
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                    const std::uint32_t dst = smem_u32(stages + stage * TN_STAGE_BYTES);
                    const std::uint32_t bar = full0 + 8 * stage;
                    mbar_arrive_expect_tx(bar, TN_STAGE_BYTES);
                    const int kc = static_cast<int>(ks * TF32_BK);
                    tma_load_2d(dst + 0 * TF32_BOX_BYTES, &tmAhi, kc, ra, bar);
                    tma_load_2d(dst + 1 * TF32_BOX_BYTES, &tmAlo, kc, ra, bar);
                    tma_load_2d(dst + 2 * TF32_BOX_BYTES, &tmBhi, kc, rb, bar);               // B_hi rows 0..127
                    tma_load_2d(dst + 3 * TF32_BOX_BYTES, &tmBhi, kc, rb + TILE, bar);        // B_hi rows 128..255 (contiguous: one 256-row K-major operand)
                    tma_load_2d(dst + 4 * TF32_BOX_BYTES, &tmBlo, kc, rb, bar);
                    tma_load_2d(dst + 5 * TF32_BOX_BYTES, &tmBlo, kc, rb + TILE, bar);
                    if (++stage == TN_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            std::uint32_t stage = 0, phase = 0, tile_iter = 0;
            for (std::uint64_t U = unit_lo + blockIdx.x; U < unit_hi; U += gridDim.x, ++tile_iter) {
                const std::uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
                tcgen05_fence_after();
                const std::uint32_t tmem_d = tmem_base + acc * TF2_NCOL;
                for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tcgen05_fence_after();
                    const std::uint32_t base = smem_u32(stages + stage * TN_STAGE_BYTES);
                    const std::uint64_t d_ahi = umma_desc_sw128(base), d_alo = umma_desc_sw128(base + TF32_BOX_BYTES);
                    const std::uint64_t d_bhi = umma_desc_sw128(base + 2 * TF32_BOX_BYTES), d_blo = umma_desc_sw128(base + 4 * TF32_BOX_BYTES);
                    #pragma unroll
                    for (std::uint32_t k = 0; k < TF32_BK / 8; ++k) {
                        const std::uint64_t koff = static_cast<std::uint64_t>((k * 8 * 4) >> 4);
                        umma_tf32_n256(tmem_d, d_alo + koff, d_bhi + koff, (ks | k) != 0u ? 1u : 0u);
                        umma_tf32_n256(tmem_d, d_ahi + koff, d_blo + koff, 1u);
                        umma_tf32_n256(tmem_d, d_ahi + koff, d_bhi + koff, 1u);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (++stage == TN_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull0 + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int et = tid - 64;
        std::uint32_t tile_iter = 0;
        for (std::uint64_t U = unit_lo + blockIdx.x; U < unit_hi; U += gridDim.x, ++tile_iter) {
            std::uint32_t I, J2;
            decode(U, I, J2);
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
                const int h = chunk >> 2;
                const std::uint32_t J = 2 * J2 + h;
                const int status = (MODE == MODE_SYM) ? (J > I ? 0 : (J == I ? 1 : 2)) : 2;
                float a[32];
                tmem_ld_32x32b_x32(taddr + chunk * 32, a);
                if (chunk == TF2_NCOL / 32 - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(tempty0 + 8 * acc); }
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
                    if (J < I && I < p.T_rows) {
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
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TF2_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
