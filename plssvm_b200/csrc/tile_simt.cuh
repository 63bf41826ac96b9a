// SIMT (FMA-pipe) tile kernel of the implicit kernel matrix, for float and double.
//
// Role: (1) the fp32 path until the tcgen05 3xTF32 kernel takes over, (2) an independent second implementation of the
// fp64 path that the GPU tests cross-check the tensor-core kernel against, (3) selectable with option "impl" = 1.
// Replaces device_kernel_{linear,polynomial,rbf} (reference svm_kernel.cu:17-222) and device_kernel_predict_*
// (predict_kernel.cu:32-74): 128x128 output tile per CTA, 8x8 register micro-tile per thread, no atomics — row sums and
// mirrored column sums go to the per-tile partial buffer and are added in a fixed order by reduce_partials_kernel.
#pragma once

#include "common.cuh"

namespace pb {

template <typename T>
__device__ __forceinline__ void load8(const T *__restrict__ g, const bool ok, T (&r)[8]) {
    if (ok) {
        if constexpr (sizeof(T) == 8) {
            const double2 *g2 = reinterpret_cast<const double2 *>(g);
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double2 v = __ldg(g2 + e);
                r[2 * e] = v.x;
                r[2 * e + 1] = v.y;
            }
        } else {
            const float4 *g4 = reinterpret_cast<const float4 *>(g);
            #pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float4 v = __ldg(g4 + e);
                r[4 * e] = v.x;
                r[4 * e + 1] = v.y;
                r[4 * e + 2] = v.z;
                r[4 * e + 3] = v.w;
            }
        }
    } else {
        #pragma unroll
        for (int e = 0; e < 8; ++e) { r[e] = T(0); }
    }
}

template <typename T, int KERNEL, int MODE>
__global__ void __launch_bounds__(256, 1) tile_kernel_simt(const TileParams<T> p) {
    constexpr int BK = 16;
    constexpr int LDS = BK + 1;  // +1: conflict-free column reads (stride 17 words / 34 for double)
    __shared__ T sA[TILE * LDS];
    __shared__ T sB[TILE * LDS];
    __shared__ T s_rowv[3][TILE];  // q_i, v_i, |x_i|^2
    __shared__ T s_colv[3][TILE];  // q_j, v_j (RECT: alpha_j), |x_j|^2
    __shared__ T s_col[8][TILE];   // per-warp mirrored column sums

    if (p.done != nullptr && *p.done != 0) { return; }

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int warp = tid >> 5;
    const int lr = tid >> 1, lh = tid & 1;  // loader: row lr, 8 consecutive k starting at lh * 8

    for (std::uint64_t L = p.tile_lo + blockIdx.x; L < p.tile_hi; L += gridDim.x) {
        std::uint32_t I, J;
        if constexpr (MODE == MODE_SYM) {
            tri_decode(p.T_rows, L, I, J);
        } else {
            rect_decode(p.T_rows, p.T_cols, L, I, J);
        }
        const std::uint32_t row0 = I * TILE, col0 = J * TILE;

        if (tid < TILE) {
            const std::uint32_t gi = row0 + tid;
            const bool ok = gi < p.n_rows;
            s_rowv[0][tid] = (MODE == MODE_SYM && ok) ? p.q[gi] : T(0);
            s_rowv[1][tid] = (MODE == MODE_SYM && ok) ? p.v[gi] : T(0);
            s_rowv[2][tid] = (KERNEL == K_RBF && ok) ? p.row_sq[gi] : T(0);
        } else {
            const int c = tid - TILE;
            const std::uint32_t gj = col0 + c;
            const bool ok = gj < p.n_cols;
            s_colv[0][c] = (MODE == MODE_SYM && ok) ? p.q[gj] : T(0);
            s_colv[1][c] = ok ? p.v[gj] : T(0);
            s_colv[2][c] = (KERNEL == K_RBF && ok) ? p.col_sq[gj] : T(0);
        }

        T acc[8][8];
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) { acc[i][j] = T(0); }
        }

        const bool a_ok = row0 + lr < p.n_rows;
        const bool b_ok = col0 + lr < p.n_cols;
        const T *ga = p.A + static_cast<std::size_t>(row0 + lr) * p.ld + lh * 8;
        const T *gb = p.B + static_cast<std::size_t>(col0 + lr) * p.ld + lh * 8;
        T ra[8], rb[8];
        load8(ga, a_ok, ra);
        load8(gb, b_ok, rb);

        for (std::uint32_t k0 = 0; k0 < p.ld; k0 += BK) {
            __syncthreads();
            #pragma unroll
            for (int e = 0; e < 8; ++e) {
                sA[lr * LDS + lh * 8 + e] = ra[e];
                sB[lr * LDS + lh * 8 + e] = rb[e];
            }
            __syncthreads();
            if (k0 + BK < p.ld) {
                load8(ga + k0 + BK, a_ok, ra);
                load8(gb + k0 + BK, b_ok, rb);
            }
            #pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                T a[8], b[8];
                #pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = sA[(ty + 16 * i) * LDS + kk]; }
                #pragma unroll
                for (int j = 0; j < 8; ++j) { b[j] = sB[(tx + 16 * j) * LDS + kk]; }
                #pragma unroll
                for (int i = 0; i < 8; ++i) {
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[i][j] = pb_fma(a[i], b[j], acc[i][j]); }
                }
            }
        }

        // ---- epilogue: kernel function, rank-structured correction, weighted row / column sums ----------------------
        const T qa = (MODE == MODE_SYM) ? *p.QA_cost : T(0);
        const bool diag = (MODE == MODE_SYM) && (I == J);
        T rowacc[8], colacc[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) { colacc[j] = T(0); }
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = ty + 16 * i;
            const T qi = s_rowv[0][rl], vi = s_rowv[1][rl], sqi = s_rowv[2][rl];
            T racc = T(0);
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cl = tx + 16 * j;
                const T kv = kernel_from_dot<KERNEL>(acc[i][j], sqi, s_colv[2][cl], p.kp);
                T t = kv;
                if constexpr (MODE == MODE_SYM) {
                    t = kv + qa - qi - s_colv[0][cl];
                    if (diag && rl == cl) { t += p.cost_inv; }
                    colacc[j] += t * vi;
                }
                racc += t * s_colv[1][cl];
            }
            rowacc[i] = racc;
        }
        // row sums: across the 16 tx lanes of a half warp
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            T s = rowacc[i];
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (tx == 0) {
                const int rl = ty + 16 * i;
                const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
                p.partial[slot * TILE + rl] = (row0 + rl < p.n_rows) ? s : T(0);
            }
        }
        if constexpr (MODE == MODE_SYM) {
            if (!diag) {  // CTA-uniform
                #pragma unroll
                for (int j = 0; j < 8; ++j) {
                    T s = colacc[j];
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if ((ty & 1) == 0) { s_col[warp][tx + 16 * j] = s; }
                }
            }
        }
        __syncthreads();
        if constexpr (MODE == MODE_SYM) {
            if (!diag && tid < TILE) {
                T s = T(0);
                #pragma unroll
                for (int w = 0; w < 8; ++w) { s += s_col[w][tid]; }
                const std::size_t slot = static_cast<std::size_t>(J) * p.T_cols + I;
                p.partial[slot * TILE + tid] = (col0 + tid < p.n_cols) ? s : T(0);
            }
        }
    }
}

}  // namespace pb
