// Shared device-side definitions: kernel-function epilogues, tile parameters, small PTX helpers.
#pragma once

#include "tile_order.hpp"

#include <cuda_runtime.h>

#include <cstdint>

namespace pb {

enum : int { K_LINEAR = 0,
             K_POLYNOMIAL = 1,
             K_RBF = 2 };

enum : int { MODE_SYM = 0,    // training matvec: lower-triangle tiles, mirrored  (run_svm_kernel)
             MODE_RECT = 1 }; // prediction: points x support vectors               (run_predict_kernel)

template <typename T>
struct KernelParams {
    int kernel;
    int degree;
    T gamma;
    T coef0;
};

// Everything one launch of a tile kernel needs.  Rows index the A operand (SYM: X; RECT: the points),
// columns the B operand (SYM: X again; RECT: the support vectors).
template <typename T>
struct TileParams {
    const T *A;            // row-major, pitch `ld` elements, rows >= n_rows are never read un-masked
    const T *B;
    const T *A_hi, *A_lo;  // fp32 tensor path only: TF32 hi / lo split of A and B (tile_tf32.cuh)
    const T *B_hi, *B_lo;
    // int8-slice tensor path only (tile_i8.cuh): digit planes of A and B in the boxed, pre-swizzled layout of split_i8_kernel (boxes of 128
    // rows; the B operand of a 128 x NH unit is a row range of such a box), and the per-row scales 2^(e - 6).  B_i8b: a second copy of B's
    // planes in boxes of 64 rows, read only by the experimental CTA-pair kernel (tile_i8_2sm.cuh)
    const std::int8_t *A_i8, *B_i8, *B_i8b;
    const T *A_scale, *B_scale;
    std::uint32_t ld8;               // features padded to a multiple of 64 (= 64 x number of slabs)
    std::uint32_t n_rows;  // valid rows of A  (SYM: n = N - 1)
    std::uint32_t n_cols;  // valid rows of B
    std::uint32_t ld;      // row pitch in elements (multiple of 128 bytes, zero padded)
    std::uint32_t T_rows;  // ceil(n_rows / TILE)
    std::uint32_t T_cols;  // ceil(n_cols / TILE)
    std::uint64_t tile_lo; // this rank's share of the banded tile order
    std::uint64_t tile_hi;
    const T *row_sq;       // squared norms of A rows (rbf)
    const T *col_sq;       // squared norms of B rows (rbf)
    const T *q;            // SYM: q vector (n)
    const T *v;            // SYM: direction vector (n);  RECT: alpha (n_cols)
    const T *QA_cost;      // SYM: device scalar k(x_N, x_N) + 1/C
    T cost_inv;            // SYM: 1 / C added on the diagonal
    KernelParams<T> kp;
    T *partial;            // [T_rows][T_cols][TILE] (SYM: [T][T][TILE], slot (A, B) = contribution of block B to output block A)
    const int *done;       // CG convergence flag: all kernels of a speculatively enqueued iteration exit when set
    unsigned long long *stats;  // optional (option "tile_stats"): per CTA 8 words of cycle counters of the int8-slice kernel's roles (tile_i8.cuh)
    int slow_drain;        // debugging / A-B measurements: 1 = the fp32 int8-slice epilogue converts before it releases TMEM (option "fp32_fast_drain" = 0)
};

// integer power by repeated squaring; matches std::pow(x, (real) degree) of kernel_function_types.hpp:86-89 to a few ulp
// and pow(real, int) of the reference's CUDA kernels (svm_kernel.cu:142)
template <typename T>
__device__ __forceinline__ T ipow(const T x, const int degree) {
    unsigned int e = degree < 0 ? static_cast<unsigned int>(-static_cast<long long>(degree)) : static_cast<unsigned int>(degree);
    T base = x;
    T r = T(1);
    while (e != 0u) {
        if (e & 1u) { r *= base; }
        base *= base;
        e >>= 1u;
    }
    return degree < 0 ? T(1) / r : r;
}

__device__ __forceinline__ float pb_exp(const float x) { return expf(x); }
__device__ __forceinline__ double pb_exp(const double x) { return exp(x); }
__device__ __forceinline__ float pb_fma(const float a, const float b, const float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double pb_fma(const double a, const double b, const double c) { return fma(a, b, c); }

// the same with the degree known at compile time (DEG = 0: the runtime value); identical operation order, hence identical rounding:
// the loop above yields x, x x, x (x x), (x x) (x x) for degree 1 .. 4
template <int DEG, typename T>
__device__ __forceinline__ T ipow_ct(const T x, const int degree) {
    if constexpr (DEG == 1) {
        return x;
    } else if constexpr (DEG == 2) {
        return x * x;
    } else if constexpr (DEG == 3) {
        return x * (x * x);
    } else if constexpr (DEG == 4) {
        const T sq = x * x;
        return sq * sq;
    } else {
        return ipow(x, degree);
    }
}

// kernel function from the contraction result: `dot` = x_i . x_j, sq_* = squared norms (rbf only)
// (kernel_function_types.hpp:75-97; rbf through |x_i|^2 + |x_j|^2 - 2 x_i.x_j instead of the reference's direct
//  sum of squared differences, clamped at 0 — see DESIGN.md "numerics")
template <int KERNEL, typename T, int DEG = 0>
__device__ __forceinline__ T kernel_from_dot(const T dot, const T sq_i, const T sq_j, const KernelParams<T> &kp) {
    if constexpr (KERNEL == K_LINEAR) {
        return dot;
    } else if constexpr (KERNEL == K_POLYNOMIAL) {
        return ipow_ct<DEG>(pb_fma(kp.gamma, dot, kp.coef0), kp.degree);
    } else {
        T d2 = pb_fma(T(-2), dot, sq_i + sq_j);
        d2 = d2 < T(0) ? T(0) : d2;  // (a NaN distance stays NaN)
        return pb_exp(-kp.gamma * d2);
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

// fixed-order block sum (deterministic for a given block size): warp shuffles, then warp 0 adds the warp results in order
template <typename T, int BLOCK>
__device__ __forceinline__ T block_sum(T v, T *smem /* BLOCK / 32 entries */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { smem[warp] = v; }
    __syncthreads();
    T r = T(0);
    #pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) { r += smem[w]; }
    return r;
}

}  // namespace pb
