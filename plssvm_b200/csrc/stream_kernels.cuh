// HBM-bound streaming kernels: data packing, row statistics (q-kernel, squared norms), partial-sum reduction,
// the device-resident CG vector algebra, the w-kernel and the linear-kernel GEMV.  All reductions have a fixed order.
#pragma once

#include "common.cuh"

namespace pb {

// ---- packing ---------------------------------------------------------------------------------------------------------
// device source, row pitch d  ->  padded copy, row pitch ld (pad columns zero).  Replaces the reference's host-side
// transform_to_soa_layout + padding (layout.hpp:93-105): the B200 layout is row-major with a 128-byte-multiple pitch,
// which is what TMA boxes and coalesced row streams both want.
template <typename T>
__global__ void pack_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, const std::size_t N, const std::uint32_t d, const std::uint32_t ld) {
    const std::size_t total = N * ld;
    for (std::size_t idx = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
        const std::size_t r = idx / ld;
        const std::uint32_t c = static_cast<std::uint32_t>(idx - r * ld);
        dst[idx] = c < d ? src[r * d + c] : T(0);
    }
}

// ---- row statistics: one warp per row ------------------------------------------------------------------------------------
// sq[i] = |x_i|^2  (rbf epilogue of the tile kernels)
// `mean` (optional, ld entries, pad columns zero): norms of the centred rows x_i - mean (rbf on centred data)
template <typename T>
__global__ void __launch_bounds__(256) row_norms_kernel(const T *__restrict__ X, const std::size_t N, const std::uint32_t ld, T *__restrict__ sq, const T *__restrict__ mean) {
    const std::size_t row = static_cast<std::size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= N) { return; }
    const int lane = threadIdx.x & 31;
    const T *x = X + row * ld;
    T s0 = T(0), s1 = T(0);
    for (std::uint32_t k = 2 * lane; k < ld; k += 64) {  // ld is a multiple of 16 elements; pad columns are zero
        const T a = mean != nullptr ? x[k] - mean[k] : x[k], b = mean != nullptr ? x[k + 1] - mean[k + 1] : x[k + 1];
        s0 = pb_fma(a, a, s0);
        s1 = pb_fma(b, b, s1);
    }
    const T s = warp_sum(s0 + s1);
    if (lane == 0) { sq[row] = s; }
}

// dst = src - mean (row-wise; pad columns stay zero because mean's pad entries are zero); dst may alias src.
// Materialised centred copy for the tile kernels that read X itself (DMMA / 3xTF32 / SIMT); the int8-slice path centres on the fly in split_i8_kernel.
template <typename T>
__global__ void center_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, const std::size_t N, const std::uint32_t ld, const T *__restrict__ mean) {
    const std::size_t total = N * ld;
    for (std::size_t idx = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
        dst[idx] = src[idx] - mean[idx % ld];
    }
}

// v *= a  (feature means from the column sums)
template <typename T>
__global__ void scale_vec_kernel(T *__restrict__ v, const T a, const std::uint32_t n) {
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { v[i] *= a; }
}

// run_q_kernel (reference q_kernel.cu:16-47, CPU twin q_kernel.cpp:18-52): q[i] = k(x_i, x_last) for ALL N rows
// (entry N-1 = k(x_last, x_last) feeds QA_cost, gpu_csvm.hpp:508).  rbf uses the direct sum of squared differences
// like the reference, not the norm expansion.
template <typename T, int KERNEL>
__global__ void __launch_bounds__(256) q_kernel(const T *__restrict__ X, const std::size_t N, const std::uint32_t ld, const KernelParams<T> kp, T *__restrict__ q) {
    const std::size_t row = static_cast<std::size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= N) { return; }
    const int lane = threadIdx.x & 31;
    const T *x = X + row * ld;
    const T *xl = X + (N - 1) * ld;
    T s0 = T(0), s1 = T(0);
    for (std::uint32_t k = 2 * lane; k < ld; k += 64) {
        if constexpr (KERNEL == K_RBF) {
            const T d0 = x[k] - xl[k], d1 = x[k + 1] - xl[k + 1];
            s0 = pb_fma(d0, d0, s0);
            s1 = pb_fma(d1, d1, s1);
        } else {
            s0 = pb_fma(x[k], xl[k], s0);
            s1 = pb_fma(x[k + 1], xl[k + 1], s1);
        }
    }
    const T s = warp_sum(s0 + s1);
    if (lane == 0) {
        if constexpr (KERNEL == K_LINEAR) {
            q[row] = s;
        } else if constexpr (KERNEL == K_POLYNOMIAL) {
            q[row] = ipow(pb_fma(kp.gamma, s, kp.coef0), kp.degree);
        } else {
            q[row] = pb_exp(-kp.gamma * s);
        }
    }
}

// ---- partial-sum reduction ----------------------------------------------------------------------------------------------
// out[A * TILE + r] = (accumulate ? out : 0) + scale * sum_B partial[A][B][r] + shift, B in increasing order.
// With several ranks only the slots of tiles this rank owns were written; ownership is recomputed from the tile order.
template <typename T, int MODE>
__global__ void __launch_bounds__(512) reduce_partials_kernel(const T *__restrict__ partial, T *__restrict__ out, const std::uint32_t n_out, const std::uint32_t T_rows,
                                                              const std::uint32_t T_cols, const std::uint64_t tile_lo, const std::uint64_t tile_hi, const int check_owner,
                                                              const int tile_shift /* 1: the schedule is over 2x2 super-tiles */, const T scale, const T shift,
                                                              const int accumulate, const int *__restrict__ done) {
    if (done != nullptr && *done != 0) { return; }
    __shared__ T s_part[4][TILE];
    const int r = threadIdx.x & (TILE - 1), grp = threadIdx.x >> 7;  // 4 groups of 128 threads
    const std::uint32_t A = blockIdx.x;
    T s = T(0);
    for (std::uint32_t B = grp; B < T_cols; B += 4) {
        if (check_owner) {
            std::uint64_t L;
            const std::uint32_t Sr = (T_rows + tile_shift) >> tile_shift, Sc = (T_cols + tile_shift) >> tile_shift;
            if constexpr (MODE == MODE_SYM) {
                L = A >= B ? tri_encode(Sr, A >> tile_shift, B >> tile_shift) : tri_encode(Sr, B >> tile_shift, A >> tile_shift);
            } else {
                L = rect_encode(Sr, Sc, A >> tile_shift, B >> tile_shift);
            }
            if (L < tile_lo || L >= tile_hi) { continue; }
        }
        s += partial[(static_cast<std::size_t>(A) * T_cols + B) * TILE + r];
    }
    s_part[grp][r] = s;
    __syncthreads();
    if (grp == 0) {
        const std::uint32_t i = A * TILE + r;
        if (i < n_out) {
            const T total = ((s_part[0][r] + s_part[1][r]) + s_part[2][r]) + s_part[3][r];
            out[i] = (accumulate ? out[i] : T(0)) + scale * total + shift;
        }
    }
}

// ---- CG state ---------------------------------------------------------------------------------------------------------
// Scalars of the reference's host loop (gpu_csvm.hpp:545-627) live on the device so no iteration needs a host round trip.
template <typename T>
struct CGState {
    T delta;      // r.r
    T delta_old;
    T delta0;
    T alpha_cd;   // delta / (d.Ad)
    T beta;
    T dAd;
    T QA_cost;    // k(x_N, x_N) + 1/C
    T k_last;     // k(x_N, x_N)
    T sum_x;
    T q_dot_x;
    T bias;
    T y_last;
    unsigned long long iter;  // completed iterations
    int done;                 // 1 = converged (delta <= eps^2 delta0)
    int pad;
};

constexpr int VEC_BLOCK = 256;
constexpr int VEC_PER_THREAD = 4;
constexpr int VEC_CHUNK = VEC_BLOCK * VEC_PER_THREAD;  // elements per block of the vector kernels

// fixed-order sum of `count` block partials by one block of VEC_BLOCK threads
template <typename T>
__device__ __forceinline__ T sum_partials(const T *__restrict__ part, const std::uint32_t count, T *smem) {
    T s = T(0);
    for (std::uint32_t i = threadIdx.x; i < count; i += VEC_BLOCK) { s += part[i]; }
    return block_sum<T, VEC_BLOCK>(s, smem);
}

// b~ = y[0..n) - y[n];  x = 1   (gpu_csvm.hpp:511-515)
template <typename T>
__global__ void cg_init_kernel(const T *__restrict__ y, const std::uint32_t n, T *__restrict__ b, T *__restrict__ x, CGState<T> *st, const T cost) {
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const T y_last = y[n];
    if (i < n) {
        b[i] = y[i] - y_last;
        x[i] = T(1);
    }
    if (i == 0) {
        st->y_last = y_last;
        st->iter = 0ull;
        st->done = 0;
    }
}

// QA_cost = k(x_N, x_N) + 1/C   (gpu_csvm.hpp:508)
template <typename T>
__global__ void cg_qa_cost_kernel(const T *__restrict__ q_full, const std::uint32_t n, CGState<T> *st, const T cost) {
    st->k_last = q_full[n];
    st->QA_cost = q_full[n] + T(1) / cost;
}

// block partials of a.b
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) dot_partial_kernel(const T *__restrict__ a, const T *__restrict__ b, const std::uint32_t n, T *__restrict__ part, const int *__restrict__ done) {
    if (done != nullptr && *done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    T s = T(0);
    const std::uint32_t base = blockIdx.x * VEC_CHUNK + threadIdx.x;
    #pragma unroll
    for (int e = 0; e < VEC_PER_THREAD; ++e) {
        const std::uint32_t i = base + e * VEC_BLOCK;
        if (i < n) { s = pb_fma(a[i], b[i], s); }
    }
    s = block_sum<T, VEC_BLOCK>(s, smem);
    if (threadIdx.x == 0) { part[blockIdx.x] = s; }
}

// r = b~ - tmp (tmp = Q~ x), block partials of r.r          (gpu_csvm.hpp:533-545 and the refresh 595-609)
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) cg_residual_kernel(const T *__restrict__ b, const T *__restrict__ tmp, T *__restrict__ r, const std::uint32_t n, T *__restrict__ part, const int *__restrict__ done) {
    if (done != nullptr && *done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    T s = T(0);
    const std::uint32_t base = blockIdx.x * VEC_CHUNK + threadIdx.x;
    #pragma unroll
    for (int e = 0; e < VEC_PER_THREAD; ++e) {
        const std::uint32_t i = base + e * VEC_BLOCK;
        if (i < n) {
            const T ri = b[i] - tmp[i];
            r[i] = ri;
            s = pb_fma(ri, ri, s);
        }
    }
    s = block_sum<T, VEC_BLOCK>(s, smem);
    if (threadIdx.x == 0) { part[blockIdx.x] = s; }
}

// delta0 = delta = r.r;  d = r    (gpu_csvm.hpp:545-554)
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) cg_start_kernel(const T *__restrict__ part, const std::uint32_t nparts, CGState<T> *st, T *__restrict__ trace) {
    __shared__ T smem[VEC_BLOCK / 32];
    const T delta = sum_partials(part, nparts, smem);
    if (threadIdx.x == 0) {
        if (trace != nullptr) { trace[0] = delta; }
        st->delta = delta;
        st->delta0 = delta;
        st->delta_old = delta;
    }
}

// alpha_cd = delta / (d.Ad)   (gpu_csvm.hpp:585)
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) cg_alpha_kernel(const T *__restrict__ part, const std::uint32_t nparts, CGState<T> *st) {
    if (st->done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    const T dAd = sum_partials(part, nparts, smem);
    if (threadIdx.x == 0) {
        st->dAd = dAd;
        st->alpha_cd = st->delta / dAd;
    }
}

// x += alpha_cd d;  REFRESH ? nothing more : (r -= alpha_cd Ad, block partials of r.r)   (gpu_csvm.hpp:588, 611-613)
template <typename T, bool REFRESH>
__global__ void __launch_bounds__(VEC_BLOCK) cg_update_xr_kernel(T *__restrict__ x, T *__restrict__ r, const T *__restrict__ d, const T *__restrict__ Ad, const std::uint32_t n,
                                                                 const CGState<T> *__restrict__ st, T *__restrict__ part) {
    if (st->done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    const T alpha_cd = st->alpha_cd;
    T s = T(0);
    const std::uint32_t base = blockIdx.x * VEC_CHUNK + threadIdx.x;
    #pragma unroll
    for (int e = 0; e < VEC_PER_THREAD; ++e) {
        const std::uint32_t i = base + e * VEC_BLOCK;
        if (i < n) {
            x[i] += alpha_cd * d[i];
            if constexpr (!REFRESH) {
                const T ri = r[i] - alpha_cd * Ad[i];
                r[i] = ri;
                s = pb_fma(ri, ri, s);
            }
        }
    }
    if constexpr (!REFRESH) {
        s = block_sum<T, VEC_BLOCK>(s, smem);
        if (threadIdx.x == 0) { part[blockIdx.x] = s; }
    }
}

// delta_old = delta; delta = r.r; stop test; beta   (gpu_csvm.hpp:616-625)
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) cg_beta_kernel(const T *__restrict__ part, const std::uint32_t nparts, CGState<T> *st, const T eps, T *__restrict__ trace,
                                                            const int ignore_convergence) {
    if (st->done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    const T delta = sum_partials(part, nparts, smem);
    if (threadIdx.x == 0) {
        const T delta_old = st->delta;
        st->delta_old = delta_old;
        st->delta = delta;
        st->iter += 1ull;
        if (trace != nullptr) { trace[st->iter] = delta; }
        if (ignore_convergence == 0 && delta <= eps * eps * st->delta0) {
            st->done = 1;
        } else {
            st->beta = delta / delta_old;
        }
    }
}

// d = beta d + r   (gpu_csvm.hpp:627);  FIRST: d = r
template <typename T, bool FIRST>
__global__ void __launch_bounds__(VEC_BLOCK) cg_update_d_kernel(T *__restrict__ d, const T *__restrict__ r, const std::uint32_t n, const CGState<T> *__restrict__ st) {
    if (st->done != 0) { return; }
    const T beta = FIRST ? T(0) : st->beta;
    const std::uint32_t base = blockIdx.x * VEC_CHUNK + threadIdx.x;
    #pragma unroll
    for (int e = 0; e < VEC_PER_THREAD; ++e) {
        const std::uint32_t i = base + e * VEC_BLOCK;
        if (i < n) { d[i] = FIRST ? r[i] : beta * d[i] + r[i]; }
    }
}

// bias = y_N + QA_cost sum(x) - q.x;  alpha_N = -sum(x)   (gpu_csvm.hpp:649-653) — one block, fixed order
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) cg_finish_kernel(const T *__restrict__ x, const T *__restrict__ q, const std::uint32_t n, CGState<T> *st) {
    __shared__ T smem[VEC_BLOCK / 32];
    T sx = T(0), qx = T(0);
    for (std::uint32_t i = threadIdx.x; i < n; i += VEC_BLOCK) {
        sx += x[i];
        qx = pb_fma(q[i], x[i], qx);
    }
    sx = block_sum<T, VEC_BLOCK>(sx, smem);
    qx = block_sum<T, VEC_BLOCK>(qx, smem);
    if (threadIdx.x == 0) {
        st->sum_x = sx;
        st->q_dot_x = qx;
        st->bias = st->y_last + st->QA_cost * sx - qx;
    }
}

// y += a * x (run_device_kernel's `ret += add * Q~ v` on top of the set-style matvec)
template <typename T>
__global__ void axpy_kernel(T *__restrict__ y, const T *__restrict__ x, const T a, const std::uint32_t n) {
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { y[i] += a * x[i]; }
}

// ---- w-kernel (reference predict_kernel.cu:17-27; CPU csvm.cpp:255-280): w[f] = sum_i alpha_i SV[i][f] ----------------------
// stage 1: one thread per feature (coalesced over f), W_ROWS rows per block row-chunk; stage 2 adds the chunks in order
constexpr int W_ROWS = 256;
template <typename T>
__global__ void __launch_bounds__(256) w_partial_kernel(const T *__restrict__ SV, const T *__restrict__ alpha, const std::size_t n_sv, const std::uint32_t d, const std::uint32_t ld,
                                                        T *__restrict__ part /* [chunks][d] */) {
    const std::uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const std::size_t r0 = static_cast<std::size_t>(blockIdx.y) * W_ROWS;
    const std::size_t r1 = r0 + W_ROWS < n_sv ? r0 + W_ROWS : n_sv;
    if (f >= d) { return; }
    T s = T(0);
    for (std::size_t i = r0; i < r1; ++i) { s = pb_fma(alpha != nullptr ? alpha[i] : T(1), SV[i * ld + f], s); }  // alpha == NULL: plain column sums
    part[static_cast<std::size_t>(blockIdx.y) * d + f] = s;
}
// stage 2: 32 features x 8 chunk groups per block; every group adds its chunks in order, the 8 group sums are added in order
template <typename T>
__global__ void __launch_bounds__(256) w_reduce_kernel(const T *__restrict__ part, const std::uint32_t chunks, const std::uint32_t d, T *__restrict__ w) {
    __shared__ T s_grp[8][32];
    const int fx = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const std::uint32_t f = blockIdx.x * 32 + fx;
    T s = T(0);
    if (f < d) {
        for (std::uint32_t c = grp; c < chunks; c += 8) { s += part[static_cast<std::size_t>(c) * d + f]; }
    }
    s_grp[grp][fx] = s;
    __syncthreads();
    if (grp == 0 && f < d) {
        T total = T(0);
        #pragma unroll
        for (int g = 0; g < 8; ++g) { total += s_grp[g][fx]; }
        w[f] = total;
    }
}

// ---- factorised linear-kernel matvec (opt-in via option "linear_factorized") --------------------------------
// For k(x_i, x_j) = x_i . x_j:  (Q~ v)_i = x_i . w + (QA_cost - q_i) S - q.v + v_i / C   with  w = sum_j v_j x_j,  S = sum_j v_j
// -> two streaming passes over X (w-kernel, then this GEMV) instead of the O(n^2 d) implicit contraction.
// block partials of S = sum(v) and q.v (finished in fixed order by every block of the apply kernel)
template <typename T>
__global__ void __launch_bounds__(VEC_BLOCK) linear_fact_sums_kernel(const T *__restrict__ v, const T *__restrict__ q, const std::uint32_t n, T *__restrict__ part /* [2][blocks] */,
                                                                     const int *__restrict__ done) {
    if (done != nullptr && *done != 0) { return; }
    __shared__ T smem[VEC_BLOCK / 32];
    T s = T(0), qv = T(0);
    const std::uint32_t base = blockIdx.x * VEC_CHUNK + threadIdx.x;
    #pragma unroll
    for (int e = 0; e < VEC_PER_THREAD; ++e) {
        const std::uint32_t i = base + e * VEC_BLOCK;
        if (i < n) {
            s += v[i];
            qv = pb_fma(q[i], v[i], qv);
        }
    }
    s = block_sum<T, VEC_BLOCK>(s, smem);
    qv = block_sum<T, VEC_BLOCK>(qv, smem);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s;
        part[gridDim.x + blockIdx.x] = qv;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) linear_fact_apply_kernel(const T *__restrict__ X, const std::uint32_t n, const std::uint32_t ld, const T *__restrict__ w, const T *__restrict__ q,
                                                                const T *__restrict__ v, const T *__restrict__ part, const std::uint32_t nparts, const T *__restrict__ QA_cost,
                                                                const T cost_inv, T *__restrict__ out, const int *__restrict__ done) {
    if (done != nullptr && *done != 0) { return; }
    __shared__ T smem[256 / 32];
    __shared__ T sums[2];
    {   // S and q.v from the block partials, same fixed order in every block
        T a = T(0), b = T(0);
        for (std::uint32_t i = threadIdx.x; i < nparts; i += 256) {
            a += part[i];
            b += part[nparts + i];
        }
        a = block_sum<T, 256>(a, smem);
        b = block_sum<T, 256>(b, smem);
        if (threadIdx.x == 0) {
            sums[0] = a;
            sums[1] = b;
        }
        __syncthreads();
    }
    const std::size_t row = static_cast<std::size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n) { return; }
    const int lane = threadIdx.x & 31;
    const T *x = X + row * ld;
    T s0 = T(0), s1 = T(0);
    for (std::uint32_t k = 2 * lane; k < ld; k += 64) {
        s0 = pb_fma(x[k], w[k], s0);
        s1 = pb_fma(x[k + 1], w[k + 1], s1);
    }
    const T s = warp_sum(s0 + s1);
    if (lane == 0) { out[row] = s + (*QA_cost - q[row]) * sums[0] - sums[1] + v[row] * cost_inv; }
}

// linear-kernel prediction on the device (the reference does this GEMV on the host: gpu_csvm.hpp:702-705):
// out[p] = w . x_p - rho, one warp per point
template <typename T>
__global__ void __launch_bounds__(256) linear_predict_kernel(const T *__restrict__ P, const std::size_t m, const std::uint32_t ld, const T *__restrict__ w /* ld, zero padded */,
                                                             const T rho, T *__restrict__ out) {
    const std::size_t row = static_cast<std::size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= m) { return; }
    const int lane = threadIdx.x & 31;
    const T *x = P + row * ld;
    T s0 = T(0), s1 = T(0);
    for (std::uint32_t k = 2 * lane; k < ld; k += 64) {
        s0 = pb_fma(x[k], w[k], s0);
        s1 = pb_fma(x[k + 1], w[k + 1], s1);
    }
    const T s = warp_sum(s0 + s1);
    if (lane == 0) { out[row] = s - rho; }
}

}  // namespace pb
