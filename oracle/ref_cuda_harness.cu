// TEST / BASELINE INFRASTRUCTURE ONLY — runs the reference's OWN CUDA kernels (src/plssvm/backends/CUDA/svm_kernel.cu, compiled
// unchanged for sm_100 in place from /root/reference by oracle/Makefile) on the B200 as a second oracle and a GPU-vs-GPU baseline.
// This file only reproduces the reference's launch convention: feature-major SoA layout with 96 zero padding rows per feature
// (layout.hpp:93-105, gpu_csvm.hpp:302-346), vectors padded to n + 96, grid ceil(n / 96)^2 x block 16 x 16 (gpu_csvm.hpp:431-447),
// argument order of csvm.cu:134-153.
#include "plssvm/backends/CUDA/svm_kernel.cuh"
#include "plssvm/constants.hpp"

#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

namespace {

#define RC(call)                                                                   \
    do {                                                                           \
        const cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                                  \
            std::fprintf(stderr, "ref_cuda_harness: %s failed: %s\n", #call, cudaGetErrorString(e__)); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

template <typename T>
__global__ void to_soa_kernel(const T *__restrict__ rowmajor, T *__restrict__ soa, const std::size_t n, const std::size_t d, const std::size_t rows_padded) {
    for (std::size_t idx = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x; idx < n * d; idx += static_cast<std::size_t>(gridDim.x) * blockDim.x) {
        const std::size_t i = idx / d, f = idx - i * d;
        soa[f * rows_padded + i] = rowmajor[idx];
    }
}

template <typename T>
int run(const int kernel, const T *X_dev_rowmajor, const std::size_t N, const std::size_t d, const T *q, const T *v, const T QA_cost, const T cost_inv, const T add,
        const int degree, const T gamma, const T coef0, T *ret_inout, float *ms_out, const int reps) {
    using plssvm::kernel_index_type;
    const std::size_t n = N - 1;
    const std::size_t boundary = static_cast<std::size_t>(plssvm::THREAD_BLOCK_SIZE) * plssvm::INTERNAL_BLOCK_SIZE;  // 96
    const std::size_t rows_padded = n + boundary;
    if (d * rows_padded >= (std::size_t{ 1 } << 31)) {
        std::fprintf(stderr, "ref_cuda_harness: d * (n + 96) exceeds the reference's int index range\n");
        return 2;
    }
    T *data_d = nullptr, *q_d = nullptr, *v_d = nullptr, *ret_d = nullptr;
    RC(cudaMalloc(&data_d, d * rows_padded * sizeof(T)));
    RC(cudaMemset(data_d, 0, d * rows_padded * sizeof(T)));
    to_soa_kernel<T><<<1184, 256>>>(X_dev_rowmajor, data_d, n, d, rows_padded);
    RC(cudaGetLastError());
    RC(cudaMalloc(&q_d, rows_padded * sizeof(T)));
    RC(cudaMalloc(&v_d, rows_padded * sizeof(T)));
    RC(cudaMalloc(&ret_d, rows_padded * sizeof(T)));
    RC(cudaMemset(q_d, 0, rows_padded * sizeof(T)));
    RC(cudaMemset(v_d, 0, rows_padded * sizeof(T)));
    RC(cudaMemcpy(q_d, q, n * sizeof(T), cudaMemcpyHostToDevice));
    RC(cudaMemcpy(v_d, v, n * sizeof(T), cudaMemcpyHostToDevice));

    const auto grid_side = static_cast<unsigned>(std::ceil(static_cast<T>(n) / static_cast<T>(boundary)));  // gpu_csvm.hpp:443
    const dim3 grid(grid_side, grid_side), block(plssvm::THREAD_BLOCK_SIZE, plssvm::THREAD_BLOCK_SIZE);
    cudaEvent_t e0, e1;
    RC(cudaEventCreate(&e0));
    RC(cudaEventCreate(&e1));
    float total_ms = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        RC(cudaMemset(ret_d, 0, rows_padded * sizeof(T)));
        RC(cudaMemcpy(ret_d, ret_inout, n * sizeof(T), cudaMemcpyHostToDevice));
        RC(cudaEventRecord(e0));
        switch (kernel) {
            case 0:
                plssvm::cuda::device_kernel_linear<<<grid, block>>>(q_d, ret_d, v_d, data_d, QA_cost, cost_inv, static_cast<kernel_index_type>(rows_padded), static_cast<kernel_index_type>(d), add, 0);
                break;
            case 1:
                plssvm::cuda::device_kernel_polynomial<<<grid, block>>>(q_d, ret_d, v_d, data_d, QA_cost, cost_inv, static_cast<kernel_index_type>(rows_padded), static_cast<kernel_index_type>(d), add, degree, gamma, coef0);
                break;
            default:
                plssvm::cuda::device_kernel_rbf<<<grid, block>>>(q_d, ret_d, v_d, data_d, QA_cost, cost_inv, static_cast<kernel_index_type>(rows_padded), static_cast<kernel_index_type>(d), add, gamma);
                break;
        }
        RC(cudaEventRecord(e1));
        RC(cudaEventSynchronize(e1));
        RC(cudaGetLastError());
        float ms = 0.f;
        RC(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 || reps == 1) { total_ms += ms; }
    }
    if (ms_out != nullptr) { *ms_out = total_ms / static_cast<float>(reps > 1 ? reps - 1 : 1); }
    RC(cudaMemcpy(ret_inout, ret_d, n * sizeof(T), cudaMemcpyDeviceToHost));
    cudaFree(data_d);
    cudaFree(q_d);
    cudaFree(v_d);
    cudaFree(ret_d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // namespace

extern "C" {
// X: DEVICE pointer to the row-major N x d matrix; q, v, ret: HOST pointers (n = N - 1 entries).  ret += add * Q~ v.
int refcuda_matvec_f64(int kernel, const double *X_dev, std::size_t N, std::size_t d, const double *q, const double *v, double QA_cost, double cost_inv, double add, int degree,
                       double gamma, double coef0, double *ret_inout, float *ms_out, int reps) {
    return run<double>(kernel, X_dev, N, d, q, v, QA_cost, cost_inv, add, degree, gamma, coef0, ret_inout, ms_out, reps);
}
int refcuda_matvec_f32(int kernel, const float *X_dev, std::size_t N, std::size_t d, const float *q, const float *v, float QA_cost, float cost_inv, float add, int degree,
                       float gamma, float coef0, float *ret_inout, float *ms_out, int reps) {
    return run<float>(kernel, X_dev, N, d, q, v, QA_cost, cost_inv, add, degree, gamma, coef0, ret_inout, ms_out, reps);
}
}
