// Measures the roofline denominators MEASURED_PEAKS.json does not hold: FP64 DMMA / DFMA issue peaks,
// FP32 FFMA peak, and cuBLAS DGEMM / SGEMM / TF32 GEMM 8192^3 — burst (best of N) and sustained (back to back for ~3 s).
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/peak_probe tools/peak_probe.cu -lcublas
// Run on the GPU box:  tools/peak_probe > gpurun_out/peaks.json
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            std::exit(1);                                                           \
        }                                                                           \
    } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dmma_loop(double *out, int iters) {
    double c[ILP][2];
    #pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) { s += c[i][0] + c[i][1]; }
    if (s == 123.456) { out[0] = s; }
}

template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_loop(T *out, int iters) {
    T c[ILP];
    #pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i] = T(i); }
    T a = T(1.0) + T(threadIdx.x) * T(1e-7), b = T(0.5);
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < ILP; ++i) { c[i] = c[i] * a + b; }
    }
    T s = 0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) { s += c[i]; }
    if (s == T(123.456)) { out[0] = s; }
}

// uniform(-1, 1) fill (LCG): real operand bit patterns, so tensor-core GEMMs draw realistic power
template <typename T>
__global__ void fill_random(T *p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        unsigned x = (unsigned) (i * 2654435761u) ^ seed;
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        p[i] = T((x & 0xFFFFFF) / double(0x800000) - 1.0);
    }
}

template <typename F>
double time_ms(F &&f, int reps, bool best) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    double bestms = 1e30, sum = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        bestms = std::min<double>(bestms, ms);
        sum += ms;
    }
    return best ? bestms : sum / reps;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *dout;
    CK(cudaMalloc(&dout, 1024));
    std::printf("{\n \"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);

    // --- issue-rate probes: `wps` warps per SM sub-partition, 16 independent accumulators per warp
    for (int wps : { 1, 2 }) {
        const int iters = 20000;
        const int blocks = sms, threads = 128 * wps > 1024 ? 1024 : 128 * wps;
        const int bl = 128 * wps > 1024 ? blocks * (128 * wps / 1024) : blocks;
        const double ms = time_ms([&] { dmma_loop<16><<<bl, threads>>>(dout, iters); }, 5, true);
        const double flops = 2.0 * 256 * 16.0 * iters * (double(bl) * threads / 32);
        std::printf(" \"dmma_tflops_%dwarp_per_smsp\": %.2f,\n", wps, flops / ms * 1e-9);
    }
    {
        const int iters = 20000;
        const double ms = time_ms([&] { fma_loop<double, 16><<<sms * 2, 256>>>(dout, iters); }, 5, true);
        std::printf(" \"dfma_tflops\": %.2f,\n", 2.0 * 16 * iters * double(sms) * 2 * 256 / ms * 1e-9);
        const double ms2 = time_ms([&] { fma_loop<float, 16><<<sms * 4, 256>>>(reinterpret_cast<float *>(dout), iters); }, 5, true);
        std::printf(" \"ffma_tflops\": %.2f,\n", 2.0 * 16 * iters * double(sms) * 4 * 256 / ms2 * 1e-9);
    }
    // sustained DMMA for ~3 s (clock behaviour under the power cap)
    {
        const int iters = 200000;
        const double ms = time_ms([&] { dmma_loop<16><<<sms, 256>>>(dout, iters); }, 1, true);
        int reps = std::max(1, int(3000.0 / ms));
        const double avg = time_ms([&] { dmma_loop<16><<<sms, 256>>>(dout, iters); }, reps, false);
        std::printf(" \"dmma_tflops_sustained_3s\": %.2f,\n", 2.0 * 256 * 16.0 * iters * (double(sms) * 8) / avg * 1e-9);
    }

    // --- cuBLAS GEMMs 8192^3
    cublasHandle_t h;
    cublasCreate(&h);
    const int n = 8192;
    {
        double *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(double) * n * n));
        CK(cudaMalloc(&B, sizeof(double) * n * n));
        CK(cudaMalloc(&C, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n));
        CK(cudaMemset(B, 0, sizeof(double) * n * n));
        const double one = 1.0, zero = 0.0;
        auto run = [&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); };
        const double best = time_ms(run, 5, true);
        const double ms1 = time_ms(run, 1, true);
        const double avg = time_ms(run, std::max(1, int(3000.0 / ms1)), false);
        std::printf(" \"cublas_dgemm_tflops_burst\": %.2f,\n \"cublas_dgemm_tflops_sustained_3s\": %.2f,\n", 2.0 * n * n * n / best * 1e-9, 2.0 * n * n * n / avg * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
        float *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(float) * n * n));
        CK(cudaMalloc(&B, sizeof(float) * n * n));
        CK(cudaMalloc(&C, sizeof(float) * n * n));
        CK(cudaMemset(A, 0, sizeof(float) * n * n));
        CK(cudaMemset(B, 0, sizeof(float) * n * n));
        const float one = 1.f, zero = 0.f;
        auto run = [&] { cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); };
        cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
        const double fp32 = time_ms(run, 5, true);
        cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
        const double tf32 = time_ms(run, 5, true);
        const double ms1 = time_ms(run, 1, true);
        const double tf32_avg = time_ms(run, std::max(1, int(3000.0 / ms1)), false);
        std::printf(" \"cublas_sgemm_fp32_tflops\": %.2f,\n \"cublas_sgemm_tf32_tflops_burst\": %.2f,\n \"cublas_sgemm_tf32_tflops_sustained_3s\": %.2f,\n", 2.0 * n * n * n / fp32 * 1e-9,
                    2.0 * n * n * n / tf32 * 1e-9, 2.0 * n * n * n / tf32_avg * 1e-9);
        // the same with uniform(-1, 1) operands: the power-limited number a real TF32 kernel can be compared with
        fill_random<float><<<1184, 256>>>(A, size_t(n) * n, 1u);
        fill_random<float><<<1184, 256>>>(B, size_t(n) * n, 2u);
        CK(cudaDeviceSynchronize());
        const double tf32_rb = time_ms(run, 5, true);
        const double ms2 = time_ms(run, 1, true);
        const double tf32_ravg = time_ms(run, std::max(1, int(4000.0 / ms2)), false);
        std::printf(" \"cublas_sgemm_tf32_random_tflops_burst\": %.2f,\n \"cublas_sgemm_tf32_random_tflops_sustained_4s\": %.2f\n", 2.0 * n * n * n / tf32_rb * 1e-9, 2.0 * n * n * n / tf32_ravg * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    std::printf("}\n");
    return 0;
}
