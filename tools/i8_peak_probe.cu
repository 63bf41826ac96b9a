// Measures the int8 tensor-core roofline denominators of the fp64 int8-slice path (plssvm_b200/csrc/tile_i8.cuh) on this GPU:
//   * tcgen05.mma kind::i8 issue loops (no global traffic): M = 128, N = 256 / 64, zero and random operands, burst and sustained
//   * cuBLAS int8 GEMM (cublasGemmEx, CUBLAS_COMPUTE_32I) 8192^3 as the library figure
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/i8_peak_probe tools/i8_peak_probe.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e = (x);                                                                         \
        if (e != cudaSuccess) {                                                                      \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            std::exit(1);                                                                            \
        }                                                                                            \
    } while (0)

__device__ __forceinline__ std::uint32_t smem_u32(const void *p) { return static_cast<std::uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ std::uint64_t desc_sw64(const std::uint32_t a) {
    return static_cast<std::uint64_t>((a & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<std::uint64_t>(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__host__ __device__ constexpr std::uint32_t idesc(const std::uint32_t n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | (8u << 24); }

// one CTA per SM; lane 0 of warp 0 issues `iters` x 8 MMAs (M = 128, N, K = 32) on operands resident in shared memory
template <int N>
__global__ void __launch_bounds__(128, 1) i8_issue_loop(const int iters, const unsigned seed, unsigned *sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ std::uint64_t bar;
    __shared__ std::uint32_t tmem_slot;
    // 8 A slices (128 x 64 B) + 256 B rows x 64 B
    constexpr int BYTES = 8 * 8192 + 256 * 64;
    for (int i = threadIdx.x; i < BYTES / 4; i += blockDim.x) {
        unsigned h = (i + 1) * 2654435761u ^ seed * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<unsigned *>(smem)[i] = seed == 0 ? 0u : h;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const std::uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const std::uint32_t base = smem_u32(smem);
        const std::uint64_t db = desc_sw64(base + 8 * 8192);
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int s = 0; s < 8; ++s) {
                const std::uint64_t da = desc_sw64(base + s * 8192) + static_cast<std::uint64_t>((s & 1) * 2);
                const std::uint32_t d = tmem + static_cast<std::uint32_t>((s & 1) * 256);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc(N)), "r"(it > 0 ? 1u : 0u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        std::uint32_t ok = 0;
        while (ok == 0) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        std::uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (r == 0x12345678u) { sink[blockIdx.x] = r; }
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <typename F>
double time_ms(F &&f, const int reps, const bool best) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double acc = 0.0, mn = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        acc += ms;
        mn = ms < mn ? ms : mn;
    }
    return best ? mn : acc / reps;
}

int main() {
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned *sink;
    CK(cudaMalloc(&sink, sms * sizeof(unsigned)));
    constexpr int SMEM = 1024 + 8 * 8192 + 256 * 64;
    CK(cudaFuncSetAttribute(i8_issue_loop<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(i8_issue_loop<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(i8_issue_loop<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    std::printf("{\n \"gpu\": \"%s\", \"sms\": %d,\n", prop.name, sms);
    const int iters = 20000;
    auto tops = [&](const int n, const double ms) { return 2.0 * 128 * n * 32 * 8.0 * iters * sms / ms * 1e-9; };
    for (unsigned seed : { 0u, 7u }) {
        const char *tag = seed == 0 ? "zero" : "random";
        i8_issue_loop<256><<<sms, 128, SMEM>>>(100, seed, sink);
        CK(cudaDeviceSynchronize());
        double ms = time_ms([&] { i8_issue_loop<256><<<sms, 128, SMEM>>>(iters, seed, sink); }, 3, true);
        std::printf(" \"i8_mma_n256_%s_tops_burst\": %.1f,\n", tag, tops(256, ms));
        ms = time_ms([&] { i8_issue_loop<128><<<sms, 128, SMEM>>>(iters, seed, sink); }, 3, true);
        std::printf(" \"i8_mma_n128_%s_tops_burst\": %.1f,\n", tag, tops(128, ms));
        ms = time_ms([&] { i8_issue_loop<64><<<sms, 128, SMEM>>>(iters, seed, sink); }, 3, true);
        std::printf(" \"i8_mma_n64_%s_tops_burst\": %.1f,\n", tag, tops(64, ms));
        // sustained: ~3 s of back-to-back launches
        const auto t0 = std::chrono::steady_clock::now();
        double acc = 0.0;
        int cnt = 0;
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 3.0) {
            acc += time_ms([&] { i8_issue_loop<256><<<sms, 128, SMEM>>>(iters, seed, sink); }, 1, true);
            ++cnt;
        }
        std::printf(" \"i8_mma_n256_%s_tops_sustained_3s\": %.1f,\n", tag, tops(256, acc / cnt));
    }
    CK(cudaGetLastError());
    // cuBLAS int8 GEMM
    {
        const int n = 8192;
        std::int8_t *A, *B;
        std::int32_t *C;
        CK(cudaMalloc(&A, size_t(n) * n));
        CK(cudaMalloc(&B, size_t(n) * n));
        CK(cudaMalloc(&C, size_t(n) * n * 4));
        std::vector<std::int8_t> h(size_t(n) * n);
        unsigned s = 12345u;
        for (auto &x : h) { s = s * 1664525u + 1013904223u; x = static_cast<std::int8_t>(s >> 24); }
        CK(cudaMemcpy(A, h.data(), h.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data(), h.size(), cudaMemcpyHostToDevice));
        cublasHandle_t hd;
        cublasCreate(&hd);
        const std::int32_t one = 1, zero = 0;
        auto run = [&] { return cublasGemmEx(hd, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, CUDA_R_8I, n, B, CUDA_R_8I, n, &zero, C, CUDA_R_32I, n, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT); };
        const cublasStatus_t st = run();
        if (st == CUBLAS_STATUS_SUCCESS && cudaDeviceSynchronize() == cudaSuccess) {
            const double best = time_ms([&] { run(); }, 5, true);
            const auto t0 = std::chrono::steady_clock::now();
            double acc = 0.0;
            int cnt = 0;
            while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 3.0) {
                acc += time_ms([&] { run(); }, 1, true);
                ++cnt;
            }
            std::printf(" \"cublas_i8gemm_8192_random_tops_burst\": %.1f,\n \"cublas_i8gemm_8192_random_tops_sustained_3s\": %.1f\n", 2.0 * n * n * n / best * 1e-9, 2.0 * n * n * n / (acc / cnt) * 1e-9);
        } else {
            std::printf(" \"cublas_i8gemm_8192_random_tops_burst\": null\n");
        }
    }
    std::printf("}\n");
    return 0;
}
