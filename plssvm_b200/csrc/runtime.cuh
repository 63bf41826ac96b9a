// Host runtime of libplssvm_b200.so: error handling, NCCL through dlopen, the per-device context with its caching allocator and
// event timers, device groups (one process driving several GPUs, one host thread per device), and pinned staging of host rows.
//
// Reference counterparts: cuda::csvm::init (src/plssvm/backends/CUDA/csvm.cu:48-86: device discovery, one queue per device),
// gpu_device_ptr (include/plssvm/backends/gpu_device_ptr.hpp:28-239), the `#pragma omp parallel for` over devices of
// gpu_csvm.hpp:331,369,521,574 and the host-staged device_reduction (gpu_csvm.hpp:449-475) — here NCCL over NVLink.
#pragma once

#include "../../include/plssvm_b200.h"

#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace pbrt {

inline thread_local std::string g_last_error;

struct api_error : std::runtime_error {
    int code;
    api_error(const int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

#define PB_CUDA(call)                                                                                                                     \
    do {                                                                                                                                  \
        const cudaError_t err__ = (call);                                                                                                 \
        if (err__ != cudaSuccess) {                                                                                                       \
            throw ::pbrt::api_error(PLSSVM_B200_ERR_CUDA, std::string("CUDA assert '") + cudaGetErrorName(err__) + "' (" + std::to_string(static_cast<int>(err__)) + "): " + \
                                                              cudaGetErrorString(err__) + " [" #call "]");                               \
        }                                                                                                                                 \
    } while (0)

#define PB_REQUIRE(cond, msg)                                                      \
    do {                                                                           \
        if (!(cond)) { throw ::pbrt::api_error(PLSSVM_B200_ERR_INVALID, (msg)); }  \
    } while (0)

// ---- NCCL through dlopen (torch ships libnccl.so.2; nothing to link at build time) ---------------------------------------
struct nccl_api {
    using comm_t = void *;
    struct unique_id {
        char internal[128];
    };
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *handle = nullptr;

    static nccl_api &get() {
        static nccl_api api = load();
        return api;
    }
    static nccl_api load() {
        nccl_api a;
        const char *names[] = { std::getenv("PLSSVM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char *nm : names) {
            if (nm == nullptr) { continue; }
            a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle != nullptr) { break; }
        }
        if (a.handle == nullptr) { throw api_error(PLSSVM_B200_ERR_CUDA, "cannot load libnccl.so.2 (set PLSSVM_B200_NCCL_LIB)"); }
        auto sym = [&](const char *name) {
            void *p = dlsym(a.handle, name);
            if (p == nullptr) { throw api_error(PLSSVM_B200_ERR_CUDA, std::string("libnccl is missing the symbol ") + name); }
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        return a;
    }
    void check(const int rc, const char *what) const {
        if (rc != 0) { throw api_error(PLSSVM_B200_ERR_CUDA, std::string("NCCL failure in ") + what + ": " + GetErrorString(rc)); }
    }
};
constexpr int NCCL_INT8 = 0, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0;

// ---- CUDA-event timer: pairs recorded on the launching stream, elapsed time summed incrementally (every pair is read once) ----
struct event_pair_timer {
    std::vector<cudaEvent_t> ev;
    std::size_t used = 0, read = 0;  // events handed out / events already summed into accum_ms
    double accum_ms = 0.0;
    ~event_pair_timer() {
        for (cudaEvent_t e : ev) { cudaEventDestroy(e); }
    }
    void reset() {
        used = read = 0;
        accum_ms = 0.0;
    }
    void begin(cudaStream_t s) {
        if (ev.size() < used + 2) {
            cudaEvent_t a, b;
            PB_CUDA(cudaEventCreate(&a));
            PB_CUDA(cudaEventCreate(&b));
            ev.push_back(a);
            ev.push_back(b);
        }
        PB_CUDA(cudaEventRecord(ev[used], s));
    }
    void end(cudaStream_t s) {
        PB_CUDA(cudaEventRecord(ev[used + 1], s));
        used += 2;
    }
    // sums the pairs that have completed since the last call (pairs still in flight are left for the next call) and recycles their events
    double collect() {
        double t = 0.0;
        for (; read + 1 < used; read += 2) {
            const cudaError_t q = cudaEventQuery(ev[read + 1]);
            if (q == cudaErrorNotReady) {
                (void) cudaGetLastError();
                break;
            }
            PB_CUDA(q);
            float ms = 0.f;
            PB_CUDA(cudaEventElapsedTime(&ms, ev[read], ev[read + 1]));
            t += ms;
        }
        if (read == used) { used = read = 0; }
        accum_ms += t;
        return t;
    }
    double total_ms() {
        collect();
        return accum_ms;
    }
};

}  // namespace pbrt

// ---- opaque handles ---------------------------------------------------------------------------------------------------------
struct plssvm_b200_dataset;

struct plssvm_b200_ctx {
    int device = 0;
    int num_sms = 0;
    int pairs_ok = 0;  // device can launch 2-CTA clusters (the CTA-pair int8-slice kernel)
    cudaStream_t stream = nullptr;
    int rank = 0, world = 1;
    pbrt::nccl_api::comm_t comm = nullptr;
    // in-process device group (plssvm_b200_create with n_dev > 1): the handle the caller holds is members[0]; every member is a
    // full context of its own device with rank = its index and a communicator from ncclCommInitAll
    std::vector<plssvm_b200_ctx *> members;
    plssvm_b200_ctx *leader = nullptr;
    // host-side agreement of the member threads before they enter a collective (leader only; see group_all_ok)
    std::mutex agree_mutex;
    std::condition_variable agree_cv;
    int agree_arrived = 0, agree_generation = 0;
    bool agree_failed = false, agree_result = false;
    bool in_process_group() const { return leader != nullptr && leader->members.size() > 1; }
    // options
    int impl = 0;            // 0 auto, 1 simt, 2 floating-point tensor tiles, 6 int8 slices on tcgen05 (tile_i8.cuh), 7 the same with the exact-input slice count, 10 int8 slices on CTA pairs (tile_i8_pair.cuh)
    int check_interval = 0;  // 0 = auto
    int verbose = 0;
    int ignore_convergence = 0;  // benchmarking: never set the convergence flag, so exactly the requested number of iterations runs
    int max_ctas = 0;            // debugging: cap the grid of the tile kernels (0 = one CTA per SM)
    int linear_factorized = 0;   // 1: linear-kernel matvec as X (X^T v) (two streaming passes) instead of the implicit tiles
    int balance = 1;             // several ranks: re-cut the tile shares from the measured tile-kernel rates every `balance_interval` iterations
    int balance_interval = 8;
    int i8_a_via_tmem = 0;       // experiment: fp64 int8-slice tiles read the doubly-used A planes from tensor memory (tcgen05.cp + TS-form MMA)
    int tile_stats = 0;          // profiling: per-role wait-cycle counters of the int8-slice tile kernel (synchronises after every tile launch)
    int fp32_pair = 1;           // automatic kernel choice, fp32: int8-slice tiles on CTA pairs (impl 10) instead of single CTAs (impl 6)
    int fp32_fast_drain = 1;     // fp32 int8-slice epilogue: release TMEM before the fp64 -> fp32 conversion (0: the round-1 order, for A/B measurements)
    int virtual_skew = 0;        // testing aid (virtual ranks): percent by which the tile shares grow from the first to the last rank
    int shard_upload = 1;        // several ranks: every rank uploads 1 / world of the rows over its own PCIe link, ncclAllGather over NVLink
    // timings of the last call (accumulated over the lifetime of an open CG session)
    plssvm_b200_timings tm{};
    pbrt::event_pair_timer tile_timer, matvec_timer;
    int open_sessions = 0;
    std::vector<plssvm_b200_dataset *> live_datasets;  // (leader) data sets and CG sessions created from this context: destroyed with it
    std::vector<struct plssvm_b200_cg *> live_sessions;
    std::vector<double> last_trace;  // residual history r.r of the last finished solve (entry k: after k iterations)
    cudaEvent_t ev_loop0 = nullptr, ev_loop1 = nullptr;
    PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
    void *pinned = nullptr;  // small pinned staging block for scalar read-backs
    // second stream + events: H2D staging of predict batches overlaps the tile kernel of the previous batch
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_computed[2] = { nullptr, nullptr };
    // pinned ring for host rows that are not one pinned block (std::vector<std::vector<T>> rows, pageable flat buffers)
    static constexpr std::size_t RING_BYTES = std::size_t{ 8 } << 20;  // (pinning host memory costs ~0.3 ms per MB and is serialised across the devices of a group)
    void *ring[2] = { nullptr, nullptr };
    cudaEvent_t ev_ring[2] = { nullptr, nullptr };
    // grow-only device workspaces kept across calls
    enum ws_slot { WS_PARTIAL = 0, WS_OUT, WS_ALPHA, WS_W, WS_STAGE0, WS_STAGE1, WS_SQ0, WS_SQ1, WS_HI0, WS_HI1, WS_LO0, WS_LO1, WS_I8_0, WS_I8_1, WS_SC0, WS_SC1, WS_MISC, WS_STATS, WS_COUNT };
    void *ws_ptr[WS_COUNT] = {};
    std::size_t ws_bytes[WS_COUNT] = {};
    // caching allocator: blocks released by RAII buffers are kept for the next call (cudaMalloc / cudaFree of multi-GB buffers cost
    // milliseconds and synchronise the device); everything runs on `stream`, so reuse is stream-ordered
    struct block {
        void *p;
        std::size_t bytes;
    };
    std::vector<block> pool;
    std::size_t pool_bytes = 0;
};

namespace pbrt {

inline void *pool_acquire(plssvm_b200_ctx *ctx, const std::size_t bytes) {
    if (bytes == 0) { return nullptr; }
    int best = -1;
    for (int i = 0; i < static_cast<int>(ctx->pool.size()); ++i) {
        const std::size_t b = ctx->pool[i].bytes;
        if (b >= bytes && b <= bytes + bytes / 4 + (std::size_t{ 1 } << 20) && (best < 0 || b < ctx->pool[best].bytes)) { best = i; }
    }
    if (best >= 0) {
        void *p = ctx->pool[best].p;
        ctx->pool_bytes -= ctx->pool[best].bytes;
        ctx->pool.erase(ctx->pool.begin() + best);
        return p;
    }
    void *p = nullptr;
    cudaError_t err = cudaMalloc(&p, bytes);
    if (err == cudaErrorMemoryAllocation) {  // hand the cached blocks back to the driver and retry once
        (void) cudaGetLastError();
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (const auto &b : ctx->pool) { cudaFree(b.p); }
        ctx->pool.clear();
        ctx->pool_bytes = 0;
        err = cudaMalloc(&p, bytes);
    }
    PB_CUDA(err);
    return p;
}
inline void pool_release(plssvm_b200_ctx *ctx, void *p, const std::size_t bytes) {
    if (p == nullptr) { return; }
    ctx->pool.push_back({ p, bytes });
    ctx->pool_bytes += bytes;
}
inline void pool_trim(plssvm_b200_ctx *ctx) {
    for (const auto &b : ctx->pool) { cudaFree(b.p); }
    ctx->pool.clear();
    ctx->pool_bytes = 0;
}

// ---- RAII device buffer (reference: gpu_device_ptr.hpp:28-239) backed by the context's caching allocator -------------------------
template <typename T>
struct dbuf {
    T *p = nullptr;
    std::size_t count = 0;
    plssvm_b200_ctx *ctx = nullptr;
    dbuf() = default;
    dbuf(plssvm_b200_ctx *c, const std::size_t n) { alloc(c, n); }
    dbuf(const dbuf &) = delete;
    dbuf &operator=(const dbuf &) = delete;
    dbuf(dbuf &&o) noexcept : p(o.p), count(o.count), ctx(o.ctx) {
        o.p = nullptr;
        o.count = 0;
    }
    dbuf &operator=(dbuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            count = o.count;
            ctx = o.ctx;
            o.p = nullptr;
            o.count = 0;
        }
        return *this;
    }
    ~dbuf() { release(); }
    void alloc(plssvm_b200_ctx *c, const std::size_t n) {
        release();
        ctx = c;
        count = n;
        if (n > 0) { p = static_cast<T *>(pool_acquire(c, n * sizeof(T))); }
    }
    void release() {
        if (p != nullptr) { pool_release(ctx, p, count * sizeof(T)); }
        p = nullptr;
        count = 0;
    }
};

template <typename T>
T *workspace(plssvm_b200_ctx *ctx, const int slot, const std::size_t count) {
    const std::size_t bytes = count * sizeof(T);
    if (ctx->ws_bytes[slot] < bytes) {
        if (ctx->ws_ptr[slot] != nullptr) {
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
            PB_CUDA(cudaFree(ctx->ws_ptr[slot]));
            ctx->ws_ptr[slot] = nullptr;
            ctx->ws_bytes[slot] = 0;
        }
        PB_CUDA(cudaMalloc(&ctx->ws_ptr[slot], bytes));
        ctx->ws_bytes[slot] = bytes;
    }
    return static_cast<T *>(ctx->ws_ptr[slot]);
}

// ---- device groups: run f(member context, rank) for every member, one host thread per device ------------------------------------
// (the reference: one OpenMP thread per device, gpu_csvm.hpp:331,369,521,574).  The first exception is rethrown on the caller.
template <typename F>
void for_each_rank(plssvm_b200_ctx *ctx, F &&f) {
    const std::size_t G = ctx->members.size();
    if (G <= 1) {
        PB_CUDA(cudaSetDevice(ctx->device));
        f(ctx, 0);
        return;
    }
    std::vector<std::exception_ptr> errs(G);
    std::vector<std::thread> th;
    th.reserve(G - 1);
    auto body = [&](const std::size_t g) {
        try {
            PB_CUDA(cudaSetDevice(ctx->members[g]->device));
            f(ctx->members[g], static_cast<int>(g));
        } catch (...) {
            errs[g] = std::current_exception();
        }
    };
    for (std::size_t g = 1; g < G; ++g) { th.emplace_back(body, g); }
    body(0);
    for (auto &t : th) { t.join(); }
    cudaSetDevice(ctx->device);
    for (const auto &e : errs) {
        if (e) { std::rethrow_exception(e); }
    }
}

// Device group only: every member thread reports whether its local preparation (allocations, argument checks) succeeded and learns whether ALL
// did — called before the first collective of an operation, so that a failure on one device becomes an error on every device instead of the
// others waiting forever inside NCCL.  (One process per GPU: the ranks are separate processes; a failing rank takes its job down.)
inline bool group_all_ok(plssvm_b200_ctx *ctx, const bool ok) {
    if (!ctx->in_process_group()) { return ok; }
    plssvm_b200_ctx *L = ctx->leader;
    const int n = static_cast<int>(L->members.size());
    std::unique_lock<std::mutex> lock(L->agree_mutex);
    const int gen = L->agree_generation;
    L->agree_failed = L->agree_failed || !ok;
    if (++L->agree_arrived == n) {
        L->agree_result = !L->agree_failed;
        L->agree_arrived = 0;
        L->agree_failed = false;
        ++L->agree_generation;
        L->agree_cv.notify_all();
    } else {
        L->agree_cv.wait(lock, [&] { return L->agree_generation != gen; });
    }
    return L->agree_result;
}

// ---- host rows --------------------------------------------------------------------------------------------------------------------
// A host matrix handed over either as one row-major block (`flat`, row pitch d) or as an array of row pointers (`rows`, the
// std::vector<std::vector<T>> of the reference's virtuals without an intermediate flat copy).
template <typename T>
struct host_matrix {
    const T *flat = nullptr;
    const T *const *rows = nullptr;
    std::size_t d = 0;
    bool valid() const { return flat != nullptr || rows != nullptr; }
    const T *row(const std::size_t i) const { return rows != nullptr ? rows[i] : flat + i * d; }
};

inline bool is_pinned_host(const void *p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void) cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// rows [r0, r1) of a host matrix -> device rows dst[0 .. r1-r0) with pitch ld (pad columns untouched), on stream `st`.
// One pinned block goes down with a single (2-D) copy; anything else is packed by a few host threads into a two-slot pinned ring
// whose H2D copies overlap the packing of the next chunk (the reference: host transpose + one blocking cudaMemcpy, gpu_csvm.hpp:302-346).
template <typename T>
void upload_rows(plssvm_b200_ctx *ctx, T *dst, const std::size_t ld, const host_matrix<T> &src, const std::size_t r0, const std::size_t r1, cudaStream_t st) {
    if (r1 <= r0) { return; }
    const std::size_t d = src.d, rows = r1 - r0;
    if (src.flat != nullptr && is_pinned_host(src.flat)) {
        const T *s = src.flat + r0 * d;
        if (ld == d) {
            PB_CUDA(cudaMemcpyAsync(dst, s, rows * d * sizeof(T), cudaMemcpyHostToDevice, st));
        } else {
            PB_CUDA(cudaMemcpy2DAsync(dst, ld * sizeof(T), s, d * sizeof(T), d * sizeof(T), rows, cudaMemcpyHostToDevice, st));
        }
        return;
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->ring[i] == nullptr) {
            PB_CUDA(cudaMallocHost(&ctx->ring[i], plssvm_b200_ctx::RING_BYTES));
            PB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming));
        }
    }
    const std::size_t chunk_rows = std::max<std::size_t>(1, plssvm_b200_ctx::RING_BYTES / (d * sizeof(T)));
    PB_REQUIRE(d * sizeof(T) <= plssvm_b200_ctx::RING_BYTES, "a single data point exceeds the pinned staging ring");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned packers = std::max(1u, std::min(4u, hw / static_cast<unsigned>(std::max(1, ctx->world))));  // every rank stages its own row share
    std::size_t slot = 0;
    for (std::size_t c0 = 0; c0 < rows; c0 += chunk_rows, slot ^= 1) {
        const std::size_t cr = std::min(chunk_rows, rows - c0);
        PB_CUDA(cudaEventSynchronize(ctx->ev_ring[slot]));  // the previous copy out of this slot has finished
        T *stage = static_cast<T *>(ctx->ring[slot]);
        auto pack = [&](const std::size_t a, const std::size_t b) {
            for (std::size_t i = a; i < b; ++i) { std::memcpy(stage + i * d, src.row(r0 + c0 + i), d * sizeof(T)); }
        };
        if (packers > 1 && cr * d * sizeof(T) >= (std::size_t{ 4 } << 20)) {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < packers; ++t) { th.emplace_back(pack, cr * t / packers, cr * (t + 1) / packers); }
            pack(0, cr / packers);
            for (auto &t : th) { t.join(); }
        } else {
            pack(0, cr);
        }
        PB_CUDA(cudaMemcpy2DAsync(dst + c0 * ld, ld * sizeof(T), stage, d * sizeof(T), d * sizeof(T), cr, cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaEventRecord(ctx->ev_ring[slot], st));
    }
}

struct host_timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace pbrt
