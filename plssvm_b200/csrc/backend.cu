// libplssvm_b200.so — host driver + C ABI (include/plssvm_b200.h).
//
// B200-native re-design of the reference's detail::gpu_csvm driver (include/plssvm/backends/gpu_csvm.hpp:45-730):
//   * X lives once in HBM, row-major with a 128-byte-multiple pitch (TMA boxes + coalesced row streams), plus |x_i|^2 and ONE copy
//     of the int8 digit planes that serves both operands of the tile kernel
//   * the CG loop is device-resident: all vectors AND scalars stay in HBM, the host only polls a convergence flag
//     (the reference does 3 blocking PCIe copies + host-serial vector algebra per iteration, gpu_csvm.hpp:582-633)
//   * the implicit matvec is a persistent tile kernel over the banded lower triangle (tile_i8.cuh / tile_dmma.cuh / tile_tf32.cuh /
//     tile_simt.cuh) followed by a fixed-order partial reduction — no atomics
//   * the rbf kernel is evaluated on data centred at the feature means of the training set / support vectors (exact translation
//     invariance), so the norm expansion |x_i|^2 + |x_j|^2 - 2 x_i.x_j never cancels against a common offset of the data
//   * several GPUs, in one process (device group: one host thread per device, ncclCommInitAll) or one process per GPU
//     (plssvm_b200_comm_init): each rank owns a contiguous, rate-weighted share of the tile order, one ncclAllReduce of the n-vector
//     per matvec; data sets are uploaded 1 / world per rank and all-gathered over NVLink; predict shards the test points
//     (the reference: feature split for the linear kernel only, summed through the host, gpu_csvm.hpp:283-299,449-475; predict on device 0)
// There is no CPU fallback: without a device every compute entry point fails.
#include "runtime.cuh"
#include "stream_kernels.cuh"
#include "tile_dmma.cuh"
#include "tile_simt.cuh"
#include "tile_tf32.cuh"
#include "tile_i8.cuh"
#include "tile_i8_pair.cuh"
#ifdef PLSSVM_B200_EXPERIMENTAL
    #include "tile_tf32_2sm.cuh"
    #include "tile_tf32_n256.cuh"
    #include "tile_i8_2sm.cuh"
#endif

#include <cmath>
#include <mutex>
#include <unordered_set>

using namespace pbrt;

struct dblock {
    void *p = nullptr;
    std::size_t bytes = 0;
};

struct plssvm_b200_dataset {
    plssvm_b200_ctx *ctx = nullptr;
    int elem_size = 0;  // 4 or 8
    std::size_t N = 0, d = 0, ld = 0;
    std::uint64_t id = 0;  // unique; the centre key of operands centred at THIS data set's feature means
    dblock X;              // [N (rounded up to equal row shares with several ranks)][ld]
    dblock sq;             // [N] squared row norms
    dblock mean;           // [ld] feature means, pad entries zero (lazy; rbf)
    // operands derived from X, each cached with the centre key it was made for (0 = the data as given, otherwise the id of the data
    // set whose feature means were subtracted); re-created on demand, never while a CG session uses them (pins)
    dblock sq_c;  std::uint64_t sq_c_key = 0;   // squared norms of the centred rows
    dblock Xc;    std::uint64_t xc_key = 0;     // materialised centred copy (only for the tile kernels that read X itself)
    dblock X_hi, X_lo;  std::uint64_t tf32_key = 0;  bool tf32_valid = false;  // fp32: TF32 hi / lo split for the 3xTF32 tiles
    // int8 digit planes (tile_i8.cuh) in the boxed layout of split_i8_kernel, boxes of 128 rows; X_i8b: experimental CTA-pair kernel only (boxes of 64 rows)
    dblock X_i8, X_i8b, rscale;
    std::size_t ld8 = 0;
    int i8_slices = 0, i8_br_b = 0, i8_slab = 0;
    std::uint64_t i8_key = 0;
    int i8_bad_rows = 0;  // rows whose elements are spread over too many orders of magnitude for the automatic choice (split_i8_kernel)
    int pins = 0;         // open CG sessions using the cached operands
    // device group: the handle the caller holds is members[0]
    std::vector<plssvm_b200_dataset *> members;
};

namespace {

using pb::CGState;
using pb::KernelParams;
using pb::TileParams;
using pb::TILE;

std::atomic<std::uint64_t> g_next_dataset_id{ 1 };

// Live handles (contexts, data sets, CG sessions).  Destroying a context destroys the data sets and sessions created from it; a later
// destroy / abort of such a handle (e.g. from a garbage collector that runs finalisers in any order) is a no-op instead of a use-after-free,
// and any other use of a dead handle is reported as an error.
std::mutex g_handles_mutex;
std::unordered_set<const void *> g_live_handles;
void handle_add(const void *h) {
    const std::lock_guard<std::mutex> lock(g_handles_mutex);
    g_live_handles.insert(h);
}
bool handle_remove(const void *h) {
    const std::lock_guard<std::mutex> lock(g_handles_mutex);
    return g_live_handles.erase(h) > 0;
}
bool handle_live(const void *h) {
    const std::lock_guard<std::mutex> lock(g_handles_mutex);
    return h != nullptr && g_live_handles.count(h) > 0;
}

void blk_alloc(plssvm_b200_ctx *ctx, dblock &b, const std::size_t bytes) {
    if (b.p != nullptr) { pool_release(ctx, b.p, b.bytes); }
    b.p = pool_acquire(ctx, bytes);
    b.bytes = bytes;
}
void blk_free(plssvm_b200_ctx *ctx, dblock &b) {
    if (b.p != nullptr) { pool_release(ctx, b.p, b.bytes); }
    b = dblock{};
}

template <typename T>
std::size_t pitch_elems(const std::size_t d) {
    const std::size_t per128 = 128 / sizeof(T);
    return (d + per128 - 1) / per128 * per128;
}

// 2-D map over a row-major matrix: box = 128 rows x 128 bytes (16 doubles / 32 floats), 128-byte swizzle, OOB rows zero-filled
template <typename T>
void make_tensor_map(plssvm_b200_ctx *ctx, CUtensorMap *tm, const T *base, const std::size_t rows, const std::size_t ld) {
    const cuuint64_t dims[2] = { static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows) };
    const cuuint64_t strides[1] = { static_cast<cuuint64_t>(ld * sizeof(T)) };
    const cuuint32_t box[2] = { static_cast<cuuint32_t>(128 / sizeof(T)), static_cast<cuuint32_t>(TILE) };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult rc = ctx->encode_tiled(tm, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<T *>(base), dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { throw api_error(PLSSVM_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(rc))); }
}

// int8 digit planes (tile_i8.cuh): features padded to whole 64-byte slabs, rows to whole 128-row boxes
inline std::size_t pitch_i8(const std::size_t d) { return (d + 63) / 64 * 64; }
inline std::size_t rows_i8(const std::size_t rows) { return (rows + 127) / 128 * 128; }

#ifdef PLSSVM_B200_EXPERIMENTAL
constexpr bool EXPERIMENTAL = true;
#else
constexpr bool EXPERIMENTAL = false;
#endif

// The CTA-pair int8-slice kernel (impl 10, tile_i8_pair.cuh) is the automatic choice for fp32 (measured 4 - 5 % faster than the single-CTA kernel at
// C3); the fp64 instantiation is measured SLOWER (88 vs 105 TFLOP/s at C2) and only exists in builds with -DPLSSVM_B200_EXPERIMENTAL.
template <typename T>
constexpr bool pair_kernel_built() { return sizeof(T) == 4 || EXPERIMENTAL; }

// number of int8 slices per operand for the tile-kernel choice `impl` (6: default, 7: exact-input count; the same for fp64)
template <typename T>
int i8_slices_for(const int impl) { return impl == 7 ? pb::I8<T>::S_EXACT : pb::I8<T>::S; }
// int8-slice tile kernels: 6 default slice count, 7 exact-input slice count; experimental: 8 default slice count with 2 x 2 CTA clusters + TMA multicast,
// 9 (fp32) CTA pairs with tcgen05.mma.cta_group::2 (tile_i8_2sm.cuh); 10: CTA pairs with the wide-N instructions (tile_i8_pair.cuh), default slice count
inline bool is_i8(const int impl) { return impl >= 6 && impl <= 11; }
// kernels whose tile range / ownership is over 256 x 256 super-tiles
inline bool super_tiled(const int impl) { return impl == 4 || impl == 5 || impl == 8 || impl == 9 || impl == 10 || impl == 11; }
// rows per box of the extra B-operand copy of the digit planes (experimental CTA-pair kernel only); TILE = no extra copy
template <typename T>
int i8_br_b_for(const int impl) { return impl == 9 ? 64 : TILE; }
// bytes (= features) per slab of the plane layout a tile kernel stages: the experimental CTA-pair kernel was written for 64-byte slabs
template <typename T>
int i8_slab_for(const int impl) { return impl == 9 ? 64 : (impl == 10 ? pb::I8PairConfig<T>::BK : pb::I8<T>::BK); }

// rows -> int8 digit planes + row scales (tile_i8.cuh)
template <typename T>
void run_split_i8(plssvm_b200_ctx *ctx, const int slices, const T *X, const std::size_t rows, const std::size_t d, const std::size_t ld, std::int8_t *planes_a,
                  std::int8_t *planes_b, const int br_b, const std::size_t ld8, T *rscale, int *bad_rows, const T *mean, const int slab, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(rows_i8(rows) / 8);  // incl. the padding rows of the last box, which get zero digits
    const std::uint32_t d32 = static_cast<std::uint32_t>(d), ld32 = static_cast<std::uint32_t>(ld), slab_w = static_cast<std::uint32_t>(slab);
    const std::uint32_t slabs = static_cast<std::uint32_t>(ld8) / slab_w;
    if (slices == pb::I8<T>::S) {
        pb::split_i8_kernel<T, pb::I8<T>::S><<<grid, 256, 0, st>>>(X, rows, d32, ld32, planes_a, planes_b, static_cast<std::uint32_t>(br_b), slabs, rscale, bad_rows, mean, slab_w);
    } else {
        pb::split_i8_kernel<T, pb::I8<T>::S_EXACT><<<grid, 256, 0, st>>>(X, rows, d32, ld32, planes_a, planes_b, static_cast<std::uint32_t>(br_b), slabs, rscale, bad_rows, mean, slab_w);
    }
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches++;
}

void require_unpinned(const plssvm_b200_dataset *ds, const char *what) {
    PB_REQUIRE(ds->pins == 0, std::string("the data set is in use by an open CG session; ") + what + " would replace the operands that session reads (finish or abort it first, or keep "
                                  "the same kernel function and \"impl\" option while it is open)");
}

// feature means of a data set (all N rows), pad entries zero: fixed-order column sums (the w-kernel with unit weights) / N
template <typename T>
const T *ensure_mean(plssvm_b200_ctx *ctx, plssvm_b200_dataset *ds) {
    if (ds->mean.p != nullptr) { return static_cast<const T *>(ds->mean.p); }
    blk_alloc(ctx, ds->mean, ds->ld * sizeof(T));
    cudaStream_t st = ctx->stream;
    const std::uint32_t d = static_cast<std::uint32_t>(ds->d);
    const std::uint32_t chunks = static_cast<std::uint32_t>((ds->N + pb::W_ROWS - 1) / pb::W_ROWS);
    dbuf<T> part(ctx, static_cast<std::size_t>(chunks) * d);
    T *mean = static_cast<T *>(ds->mean.p);
    PB_CUDA(cudaMemsetAsync(mean, 0, ds->ld * sizeof(T), st));
    pb::w_partial_kernel<T><<<dim3((d + 255) / 256, chunks), 256, 0, st>>>(static_cast<const T *>(ds->X.p), nullptr, ds->N, d, static_cast<std::uint32_t>(ds->ld), part.p);
    pb::w_reduce_kernel<T><<<(d + 31) / 32, 256, 0, st>>>(part.p, chunks, d, mean);
    pb::scale_vec_kernel<T><<<(d + 255) / 256, 256, 0, st>>>(mean, T(1) / static_cast<T>(ds->N), d);
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches += 3;
    return mean;
}

// What a tile kernel needs of one operand (rows of A or of B)
template <typename T>
struct operand {
    const T *X = nullptr, *sq = nullptr, *hi = nullptr, *lo = nullptr, *scale = nullptr;
    const std::int8_t *i8 = nullptr, *i8b = nullptr;
    std::uint32_t ld8 = 0;
    int bad_rows = 0;
};

// Makes (or finds cached) the buffers tile kernel `impl` reads for data set `ds`, centred at the feature means of `centre` (NULL: as given).
template <typename T>
operand<T> prepare_operand(plssvm_b200_ctx *ctx, plssvm_b200_dataset *ds, plssvm_b200_dataset *centre, const int impl, const bool need_sq) {
    operand<T> op;
    cudaStream_t st = ctx->stream;
    const std::uint64_t key = centre != nullptr ? centre->id : 0;
    const T *mean = centre != nullptr ? ensure_mean<T>(ctx, centre) : nullptr;
    const T *X = static_cast<const T *>(ds->X.p);
    if (key == 0) {
        op.sq = static_cast<const T *>(ds->sq.p);
    } else if (need_sq) {
        if (ds->sq_c.p == nullptr || ds->sq_c_key != key) {
            if (ds->sq_c.p != nullptr) { require_unpinned(ds, "centring it at other feature means"); }
            blk_alloc(ctx, ds->sq_c, ds->N * sizeof(T));
            pb::row_norms_kernel<T><<<static_cast<unsigned>((ds->N + 7) / 8), 256, 0, st>>>(X, ds->N, static_cast<std::uint32_t>(ds->ld), static_cast<T *>(ds->sq_c.p), mean);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            ds->sq_c_key = key;
        }
        op.sq = static_cast<const T *>(ds->sq_c.p);
    }
    if (is_i8(impl)) {
        const int slices = i8_slices_for<T>(impl), br_b = i8_br_b_for<T>(impl), slab = i8_slab_for<T>(impl);
        if (ds->X_i8.p == nullptr || ds->i8_slices != slices || ds->i8_br_b != br_b || ds->i8_key != key || ds->i8_slab != slab) {
            if (ds->X_i8.p != nullptr) { require_unpinned(ds, "splitting it into other digit planes"); }
            ds->ld8 = pitch_i8(ds->d);
            const std::size_t plane_bytes = static_cast<std::size_t>(slices) * rows_i8(ds->N) * ds->ld8;
            blk_alloc(ctx, ds->X_i8, plane_bytes);
            if (br_b != TILE) {
                blk_alloc(ctx, ds->X_i8b, plane_bytes);
            } else {
                blk_free(ctx, ds->X_i8b);
            }
            blk_alloc(ctx, ds->rscale, (ds->N + 2) * sizeof(T));
            int *bad_d = reinterpret_cast<int *>(static_cast<T *>(ds->rscale.p) + ds->N);  // scratch word behind the scales
            PB_CUDA(cudaMemsetAsync(bad_d, 0, sizeof(int), st));
            std::int8_t *pa = static_cast<std::int8_t *>(ds->X_i8.p);
            run_split_i8<T>(ctx, slices, X, ds->N, ds->d, ds->ld, pa, br_b != TILE ? static_cast<std::int8_t *>(ds->X_i8b.p) : pa, br_b, ds->ld8, static_cast<T *>(ds->rscale.p), bad_d, mean, slab, st);
            int *h = static_cast<int *>(ctx->pinned) + 512;  // second half of the pinned block (the first holds the CG state read-back)
            PB_CUDA(cudaMemcpyAsync(h, bad_d, sizeof(int), cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
            ds->i8_bad_rows = *h;
            ds->i8_slices = slices;
            ds->i8_br_b = br_b;
            ds->i8_slab = slab;
            ds->i8_key = key;
        }
        op.i8 = static_cast<const std::int8_t *>(ds->X_i8.p);
        op.i8b = br_b != TILE ? static_cast<const std::int8_t *>(ds->X_i8b.p) : op.i8;
        op.scale = static_cast<const T *>(ds->rscale.p);
        op.ld8 = static_cast<std::uint32_t>(ds->ld8);
        op.bad_rows = ds->i8_bad_rows;
        op.X = X;  // (not read by the int8-slice kernels)
        return op;
    }
    // tile kernels that read X itself: materialise the centred copy
    if (key != 0) {
        if (ds->Xc.p == nullptr || ds->xc_key != key) {
            if (ds->Xc.p != nullptr) { require_unpinned(ds, "centring it at other feature means"); }
            blk_alloc(ctx, ds->Xc, ds->N * ds->ld * sizeof(T));
            const std::size_t total = ds->N * ds->ld;
            const unsigned grid = static_cast<unsigned>(std::min<std::size_t>((total + 255) / 256, static_cast<std::size_t>(ctx->num_sms) * 32));
            pb::center_rows_kernel<T><<<grid, 256, 0, st>>>(X, static_cast<T *>(ds->Xc.p), ds->N, static_cast<std::uint32_t>(ds->ld), mean);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            ds->xc_key = key;
        }
        X = static_cast<const T *>(ds->Xc.p);
    }
    op.X = X;
    if constexpr (sizeof(T) == 4) {
        if (impl >= 2) {  // TF32 hi / lo split for the tcgen05 3xTF32 tiles (tile_tf32*.cuh)
            if (!ds->tf32_valid || ds->tf32_key != key) {
                if (ds->tf32_valid) { require_unpinned(ds, "re-splitting it for the 3xTF32 tiles"); }
                const std::size_t total = ds->N * ds->ld;
                blk_alloc(ctx, ds->X_hi, total * sizeof(float));
                blk_alloc(ctx, ds->X_lo, total * sizeof(float));
                const unsigned grid = static_cast<unsigned>(std::min<std::size_t>((total + 255) / 256, static_cast<std::size_t>(ctx->num_sms) * 32));
                pb::split_tf32_kernel<<<grid, 256, 0, st>>>(X, static_cast<float *>(ds->X_hi.p), static_cast<float *>(ds->X_lo.p), total);
                PB_CUDA(cudaGetLastError());
                ctx->tm.kernel_launches++;
                ds->tf32_valid = true;
                ds->tf32_key = key;
            }
            op.hi = static_cast<const float *>(ds->X_hi.p);
            op.lo = static_cast<const float *>(ds->X_lo.p);
        }
    }
    return op;
}

template <typename T, int KERNEL, int MODE>
void launch_tiles_t(plssvm_b200_ctx *ctx, const TileParams<T> &p, const int impl) {
    const std::uint64_t ntiles = p.tile_hi - p.tile_lo;
    if (ntiles == 0) { return; }
    const unsigned grid = static_cast<unsigned>(std::min<std::uint64_t>(ntiles, static_cast<std::uint64_t>(ctx->max_ctas > 0 ? std::min(ctx->max_ctas, ctx->num_sms) : ctx->num_sms)));
#ifdef PLSSVM_B200_EXPERIMENTAL
    if constexpr (sizeof(T) == 4) {
        if (impl == 9) {  // int8-slice tiles on CTA pairs (tcgen05.mma.cta_group::2): the operand boxes are contiguous -> 2-D boxes of 128-byte lines
            using L8 = pb::I8PairLayout<pb::I8<float>::S>;
            PB_REQUIRE(p.A_i8 != nullptr && p.B_i8 != nullptr && p.A_scale != nullptr && p.B_scale != nullptr, "int8-slice tensor path needs the digit planes of both operands");
            auto line_map = [&](CUtensorMap *tm, const std::int8_t *base, const std::size_t rows, const std::uint32_t box_lines) {
                const cuuint64_t dims[2] = { 128, static_cast<cuuint64_t>(static_cast<std::size_t>(L8::S) * rows_i8(rows) * p.ld8 / 128) };
                const cuuint64_t strides[1] = { 128 };
                const cuuint32_t box[2] = { 128, box_lines };
                const cuuint32_t estr[2] = { 1, 1 };
                const CUresult rc = ctx->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<std::int8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (rc != CUDA_SUCCESS) { throw api_error(PLSSVM_B200_ERR_CUDA, "cuTensorMapEncodeTiled (digit-plane lines) failed with code " + std::to_string(static_cast<int>(rc))); }
            };
            CUtensorMap tmA, tmB;
            line_map(&tmA, p.A_i8, static_cast<std::size_t>(p.T_rows) * TILE, L8::A_BYTES / 128);
            line_map(&tmB, p.B_i8b, static_cast<std::size_t>(p.T_cols) * TILE, L8::BH_BYTES / 128);
            const unsigned clusters = static_cast<unsigned>(std::min<std::uint64_t>(ntiles, static_cast<std::uint64_t>(ctx->num_sms / 2)));
            auto kern = pb::tile_kernel_i8_2sm<pb::I8<float>::S, KERNEL, MODE>;
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L8::SMEM_BYTES));
            kern<<<2 * clusters, pb::I8_THREADS, L8::SMEM_BYTES, ctx->stream>>>(tmA, tmB, p);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            return;
        }
    }
#endif
    if constexpr (pair_kernel_built<T>()) {
    if (impl == 10) {  // int8-slice tiles on CTA pairs (tile_i8_pair.cuh): contiguous operand boxes -> 2-D boxes of 128-byte lines
        using L8 = pb::I8PairLayout2<T, pb::I8<T>::S>;
        PB_REQUIRE(p.A_i8 != nullptr && p.B_i8 != nullptr && p.A_scale != nullptr && p.B_scale != nullptr, "int8-slice tensor path needs the digit planes of both operands");
        auto line_map = [&](CUtensorMap *tm, const std::int8_t *base, const std::size_t rows, const std::uint32_t box_lines) {
            const cuuint64_t dims[2] = { 128, static_cast<cuuint64_t>(static_cast<std::size_t>(L8::S) * rows_i8(rows) * p.ld8 / 128) };
            const cuuint64_t strides[1] = { 128 };
            const cuuint32_t box[2] = { 128, box_lines };
            const cuuint32_t estr[2] = { 1, 1 };
            const CUresult rc = ctx->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<std::int8_t *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) { throw api_error(PLSSVM_B200_ERR_CUDA, "cuTensorMapEncodeTiled (digit-plane lines) failed with code " + std::to_string(static_cast<int>(rc))); }
        };
        CUtensorMap tmA, tmR1, tmBig, tmSmall;
        const std::size_t rows_a = static_cast<std::size_t>(p.T_rows) * TILE, rows_b = static_cast<std::size_t>(p.T_cols) * TILE;
        line_map(&tmA, p.A_i8, rows_a, L8::A_BOX_LINES);
        line_map(&tmR1, p.B_i8, rows_b, L8::R1_BOX_LINES);
        line_map(&tmBig, p.B_i8, rows_b, L8::BIG_LINES);
        line_map(&tmSmall, p.B_i8, rows_b, L8::SMALL_LINES);
        const unsigned max_pairs = static_cast<unsigned>((ctx->max_ctas > 0 ? std::min(ctx->max_ctas, ctx->num_sms) : ctx->num_sms) / 2);
        const unsigned clusters = static_cast<unsigned>(std::min<std::uint64_t>(ntiles, std::max(1u, max_pairs)));
        auto kern = pb::tile_kernel_i8_pair<T, pb::I8<T>::S, KERNEL, MODE>;
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L8::SMEM_BYTES));
        kern<<<2 * clusters, pb::I8_THREADS, L8::SMEM_BYTES, ctx->stream>>>(tmA, tmR1, tmBig, tmSmall, p);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        return;
    }
    }
    if (is_i8(impl)) {  // int8-slice tcgen05 tiles: S exact int32 accumulators in TMEM (fp64: S = 7, units of 128 x 64; fp32: S = 3 or 4, units of 128 x 128)
        PB_REQUIRE(p.A_i8 != nullptr && p.B_i8 != nullptr && p.A_scale != nullptr && p.B_scale != nullptr, "int8-slice tensor path needs the digit planes of both operands");
        auto launch = [&](auto slices, auto cluster) {
            constexpr int S = decltype(slices)::value, CL = decltype(cluster)::value;
            using L8 = pb::I8Layout<T, S>;
            auto kern = pb::tile_kernel_i8<T, S, KERNEL, MODE, CL>;
            if constexpr (sizeof(T) == 8 && S == 7 && CL == 1) {
                if (ctx->i8_a_via_tmem != 0) { kern = pb::tile_kernel_i8<T, S, KERNEL, MODE, CL, true>; }  // doubly-read A planes through tensor memory
            }
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L8::SMEM_BYTES));
            if constexpr (CL == 1) {
                kern<<<grid, pb::I8_THREADS, L8::SMEM_BYTES, ctx->stream>>>(p);
            } else {
                cudaLaunchConfig_t cfg{};
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = CL;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.blockDim = dim3(pb::I8_THREADS);
                cfg.dynamicSmemBytes = L8::SMEM_BYTES;
                cfg.stream = ctx->stream;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                cfg.gridDim = dim3(static_cast<unsigned>(ctx->num_sms / CL * CL));
                int max_clusters = 0;  // clusters that can be resident at once (GPC boundaries can leave a few SMs unused)
                PB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
                PB_REQUIRE(max_clusters > 0, "no CTA cluster of the int8-slice kernel fits on this device");
                const unsigned clusters = static_cast<unsigned>(std::min<std::uint64_t>(ntiles, static_cast<std::uint64_t>(max_clusters)));
                cfg.gridDim = dim3(CL * clusters);
                PB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
            }
        };
        const bool dflt = i8_slices_for<T>(impl) == pb::I8<T>::S;
        if (EXPERIMENTAL && impl == 8) {
#ifdef PLSSVM_B200_EXPERIMENTAL
            launch(std::integral_constant<int, pb::I8<T>::S>{}, std::integral_constant<int, 4>{});
#endif
        } else if (EXPERIMENTAL && impl == 11) {
#ifdef PLSSVM_B200_EXPERIMENTAL
            launch(std::integral_constant<int, pb::I8<T>::S>{}, std::integral_constant<int, 2>{});
#endif
        } else if (dflt) {
            launch(std::integral_constant<int, pb::I8<T>::S>{}, std::integral_constant<int, 1>{});
        } else {
            launch(std::integral_constant<int, pb::I8<T>::S_EXACT>{}, std::integral_constant<int, 1>{});
        }
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        return;
    }
    if constexpr (sizeof(T) == 8) {
        if (impl == 2) {
            CUtensorMap tmA, tmB;
            make_tensor_map<double>(ctx, &tmA, p.A, p.n_rows, p.ld);
            make_tensor_map<double>(ctx, &tmB, p.B, p.n_cols, p.ld);
            PB_CUDA(cudaFuncSetAttribute(pb::tile_kernel_dmma<KERNEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::DMMA_SMEM_BYTES));
            pb::tile_kernel_dmma<KERNEL, MODE><<<grid, pb::DMMA_THREADS, pb::DMMA_SMEM_BYTES, ctx->stream>>>(tmA, tmB, p);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            return;
        }
    } else {
        CUtensorMap tmAhi, tmAlo, tmBhi, tmBlo;  // TF32 hi / lo operands of the tcgen05 kernels (impl 2, 4, 5)
        if (impl >= 2) {
            PB_REQUIRE(p.A_hi != nullptr && p.A_lo != nullptr && p.B_hi != nullptr && p.B_lo != nullptr, "fp32 tensor path needs the hi / lo split of both operands");
            make_tensor_map<float>(ctx, &tmAhi, p.A_hi, p.n_rows, p.ld);
            make_tensor_map<float>(ctx, &tmAlo, p.A_lo, p.n_rows, p.ld);
            make_tensor_map<float>(ctx, &tmBhi, p.B_hi, p.n_cols, p.ld);
            make_tensor_map<float>(ctx, &tmBlo, p.B_lo, p.n_cols, p.ld);
        }
#ifdef PLSSVM_B200_EXPERIMENTAL
        if (impl == 4) {  // CTA-pair tcgen05 kernel: tile range is in 256 x 256 super-tiles, one cluster of two CTAs per SM pair
            const unsigned clusters = static_cast<unsigned>(std::min<std::uint64_t>(ntiles, static_cast<std::uint64_t>(ctx->num_sms / 2)));
            PB_CUDA(cudaFuncSetAttribute(pb::tile_kernel_tf32_2sm<KERNEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::TF2_SMEM_BYTES));
            pb::tile_kernel_tf32_2sm<KERNEL, MODE><<<2 * clusters, pb::TF2_THREADS, pb::TF2_SMEM_BYTES, ctx->stream>>>(tmAhi, tmAlo, tmBhi, tmBlo, p);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            return;
        }
        if (impl == 5) {  // 128 x 256 tiles per CTA: units are halves of the 256 x 256 super-tiles (tile range in super-tiles)
            const unsigned g5 = static_cast<unsigned>(std::min<std::uint64_t>(2 * ntiles, static_cast<std::uint64_t>(ctx->num_sms)));
            PB_CUDA(cudaFuncSetAttribute(pb::tile_kernel_tf32_n256<KERNEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::TN_SMEM_BYTES));
            pb::tile_kernel_tf32_n256<KERNEL, MODE><<<g5, pb::TN_THREADS, pb::TN_SMEM_BYTES, ctx->stream>>>(tmAhi, tmAlo, tmBhi, tmBlo, p);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            return;
        }
#endif
        if (impl == 2) {
            PB_CUDA(cudaFuncSetAttribute(pb::tile_kernel_tf32<KERNEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb::TF32_SMEM_BYTES));
            pb::tile_kernel_tf32<KERNEL, MODE><<<grid, pb::TF32_THREADS, pb::TF32_SMEM_BYTES, ctx->stream>>>(tmAhi, tmAlo, tmBhi, tmBlo, p);
            PB_CUDA(cudaGetLastError());
            ctx->tm.kernel_launches++;
            return;
        }
    }
    pb::tile_kernel_simt<T, KERNEL, MODE><<<grid, 256, 0, ctx->stream>>>(p);
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches++;
}

template <typename T>
int resolve_impl(const plssvm_b200_ctx *ctx, const std::size_t features = 0, const std::size_t rows = 0) {
    // int8-slice tcgen05 tiles: beyond I8_MAX_FEATURES the int32 accumulators could overflow -> DMMA / 3xTF32 tiles
    if (is_i8(ctx->impl)) {
        if (features > pb::I8_MAX_FEATURES) { return 2; }
        if (sizeof(T) == 8 && (ctx->impl == 7 || ctx->impl == 9)) { return 6; }
        if (ctx->impl == 10 && !pair_kernel_built<T>()) { return 6; }
        if (ctx->impl == 11 && !EXPERIMENTAL) { return 6; }
        return ctx->impl;
    }
    if (ctx->impl == 4 || ctx->impl == 5) { return sizeof(T) == 4 ? ctx->impl : 2; }  // CTA-pair / wide-tile tcgen05 kernels exist for fp32 only
    if (ctx->impl != 0) { return ctx->impl; }
    // auto: int8 slices on tcgen05 where the int32 accumulators cannot overflow (fp32 with at least one 256-row block of the A operand: on CTA
    // pairs); callers fall back to 2 for badly scaled rows: fp64 -> TMA + DMMA (tile_dmma.cuh), fp32 -> TMA + tcgen05 3xTF32 + TMEM (tile_tf32.cuh)
    if (features > 0 && features <= pb::I8_MAX_FEATURES) { return (sizeof(T) == 4 && ctx->fp32_pair != 0 && rows >= 2 * TILE && ctx->pairs_ok != 0) ? 10 : 6; }
    return 2;
}
// automatic kernel choice only: the int8-slice tiles are used unless an operand holds badly scaled rows (split_i8_kernel)
inline bool i8_forced(const plssvm_b200_ctx *ctx) { return is_i8(ctx->impl); }

template <typename T, int MODE>
void launch_tiles(plssvm_b200_ctx *ctx, const TileParams<T> &p_in, const int impl) {
    ctx->tm.impl_used = impl;
    TileParams<T> p = p_in;
    p.slow_drain = ctx->fp32_fast_drain != 0 ? 0 : 1;
    const bool stats = ctx->tile_stats != 0 && (impl == 6 || impl == 7 || impl == 10 || impl == 11);
    const std::size_t stat_words = static_cast<std::size_t>(ctx->num_sms) * 8;
    if (stats) {
        p.stats = workspace<unsigned long long>(ctx, plssvm_b200_ctx::WS_STATS, stat_words);
        PB_CUDA(cudaMemsetAsync(p.stats, 0, stat_words * sizeof(unsigned long long), ctx->stream));
    }
    switch (p.kp.kernel) {
        case pb::K_LINEAR: launch_tiles_t<T, pb::K_LINEAR, MODE>(ctx, p, impl); break;
        case pb::K_POLYNOMIAL: launch_tiles_t<T, pb::K_POLYNOMIAL, MODE>(ctx, p, impl); break;
        default: launch_tiles_t<T, pb::K_RBF, MODE>(ctx, p, impl); break;
    }
    if (stats) {  // profiling only: read the per-CTA role counters of this launch
        std::vector<unsigned long long> h(stat_words);
        PB_CUDA(cudaMemcpyAsync(h.data(), p.stats, stat_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        double f[7] = { 0, 0, 0, 0, 0, 0, 0 };
        int ctas = 0;
        for (int b = 0; b < ctx->num_sms; ++b) {
            const double total = static_cast<double>(h[b * 8]);
            if (total <= 0.0) { continue; }
            ++ctas;
            for (int k = 0; k < 7; ++k) { f[k] += static_cast<double>(h[b * 8 + 1 + k]) / total; }
        }
#ifdef PB_TILE_STATS_FINE
        if (ctas > 0 && ctx->tile_stats >= 2) {  // the epilogue's own split (one thread): wait / drain / vector loads + barriers + stores; the remainder is the kernel function + sums
            std::fprintf(stderr, "{\"tile_stats\": {\"impl\": %d, \"ctas\": %d, \"mma_wait_operands\": %.4f, \"mma_wait_drain\": %.4f, \"producer_wait\": %.4f, "
                                 "\"epilogue_wait_accumulators\": %.4f, \"epilogue_drain\": %.4f, \"epilogue_loads_barriers_stores\": %.4f, \"epilogue_math\": %.4f}}\n",
                         impl, ctas, f[0] / ctas, f[1] / ctas, f[2] / ctas, f[3] / ctas, f[4] / ctas, f[6] / ctas, 1.0 - (f[3] + f[4] + f[6]) / ctas);
        }
#endif
        if (ctas > 0) {
            ctx->tm.tile_mma_wait_operands = f[0] / ctas;
            ctx->tm.tile_mma_wait_drain = f[1] / ctas;
            ctx->tm.tile_producer_wait = f[2] / ctas;
            ctx->tm.tile_epilogue_wait = f[3] / ctas;
        }
    }
}

template <typename T>
constexpr int nccl_type() { return sizeof(T) == 8 ? NCCL_FLOAT64 : NCCL_FLOAT32; }

template <typename T>
void all_reduce_sum(plssvm_b200_ctx *ctx, T *buf, const std::size_t count) {
    if (ctx->world <= 1 || ctx->comm == nullptr) { return; }  // (no communicator: a virtual rank, see option "virtual_world")
    const nccl_api &nccl = nccl_api::get();
    nccl.check(nccl.AllReduce(buf, buf, count, nccl_type<T>(), NCCL_SUM, ctx->comm, ctx->stream), "ncclAllReduce");
}

void validate_kernel_args(const int kernel, const double gamma) {
    PB_REQUIRE(kernel >= 0 && kernel <= 2, "unknown kernel function type " + std::to_string(kernel));
    if (kernel != pb::K_LINEAR) { PB_REQUIRE(gamma > 0.0, "gamma must be greater than 0, but is " + std::to_string(gamma) + "!"); }
}

template <typename T>
void set_operands(TileParams<T> &p, const operand<T> &a, const operand<T> &b) {
    p.A = a.X;
    p.B = b.X;
    p.A_hi = a.hi;
    p.A_lo = a.lo;
    p.B_hi = b.hi;
    p.B_lo = b.lo;
    p.A_i8 = a.i8;
    p.B_i8 = b.i8;
    p.B_i8b = b.i8b;
    p.A_scale = a.scale;
    p.B_scale = b.scale;
    p.ld8 = b.ld8 != 0 ? b.ld8 : a.ld8;
    p.row_sq = a.sq;
    p.col_sq = b.sq;
}

// ---- implicit matvec: out = Q~ v  (set semantics; callers add / subtract) ----------------------------------------------------
template <typename T>
struct matvec_plan {
    plssvm_b200_ctx *ctx;
    plssvm_b200_dataset *ds;
    std::uint32_t n;   // N - 1
    std::uint32_t Tb;  // tiles per side
    int tile_shift = 0;
    int impl = 2;
    std::uint64_t total_tiles = 0, tile_lo = 0, tile_hi = 0;
    std::vector<double> weights;  // rate-weighted shares of the tile order, one per rank (all 1 = equal shares)
    dbuf<T> partial;
    TileParams<T> base;
    bool pinned = false;

    matvec_plan(plssvm_b200_ctx *c, plssvm_b200_dataset *data, const KernelParams<T> &kp, const T *q, const T *QA_cost_dev, const T cost_inv, const int *done) :
        ctx(c), ds(data) {
        n = static_cast<std::uint32_t>(data->N - 1);
        Tb = (n + TILE - 1) / TILE;
        const bool tiles_needed = !(c->linear_factorized != 0 && kp.kernel == pb::K_LINEAR);
        plssvm_b200_dataset *centre = kp.kernel == pb::K_RBF ? data : nullptr;  // rbf: on data centred at its own feature means
        impl = resolve_impl<T>(c, data->ld, data->N - 1);
        operand<T> op;
        if (tiles_needed) {
            op = prepare_operand<T>(c, data, centre, impl, kp.kernel == pb::K_RBF);
            if (is_i8(impl) && !i8_forced(c) && op.bad_rows != 0) {
                impl = 2;
                op = prepare_operand<T>(c, data, centre, impl, kp.kernel == pb::K_RBF);
            }
        } else {
            op.X = static_cast<const T *>(data->X.p);
        }
        tile_shift = super_tiled(impl) ? 1 : 0;  // CTA-pair kernel: the schedule (and rank ownership) is over 256 x 256 super-tiles
        total_tiles = pb::tri_num_tiles((Tb + tile_shift) >> tile_shift);
        weights.assign(static_cast<std::size_t>(c->world), 1.0);
        if (c->comm == nullptr && c->world > 1 && c->virtual_skew != 0) {  // virtual ranks: deliberately unequal shares, to test the rate-weighted cut
            for (int g = 0; g < c->world; ++g) { weights[static_cast<std::size_t>(g)] = 1.0 + 0.01 * c->virtual_skew * (static_cast<double>(g) / (c->world - 1) - 0.5); }
        }
        pb::weighted_range(total_tiles, c->rank, c->world, weights.data(), tile_lo, tile_hi);
        if (tiles_needed) { partial.alloc(c, static_cast<std::size_t>(Tb) * Tb * TILE); }
        base = TileParams<T>{};
        set_operands(base, op, op);
        base.n_rows = n;
        base.n_cols = n;
        base.ld = static_cast<std::uint32_t>(data->ld);
        base.T_rows = Tb;
        base.T_cols = Tb;
        base.q = q;
        base.QA_cost = QA_cost_dev;
        base.cost_inv = cost_inv;
        base.kp = kp;
        base.partial = partial.p;
        base.done = done;
        ds->pins++;
        pinned = true;
    }
    ~matvec_plan() {
        if (pinned) { ds->pins--; }
    }
    matvec_plan(const matvec_plan &) = delete;
    matvec_plan &operator=(const matvec_plan &) = delete;

    // new shares of the tile order (identical on every rank: the weights come out of an all-gather)
    void set_weights(const std::vector<double> &w) {
        weights = w;
        pb::weighted_range(total_tiles, ctx->rank, ctx->world, weights.data(), tile_lo, tile_hi);
    }

    // linear kernel, factorised: out = X (X^T v) + (QA_cost - q) S - q.v + v / C   — identical on every rank, no collective
    dbuf<T> fact_w, fact_part, fact_sums;
    void run_factorized(const T *v, T *out) {
        cudaStream_t st = ctx->stream;
        const std::uint32_t d = static_cast<std::uint32_t>(ds->d), ld = static_cast<std::uint32_t>(ds->ld);
        const std::uint32_t chunks = (n + pb::W_ROWS - 1) / pb::W_ROWS;
        if (fact_w.count == 0) {
            fact_w.alloc(ctx, ld);
            fact_part.alloc(ctx, static_cast<std::size_t>(chunks) * d);
            fact_sums.alloc(ctx, 2 * static_cast<std::size_t>((n + pb::VEC_CHUNK - 1) / pb::VEC_CHUNK));
            PB_CUDA(cudaMemsetAsync(fact_w.p, 0, ld * sizeof(T), st));
        }
        ctx->matvec_timer.begin(st);
        pb::w_partial_kernel<T><<<dim3((d + 255) / 256, chunks), 256, 0, st>>>(base.A, v, n, d, ld, fact_part.p);
        const std::uint32_t vb = (n + pb::VEC_CHUNK - 1) / pb::VEC_CHUNK;
        pb::w_reduce_kernel<T><<<(d + 31) / 32, 256, 0, st>>>(fact_part.p, chunks, d, fact_w.p);
        pb::linear_fact_sums_kernel<T><<<vb, pb::VEC_BLOCK, 0, st>>>(v, base.q, n, fact_sums.p, base.done);
        pb::linear_fact_apply_kernel<T><<<(n + 7) / 8, 256, 0, st>>>(base.A, n, ld, fact_w.p, base.q, v, fact_sums.p, vb, base.QA_cost, base.cost_inv, out, base.done);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches += 4;
        ctx->matvec_timer.end(st);
        ctx->tm.matvec_calls++;
        ctx->tm.impl_used = 3;
    }

    // out = Q~ v
    void run(const T *v, T *out) {
        if (ctx->linear_factorized != 0 && base.kp.kernel == pb::K_LINEAR) {
            run_factorized(v, out);
            return;
        }
        TileParams<T> p = base;
        p.v = v;
        p.tile_lo = tile_lo;
        p.tile_hi = tile_hi;
        ctx->matvec_timer.begin(ctx->stream);
        ctx->tile_timer.begin(ctx->stream);
        launch_tiles<T, pb::MODE_SYM>(ctx, p, impl);
        ctx->tile_timer.end(ctx->stream);
        pb::reduce_partials_kernel<T, pb::MODE_SYM><<<Tb, 512, 0, ctx->stream>>>(partial.p, out, n, Tb, Tb, tile_lo, tile_hi, ctx->world > 1 ? 1 : 0, tile_shift, T(1), T(0), 0, base.done);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        all_reduce_sum(ctx, out, n);
        ctx->matvec_timer.end(ctx->stream);
        ctx->tm.matvec_calls++;
    }
};

template <typename T, typename F>
void dispatch_kernel(const int kernel, F &&f) {
    switch (kernel) {
        case pb::K_LINEAR: f(std::integral_constant<int, pb::K_LINEAR>{}); break;
        case pb::K_POLYNOMIAL: f(std::integral_constant<int, pb::K_POLYNOMIAL>{}); break;
        default: f(std::integral_constant<int, pb::K_RBF>{}); break;
    }
}

template <typename T>
void run_q_kernel(plssvm_b200_ctx *ctx, const plssvm_b200_dataset *ds, const KernelParams<T> &kp, T *q_full /* N */) {
    const unsigned grid = static_cast<unsigned>((ds->N + 7) / 8);
    dispatch_kernel<T>(kp.kernel, [&](auto K) {
        pb::q_kernel<T, decltype(K)::value><<<grid, 256, 0, ctx->stream>>>(static_cast<const T *>(ds->X.p), ds->N, static_cast<std::uint32_t>(ds->ld), kp, q_full);
    });
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches++;
}

// timings describe the last call; while a CG session is open they keep accumulating over its lifetime
void reset_timings(plssvm_b200_ctx *ctx) {
    if (ctx->open_sessions > 0) { return; }
    ctx->tm = plssvm_b200_timings{};
    ctx->tm.n_devices = ctx->world;
    ctx->tile_timer.reset();
    ctx->matvec_timer.reset();
}

void check_dataset(const plssvm_b200_ctx *ctx, const plssvm_b200_dataset *ds, const std::size_t elem, const char *what) {
    PB_REQUIRE(ctx != nullptr, "context is NULL");
    PB_REQUIRE(ds != nullptr, std::string(what) + " dataset is NULL");
    PB_REQUIRE(ds->ctx == ctx, std::string(what) + " dataset belongs to another context");
    PB_REQUIRE(ds->elem_size == static_cast<int>(elem), std::string(what) + " dataset has the wrong real_type");
}

// ---- dataset ----------------------------------------------------------------------------------------------------------------
void dataset_free_buffers(plssvm_b200_dataset *ds) {
    plssvm_b200_ctx *ctx = ds->ctx;
    for (dblock *b : { &ds->X, &ds->sq, &ds->mean, &ds->sq_c, &ds->Xc, &ds->X_hi, &ds->X_lo, &ds->X_i8, &ds->X_i8b, &ds->rscale }) { blk_free(ctx, *b); }
}

// One rank's part of creating a data set.  Host source with several ranks: this rank uploads rows [rank, rank + 1) * ceil(N / world) over its own
// PCIe link and the row shares are all-gathered over NVLink (every rank of a multi-process run passes the same matrix).  Device source: packed
// on the device that holds it; in a device group rank 0 holds the source and broadcasts the packed copy.
template <typename T>
plssvm_b200_dataset *dataset_create_rank(plssvm_b200_ctx *ctx, const host_matrix<T> &host, const T *dev_src, const std::size_t N, const std::size_t d) {
    PB_REQUIRE(ctx != nullptr, "context is NULL");
    PB_REQUIRE(host.valid() || dev_src != nullptr, "The data must not be empty!");
    PB_REQUIRE(N > 0, "The data must not be empty!");
    PB_REQUIRE(d > 0, "The data points must contain at least one feature!");
    PB_REQUIRE(N < (1ull << 31) && d < (1ull << 31), "matrix dimensions must be below 2^31");
    PB_CUDA(cudaSetDevice(ctx->device));
    auto *ds = new plssvm_b200_dataset{};
    try {
        cudaStream_t st = ctx->stream;
        ds->ctx = ctx;
        ds->elem_size = static_cast<int>(sizeof(T));
        ds->N = N;
        ds->d = d;
        ds->ld = pitch_elems<T>(d);
        ds->id = g_next_dataset_id.fetch_add(1);
        const std::size_t G = static_cast<std::size_t>(ctx->world);
        const bool sharded = G > 1 && ctx->comm != nullptr && host.valid() && ctx->shard_upload != 0;
        const std::size_t share = sharded ? (N + G - 1) / G : N;  // rows per rank (the last share may be short; X is allocated for G full shares)
        const std::size_t rows_alloc = sharded ? share * G : N;
        {   // allocate, then make sure every device of a group got its memory before anyone enters the all-gather / broadcast
            std::exception_ptr err;
            try {
                blk_alloc(ctx, ds->X, rows_alloc * ds->ld * sizeof(T));
                blk_alloc(ctx, ds->sq, N * sizeof(T));
            } catch (...) {
                err = std::current_exception();
            }
            const bool all_ok = group_all_ok(ctx, !err);
            if (err) { std::rethrow_exception(err); }
            if (!all_ok) { throw api_error(PLSSVM_B200_ERR_CUDA, "another device of the group failed to allocate the data set"); }
        }
        T *X = static_cast<T *>(ds->X.p);
        const nccl_api *nccl = (G > 1 && ctx->comm != nullptr) ? &nccl_api::get() : nullptr;
        if (dev_src != nullptr) {
            if (!ctx->in_process_group() || ctx->rank == 0) {
                const std::size_t total = N * ds->ld;
                const unsigned grid = static_cast<unsigned>(std::min<std::size_t>((total + 255) / 256, static_cast<std::size_t>(ctx->num_sms) * 32));
                pb::pack_rows_kernel<T><<<grid, 256, 0, st>>>(dev_src, X, N, static_cast<std::uint32_t>(d), static_cast<std::uint32_t>(ds->ld));
                PB_CUDA(cudaGetLastError());
                ctx->tm.kernel_launches++;
            }
            if (ctx->in_process_group()) { nccl->check(nccl->Broadcast(X, X, N * ds->ld * sizeof(T), NCCL_INT8, 0, ctx->comm, st), "ncclBroadcast"); }
        } else {
            const std::size_t r0 = sharded ? std::min(N, share * static_cast<std::size_t>(ctx->rank)) : 0, r1 = sharded ? std::min(N, r0 + share) : N;
            if (ds->ld != d && r1 > r0) { PB_CUDA(cudaMemsetAsync(X + r0 * ds->ld, 0, (r1 - r0) * ds->ld * sizeof(T), st)); }
            upload_rows<T>(ctx, X + r0 * ds->ld, ds->ld, host, r0, r1, st);
            ctx->tm.h2d_bytes += static_cast<double>((r1 - r0) * d * sizeof(T));
            if (sharded) {
                nccl->check(nccl->AllGather(X + share * static_cast<std::size_t>(ctx->rank) * ds->ld, X, share * ds->ld * sizeof(T), NCCL_INT8, ctx->comm, st), "ncclAllGather");
            }
        }
        pb::row_norms_kernel<T><<<static_cast<unsigned>((N + 7) / 8), 256, 0, st>>>(X, N, static_cast<std::uint32_t>(ds->ld), static_cast<T *>(ds->sq.p), nullptr);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        PB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        dataset_free_buffers(ds);
        delete ds;
        throw;
    }
    return ds;
}

void dataset_destroy_rank(plssvm_b200_dataset *ds) {
    if (ds == nullptr) { return; }
    cudaSetDevice(ds->ctx->device);
    dataset_free_buffers(ds);
    delete ds;
}

// creates the data set on every device of the context; returns the handle (member 0)
template <typename T>
plssvm_b200_dataset *dataset_create(plssvm_b200_ctx *ctx, const host_matrix<T> &host, const T *dev_src, const std::size_t N, const std::size_t d) {
    PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");
    const std::size_t G = std::max<std::size_t>(1, ctx->members.size());
    std::vector<plssvm_b200_dataset *> parts(G, nullptr);
    try {
        for_each_rank(ctx, [&](plssvm_b200_ctx *c, const int g) { parts[g] = dataset_create_rank<T>(c, host, dev_src, N, d); });
    } catch (...) {
        for (auto *p : parts) { dataset_destroy_rank(p); }
        throw;
    }
    if (G > 1) { parts[0]->members = parts; }
    ctx->live_datasets.push_back(parts[0]);
    handle_add(parts[0]);
    return parts[0];
}

void dataset_destroy(plssvm_b200_dataset *ds) {
    if (!handle_remove(ds)) { return; }  // NULL, or already destroyed together with its context
    auto &live = ds->ctx->live_datasets;
    live.erase(std::remove(live.begin(), live.end(), ds), live.end());
    const std::vector<plssvm_b200_dataset *> members = ds->members;
    if (members.size() > 1) {
        for (auto *m : members) { dataset_destroy_rank(m); }
    } else {
        dataset_destroy_rank(ds);
    }
}

plssvm_b200_dataset *member_of(plssvm_b200_dataset *ds, const int g) { return ds->members.size() > 1 ? ds->members[static_cast<std::size_t>(g)] : ds; }

void check_group_dataset(const plssvm_b200_ctx *ctx, const plssvm_b200_dataset *ds, const std::size_t elem, const char *what) {
    PB_REQUIRE(handle_live(ctx), "the context is NULL or has been destroyed");
    PB_REQUIRE(handle_live(ds), std::string(what) + " dataset is NULL or has been destroyed (data sets die with their context)");
    check_dataset(ctx, ds, elem, what);
    PB_REQUIRE(ds->members.size() == ctx->members.size() || (ds->members.empty() && ctx->members.size() <= 1), std::string(what) + " dataset was not created by this context");
}

// ---- solve: csvm::solve_system_of_linear_equations (gpu_csvm.hpp:477-654) ----------------------------------------------------
// One CG solve as a session: begin (b~, x0 = 1, q, QA_cost, r0 = b~ - Q~ x0, d0 = r0), step (k iterations enqueued back to
// back, then one poll of the device-side state), finish (bias, alpha_N, download).  plssvm_b200_solve_* is begin + step
// until converged / max_iter + finish; the benchmark drives step() directly so that exactly K iterations are timed.
struct cg_session_base {
    plssvm_b200_ctx *ctx = nullptr;
    int elem_size = 0;
    virtual ~cg_session_base() = default;
};

template <typename T>
struct cg_session : cg_session_base {
    plssvm_b200_dataset *ds;
    KernelParams<T> kp;
    T cost, eps;
    std::uint32_t n;
    unsigned vblocks;
    dbuf<T> y_d, q_full, b, x, r, dvec, Ad, part, trace;  // trace[k] = r.r after k iterations (k <= TRACE_CAP)
    static constexpr std::uint64_t TRACE_CAP = 4096;
    dbuf<CGState<T>> state;
    dbuf<double> rates_d;  // one measured tile rate per rank (rate-weighted shares)
    std::unique_ptr<matvec_plan<T>> mv;
    std::uint64_t iters_enqueued = 0, since_balance = 0;
    bool converged = false, counted = false;
    CGState<T> last{};  // last polled copy of the device state
    double loop_wall_ms = 0.0;

    cg_session(plssvm_b200_ctx *c, plssvm_b200_dataset *data, const T *y, const int kernel, const int degree, const T gamma, const T coef0, const T cost_, const T eps_) :
        ds(data), kp{ kernel, degree, gamma, coef0 }, cost(cost_), eps(eps_) {
        ctx = c;
        elem_size = static_cast<int>(sizeof(T));
        check_dataset(ctx, ds, sizeof(T), "training");
        PB_REQUIRE(y != nullptr, "y must not be NULL");
        PB_REQUIRE(ds->N >= 2, "The data must contain at least two data points!");
        PB_REQUIRE(eps > T(0), "The stopping criterion in the CG algorithm must be greater than 0.0, but is " + std::to_string(eps) + "!");
        PB_REQUIRE(cost != T(0), "cost must not be 0.0!");
        validate_kernel_args(kernel, static_cast<double>(gamma));
        PB_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const std::size_t N = ds->N;
        n = static_cast<std::uint32_t>(N - 1);
        vblocks = (n + pb::VEC_CHUNK - 1) / pb::VEC_CHUNK;
        y_d.alloc(c, N);
        q_full.alloc(c, N);
        b.alloc(c, n);
        x.alloc(c, n);
        r.alloc(c, n);
        dvec.alloc(c, n);
        Ad.alloc(c, n);
        part.alloc(c, vblocks);
        trace.alloc(c, TRACE_CAP + 1);
        state.alloc(c, 1);
        if (ctx->world > 1) { rates_d.alloc(c, static_cast<std::size_t>(ctx->world)); }

        PB_CUDA(cudaMemcpyAsync(y_d.p, y, N * sizeof(T), cudaMemcpyHostToDevice, st));
        ctx->tm.h2d_bytes += static_cast<double>(N * sizeof(T));
        PB_CUDA(cudaMemsetAsync(state.p, 0, sizeof(CGState<T>), st));
        pb::cg_init_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(y_d.p, n, b.p, x.p, state.p, cost);
        run_q_kernel<T>(ctx, ds, kp, q_full.p);
        pb::cg_qa_cost_kernel<T><<<1, 1, 0, st>>>(q_full.p, n, state.p, cost);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches += 2;

        {   // operands and the partial buffer are allocated here: every device of a group must have succeeded before the first all-reduce
            std::exception_ptr err;
            try {
                mv = std::make_unique<matvec_plan<T>>(ctx, ds, kp, q_full.p, &state.p->QA_cost, T(1) / cost, &state.p->done);
            } catch (...) {
                err = std::current_exception();
            }
            const bool all_ok = group_all_ok(ctx, !err);
            if (err) { std::rethrow_exception(err); }
            if (!all_ok) { throw api_error(PLSSVM_B200_ERR_CUDA, "another device of the group failed to set up the CG solve"); }
        }
        ctx->tm.matvec_flops = static_cast<double>(ds->d) * static_cast<double>(n) * (static_cast<double>(n) + 1.0);
        ctx->tm.cg_epsilon = static_cast<double>(eps);

        // r = b - Q~ x0,  delta0 = r.r,  d = r     (gpu_csvm.hpp:515-554)
        mv->run(x.p, Ad.p);
        pb::cg_residual_kernel<T><<<vblocks, pb::VEC_BLOCK, 0, st>>>(b.p, Ad.p, r.p, n, part.p, nullptr);
        pb::cg_start_kernel<T><<<1, pb::VEC_BLOCK, 0, st>>>(part.p, vblocks, state.p, trace.p);
        pb::cg_update_d_kernel<T, true><<<vblocks, pb::VEC_BLOCK, 0, st>>>(dvec.p, r.p, n, state.p);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches += 3;
        PB_CUDA(cudaStreamSynchronize(st));  // (the event timers below assume that everything recorded before an iteration chunk has completed)
        ctx->open_sessions++;
        counted = true;
    }
    ~cg_session() override {
        if (counted) { ctx->open_sessions--; }
    }

    void enqueue_iteration(const std::uint64_t iter) {
        cudaStream_t st = ctx->stream;
        const int *done = &state.p->done;
        mv->run(dvec.p, Ad.p);                                                                                       // Ad = Q~ d        (574-582)
        pb::dot_partial_kernel<T><<<vblocks, pb::VEC_BLOCK, 0, st>>>(dvec.p, Ad.p, n, part.p, done);
        pb::cg_alpha_kernel<T><<<1, pb::VEC_BLOCK, 0, st>>>(part.p, vblocks, state.p);                               // alpha = delta / d.Ad (585)
        if (iter % 50 == 49) {                                                                                      // residual refresh (595-609)
            pb::cg_update_xr_kernel<T, true><<<vblocks, pb::VEC_BLOCK, 0, st>>>(x.p, r.p, dvec.p, Ad.p, n, state.p, part.p);
            mv->run(x.p, Ad.p);
            pb::cg_residual_kernel<T><<<vblocks, pb::VEC_BLOCK, 0, st>>>(b.p, Ad.p, r.p, n, part.p, done);
        } else {
            pb::cg_update_xr_kernel<T, false><<<vblocks, pb::VEC_BLOCK, 0, st>>>(x.p, r.p, dvec.p, Ad.p, n, state.p, part.p);  // x += a d; r -= a Ad (588, 611-613)
        }
        pb::cg_beta_kernel<T><<<1, pb::VEC_BLOCK, 0, st>>>(part.p, vblocks, state.p, eps, iter < TRACE_CAP ? trace.p : nullptr, ctx->ignore_convergence);                 // delta, stop test, beta (616-625)
        pb::cg_update_d_kernel<T, false><<<vblocks, pb::VEC_BLOCK, 0, st>>>(dvec.p, r.p, n, state.p);                // d = beta d + r   (627)
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches += (iter % 50 == 49) ? 6 : 5;
    }

    bool balancing() const { return ctx->world > 1 && ctx->comm != nullptr && ctx->balance != 0 && !(ctx->linear_factorized != 0 && kp.kernel == pb::K_LINEAR); }

    // Rate-weighted tile shares: every rank contributes the rate (tiles per ms) its tile kernels ran at since the last re-cut; the all-gathered
    // rates — identical on every rank — become the new weights of the contiguous shares of the tile order.  Ownership is recomputed from the
    // shares by the reduction kernel, so nothing else changes.  GPUs under the 1 kW cap do not hold the same clock: with equal shares every
    // matvec waits in the all-reduce for the slowest rank.
    void rebalance(const double tile_ms, const std::uint64_t matvecs) {
        cudaStream_t st = ctx->stream;
        const std::size_t G = static_cast<std::size_t>(ctx->world);
        double *h = reinterpret_cast<double *>(static_cast<char *>(ctx->pinned) + 1024);  // G <= 64 doubles
        const double tiles = static_cast<double>(mv->tile_hi - mv->tile_lo) * static_cast<double>(matvecs);
        h[0] = (tile_ms > 0.0 && tiles > 0.0) ? tiles / tile_ms : 0.0;
        PB_CUDA(cudaMemcpyAsync(rates_d.p + ctx->rank, h, sizeof(double), cudaMemcpyHostToDevice, st));
        const nccl_api &nccl = nccl_api::get();
        nccl.check(nccl.AllGather(rates_d.p + ctx->rank, rates_d.p, 1, NCCL_FLOAT64, ctx->comm, st), "ncclAllGather");
        PB_CUDA(cudaMemcpyAsync(h, rates_d.p, G * sizeof(double), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        std::vector<double> w(G);
        double mean_rate = 0.0;
        bool ok = true;
        for (std::size_t g = 0; g < G; ++g) {
            ok = ok && h[g] > 0.0;
            mean_rate += h[g] / static_cast<double>(G);
        }
        if (!ok) { return; }
        // damped update, bounded: a share never moves further than +-25 % from equal
        for (std::size_t g = 0; g < G; ++g) { w[g] = std::min(1.25, std::max(0.75, 0.5 * mv->weights[g] + 0.5 * h[g] / mean_rate)); }
        mv->set_weights(w);
        ctx->tm.rebalances++;
    }

    // enqueue `count` iterations back to back, then poll the device state once (one ~100-byte read-back + sync).
    // Kernels of iterations enqueued past convergence exit immediately (device-side `done` flag), so x is never over-updated.
    void step(const std::uint64_t count) {
        PB_CUDA(cudaSetDevice(ctx->device));
        if (converged) { return; }
        const host_timer wall;
        std::uint64_t left = count;
        while (left > 0 && !converged) {
            // with rate-weighted shares the iterations go out in chunks that end where the next re-cut is due
            const std::uint64_t interval = static_cast<std::uint64_t>(std::max(1, ctx->balance_interval));
            const std::uint64_t chunk = balancing() ? std::min(left, interval - since_balance % interval) : left;
            const std::uint64_t mv0 = ctx->tm.matvec_calls;
            ctx->tile_timer.collect();
            const double tile0 = ctx->tile_timer.accum_ms;
            PB_CUDA(cudaEventRecord(ctx->ev_loop0, ctx->stream));
            for (std::uint64_t k = 0; k < chunk; ++k) { enqueue_iteration(iters_enqueued + k); }
            PB_CUDA(cudaEventRecord(ctx->ev_loop1, ctx->stream));
            iters_enqueued += chunk;
            since_balance += chunk;
            left -= chunk;
            poll();
            float ms = 0.f;
            PB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_loop0, ctx->ev_loop1));
            ctx->tm.cg_loop_ms += ms;  // device time of the iterations alone (events on the launching stream)
            if (balancing() && since_balance % interval == 0 && !converged) {
                ctx->tile_timer.collect();
                rebalance(ctx->tile_timer.accum_ms - tile0, ctx->tm.matvec_calls - mv0);
            }
        }
        loop_wall_ms += wall.ms();
        publish_stats();
    }

    void publish_stats() {
        ctx->tm.matvec_tile_ms = ctx->tile_timer.total_ms();
        ctx->tm.matvec_ms = ctx->matvec_timer.total_ms();
        ctx->tm.cg_iterations = last.iter;
        ctx->tm.cg_residuum = static_cast<double>(last.delta);
        ctx->tm.cg_target_residuum = static_cast<double>(eps) * static_cast<double>(eps) * static_cast<double>(last.delta0);
        ctx->tm.cg_avg_iteration_ms = last.iter > 0 ? loop_wall_ms / static_cast<double>(last.iter) : 0.0;
    }

    void poll() {
        CGState<T> *h_state = static_cast<CGState<T> *>(ctx->pinned);
        PB_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(CGState<T>), cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        last = *h_state;
        converged = last.done != 0;
        if (ctx->verbose != 0 && ctx->rank == 0) {
            std::printf("[plssvm_b200] iteration %llu residuum %g (target: %g)\n", static_cast<unsigned long long>(last.iter), static_cast<double>(last.delta),
                        static_cast<double>(eps * eps * last.delta0));
        }
    }

    // residual history: out[k] = r.r after k iterations, k = 0 .. min(iterations, TRACE_CAP)
    std::size_t get_trace(T *out, const std::size_t capacity) {
        PB_CUDA(cudaSetDevice(ctx->device));
        const std::size_t count = std::min<std::size_t>({ static_cast<std::size_t>(last.iter) + 1, static_cast<std::size_t>(TRACE_CAP) + 1, capacity });
        if (count > 0) {
            PB_CUDA(cudaMemcpyAsync(out, trace.p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        return count;
    }

    // bias and the last alpha (gpu_csvm.hpp:649-653); only the rank with `download` set copies the result to the host
    void finish(T *alpha_out, T *rho_out, std::uint64_t *iters_out, T *residual_out, const bool download) {
        PB_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        pb::cg_finish_kernel<T><<<1, pb::VEC_BLOCK, 0, st>>>(x.p, q_full.p, n, state.p);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        if (download) {
            PB_CUDA(cudaMemcpyAsync(alpha_out, x.p, n * sizeof(T), cudaMemcpyDeviceToHost, st));
            ctx->tm.d2h_bytes += static_cast<double>(n * sizeof(T));
        }
        poll();
        if (download) {
            alpha_out[n] = -last.sum_x;
            *rho_out = -last.bias;
            if (iters_out != nullptr) { *iters_out = last.iter; }
            if (residual_out != nullptr) {
                residual_out[0] = last.delta;
                residual_out[1] = last.delta0;
            }
        }
        if (ctx->verbose != 0 && ctx->rank == 0) { std::printf("[plssvm_b200] optimization finished, #iter = %llu\n", static_cast<unsigned long long>(last.iter)); }
        if (download) {  // residual history of this solve, kept for plssvm_b200_last_trace (the reference logs it per iteration, gpu_csvm.hpp:569-571)
            std::vector<T> tr(static_cast<std::size_t>(TRACE_CAP) + 1);
            const std::size_t cnt = get_trace(tr.data(), tr.size());
            ctx->last_trace.assign(tr.begin(), tr.begin() + static_cast<std::ptrdiff_t>(cnt));
        }
        publish_stats();
    }
};

template <typename T>
void solve_dataset_rank(plssvm_b200_ctx *ctx, plssvm_b200_dataset *ds, const T *y, const int kernel, const int degree, const T gamma, const T coef0, const T cost, const T eps,
                        const std::uint64_t max_iter, T *alpha_out, T *rho_out, std::uint64_t *iters_out, T *residual_out, const bool download) {
    cg_session<T> cg(ctx, ds, y, kernel, degree, gamma, coef0, cost, eps);
    ctx->tm.cg_max_iterations = max_iter;
    const std::uint64_t interval = ctx->check_interval > 0 ? static_cast<std::uint64_t>(ctx->check_interval) : (cg.n >= 16384 ? 1 : 8);
    while (cg.iters_enqueued < max_iter && !cg.converged) { cg.step(std::min<std::uint64_t>(interval, max_iter - cg.iters_enqueued)); }
    cg.finish(alpha_out, rho_out, iters_out, residual_out, download);
}

template <typename T>
void solve_dataset(plssvm_b200_ctx *ctx, plssvm_b200_dataset *ds, const T *y, const int kernel, const int degree, const T gamma, const T coef0, const T cost, const T eps,
                   const std::uint64_t max_iter, T *alpha_out, T *rho_out, std::uint64_t *iters_out, T *residual_out) {
    PB_REQUIRE(max_iter > 0, "The number of CG iterations must be greater than 0!");
    PB_REQUIRE(alpha_out != nullptr && rho_out != nullptr, "alpha_out and rho_out must not be NULL");
    check_group_dataset(ctx, ds, sizeof(T), "training");
    for_each_rank(ctx, [&](plssvm_b200_ctx *c, const int g) {
        solve_dataset_rank<T>(c, member_of(ds, g), y, kernel, degree, gamma, coef0, cost, eps, max_iter, alpha_out, rho_out, iters_out, residual_out, g == 0);
    });
}

// ---- w-kernel -----------------------------------------------------------------------------------------------------------------
template <typename T>
void run_w_kernel(plssvm_b200_ctx *ctx, const plssvm_b200_dataset *sv, const T *alpha_d, T *w_d /* ld entries, zero padded */) {
    const std::uint32_t d = static_cast<std::uint32_t>(sv->d);
    const std::uint32_t chunks = static_cast<std::uint32_t>((sv->N + pb::W_ROWS - 1) / pb::W_ROWS);
    dbuf<T> part(ctx, static_cast<std::size_t>(chunks) * d);
    PB_CUDA(cudaMemsetAsync(w_d, 0, sv->ld * sizeof(T), ctx->stream));
    const dim3 grid((d + 255) / 256, chunks);
    pb::w_partial_kernel<T><<<grid, 256, 0, ctx->stream>>>(static_cast<const T *>(sv->X.p), alpha_d, sv->N, d, static_cast<std::uint32_t>(sv->ld), part.p);
    pb::w_reduce_kernel<T><<<(d + 31) / 32, 256, 0, ctx->stream>>>(part.p, chunks, d, w_d);
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches += 2;
}

// ---- predict: csvm::predict_values (gpu_csvm.hpp:656-730) -------------------------------------------------------------------------
// m points (operand `pt`, already offset to the first of them) against all support vectors; out_d: m values on the device
template <typename T>
void predict_rows_device(plssvm_b200_ctx *ctx, const plssvm_b200_dataset *sv, const operand<T> &svo, const T *alpha_d, const T *w_d, const T rho, const operand<T> &pt,
                         const std::size_t m, const KernelParams<T> &kp, const int impl, T *out_d) {
    const std::uint32_t ld = static_cast<std::uint32_t>(sv->ld);
    if (kp.kernel == pb::K_LINEAR) {
        pb::linear_predict_kernel<T><<<static_cast<unsigned>((m + 7) / 8), 256, 0, ctx->stream>>>(pt.X, m, ld, w_d, rho, out_d);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        return;
    }
    TileParams<T> p{};
    set_operands(p, pt, svo);
    p.n_rows = static_cast<std::uint32_t>(m);
    p.n_cols = static_cast<std::uint32_t>(sv->N);
    p.ld = ld;
    p.T_rows = (p.n_rows + TILE - 1) / TILE;
    p.T_cols = (p.n_cols + TILE - 1) / TILE;
    p.tile_lo = 0;
    p.tile_hi = static_cast<std::uint64_t>(p.T_rows) * p.T_cols;
    if (super_tiled(impl)) { p.tile_hi = static_cast<std::uint64_t>((p.T_rows + 1) / 2) * ((p.T_cols + 1) / 2); }
    if (is_i8(impl)) { PB_REQUIRE(pt.i8 != nullptr && pt.scale != nullptr, "int8-slice tensor path needs the digit planes of the predict points"); }
    p.v = alpha_d;
    p.kp = kp;
    p.partial = workspace<T>(ctx, plssvm_b200_ctx::WS_PARTIAL, static_cast<std::size_t>(p.T_rows) * p.T_cols * TILE);
    ctx->tile_timer.begin(ctx->stream);
    launch_tiles<T, pb::MODE_RECT>(ctx, p, impl);
    ctx->tile_timer.end(ctx->stream);
    pb::reduce_partials_kernel<T, pb::MODE_RECT><<<p.T_rows, 512, 0, ctx->stream>>>(p.partial, out_d, p.n_rows, p.T_rows, p.T_cols, 0, p.tile_hi, 0, 0, T(1), -rho, 0, nullptr);
    PB_CUDA(cudaGetLastError());
    ctx->tm.kernel_launches++;
}

constexpr std::size_t PREDICT_BATCH = 32768;  // test points per pass (bounds the partial buffer: T_rows x T_cols x 128 values)

// contiguous share [lo, hi) of m test points owned by `rank`: multiples of 128 rows (whole digit-plane boxes)
inline void point_range(const std::size_t m, const int rank, const int world, std::size_t &lo, std::size_t &hi) {
    const std::size_t blocks = (m + TILE - 1) / TILE;
    lo = std::min(m, blocks * static_cast<std::size_t>(rank) / static_cast<std::size_t>(world) * TILE);
    hi = std::min(m, blocks * (static_cast<std::size_t>(rank) + 1) / static_cast<std::size_t>(world) * TILE);
}

// One rank's part of a predict call: the points [lo, hi) of its share, written to out[lo .. hi).  Several ranks: the test points are
// independent units, sharded by ranges with no data-path collective (the reference predicts on device 0 only, gpu_csvm.hpp:722).
template <typename T>
void predict_rank(plssvm_b200_ctx *ctx, plssvm_b200_dataset *sv, const T *alpha, const T rho, T *w_inout, int *w_valid, plssvm_b200_dataset *pts_ds, const host_matrix<T> &pts_host,
                  const std::size_t m, const int kernel, const int degree, const T gamma, const T coef0, T *out, const bool with_rho) {
    check_dataset(ctx, sv, sizeof(T), "support vector");
    PB_REQUIRE(alpha != nullptr && out != nullptr, "alpha and out must not be NULL");
    PB_REQUIRE(m > 0, "The data points to predict must not be empty!");
    validate_kernel_args(kernel, static_cast<double>(gamma));
    if (pts_ds != nullptr) {
        check_dataset(ctx, pts_ds, sizeof(T), "predict points");
        PB_REQUIRE(pts_ds->d == sv->d, "The number of features in the support vectors (" + std::to_string(sv->d) + ") must be the same as in the data points to predict (" +
                                           std::to_string(pts_ds->d) + ")!");
    }
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const KernelParams<T> kp{ kernel, degree, gamma, coef0 };
    const T shift_rho = with_rho ? rho : T(0);
    const bool leader = ctx->rank == 0 || !ctx->in_process_group();  // who writes the caller's w cache

    using ctx_t = plssvm_b200_ctx;
    T *alpha_d = workspace<T>(ctx, ctx_t::WS_ALPHA, sv->N);
    T *w_d = nullptr;
    PB_CUDA(cudaMemcpyAsync(alpha_d, alpha, sv->N * sizeof(T), cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += static_cast<double>(sv->N * sizeof(T));
    if (kernel == pb::K_LINEAR) {
        w_d = workspace<T>(ctx, ctx_t::WS_W, sv->ld);
        if (w_valid != nullptr && *w_valid != 0 && w_inout != nullptr) {
            PB_CUDA(cudaMemsetAsync(w_d, 0, sv->ld * sizeof(T), st));
            PB_CUDA(cudaMemcpyAsync(w_d, w_inout, sv->d * sizeof(T), cudaMemcpyHostToDevice, st));
            PB_CUDA(cudaStreamSynchronize(st));  // (another rank of the group may be about to fill the cache)
        } else {
            run_w_kernel<T>(ctx, sv, alpha_d, w_d);
            if (w_inout != nullptr && leader) {
                PB_CUDA(cudaMemcpyAsync(w_inout, w_d, sv->d * sizeof(T), cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
            }
        }
    }

    std::size_t lo = 0, hi = m;
    point_range(m, ctx->rank, ctx->world, lo, hi);

    // rbf: both operands centred at the feature means of the support vectors
    plssvm_b200_dataset *centre = kernel == pb::K_RBF ? sv : nullptr;
    const T *mean = centre != nullptr ? ensure_mean<T>(ctx, centre) : nullptr;
    int impl = resolve_impl<T>(ctx, sv->ld, hi - lo);
    operand<T> svo, pto;
    const bool tiles = kernel != pb::K_LINEAR;
    if (tiles) {
        svo = prepare_operand<T>(ctx, sv, centre, impl, kernel == pb::K_RBF);
        if (pts_ds != nullptr) { pto = prepare_operand<T>(ctx, pts_ds, centre, impl, kernel == pb::K_RBF); }
        if (is_i8(impl) && !i8_forced(ctx) && (svo.bad_rows != 0 || pto.bad_rows != 0)) {
            impl = 2;
            svo = prepare_operand<T>(ctx, sv, centre, impl, kernel == pb::K_RBF);
            if (pts_ds != nullptr) { pto = prepare_operand<T>(ctx, pts_ds, centre, impl, kernel == pb::K_RBF); }
        }
    } else if (pts_ds != nullptr) {
        pto.X = static_cast<const T *>(pts_ds->X.p);
    }
    const bool auto_i8 = tiles && is_i8(impl) && !i8_forced(ctx);  // host-staged batches with badly scaled points are re-run with the floating-point tiles
    operand<T> svo_fallback;
    bool have_fallback = false;

    // Test points are processed in batches of PREDICT_BATCH rows with 64-bit offsets (the reference's int indexing overflows at
    // this size: predict_kernel.cu:40-42).  Host points are staged through two HBM buffers: the H2D copy of batch b + 1 runs
    // on the copy stream while the tile kernel of batch b runs on the compute stream.  Values collect in HBM and are
    // downloaded once per super-batch.
    constexpr std::size_t SUPER_BATCH = std::size_t{ 1 } << 22;
    const std::size_t m_own = hi - lo;
    const std::size_t stage_rows = std::min(m_own, PREDICT_BATCH);
    T *stage_X[2] = { nullptr, nullptr }, *stage_sq[2] = { nullptr, nullptr }, *stage_hi[2] = { nullptr, nullptr }, *stage_lo[2] = { nullptr, nullptr };
    const std::size_t ld8 = pitch_i8(sv->d);
    std::int8_t *stage_i8[2] = { nullptr, nullptr };
    T *stage_sc[2] = { nullptr, nullptr };
    int *bad_d = nullptr;
    auto alloc_split_stages = [&](const int n_stage) {
        for (int i = 0; i < n_stage; ++i) {
            stage_hi[i] = workspace<T>(ctx, ctx_t::WS_HI0 + i, stage_rows * sv->ld);
            stage_lo[i] = workspace<T>(ctx, ctx_t::WS_LO0 + i, stage_rows * sv->ld);
        }
    };
    const int n_stage = m_own > PREDICT_BATCH ? 2 : 1;
    if (pts_ds == nullptr && m_own > 0) {
        for (int i = 0; i < n_stage; ++i) {
            stage_X[i] = workspace<T>(ctx, ctx_t::WS_STAGE0 + i, stage_rows * sv->ld);
            stage_sq[i] = workspace<T>(ctx, ctx_t::WS_SQ0 + i, stage_rows);
            if (tiles && is_i8(impl)) {
                stage_i8[i] = workspace<std::int8_t>(ctx, ctx_t::WS_I8_0 + i, static_cast<std::size_t>(pb::I8<T>::S_EXACT) * rows_i8(stage_rows) * ld8);
                stage_sc[i] = workspace<T>(ctx, ctx_t::WS_SC0 + i, stage_rows);
            }
            if (sv->ld != sv->d) { PB_CUDA(cudaMemsetAsync(stage_X[i], 0, stage_rows * sv->ld * sizeof(T), st)); }  // pad columns stay zero
        }
        if (sizeof(T) == 4 && tiles && impl >= 2 && !is_i8(impl)) { alloc_split_stages(n_stage); }
        bad_d = workspace<int>(ctx, ctx_t::WS_MISC, 16);
        PB_CUDA(cudaEventRecord(ctx->ev_computed[0], st));  // the copy stream must not start before the memsets above
        PB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_computed[0], 0));
    }
    T *out_d = workspace<T>(ctx, ctx_t::WS_OUT, std::max<std::size_t>(1, std::min(m_own, SUPER_BATCH)));
    std::size_t batch_index = 0;
    for (std::size_t s0 = lo; s0 < hi; s0 += SUPER_BATCH) {
        const std::size_t ms = std::min(SUPER_BATCH, hi - s0);
        for (std::size_t p0 = s0; p0 < s0 + ms; p0 += PREDICT_BATCH, ++batch_index) {
            const std::size_t mb = std::min(PREDICT_BATCH, s0 + ms - p0);
            operand<T> pt;
            int batch_impl = impl;
            if (pts_ds != nullptr) {
                pt = pto;
                pt.X = pto.X + p0 * pts_ds->ld;
                if (pto.sq != nullptr) { pt.sq = pto.sq + p0; }
                if (pto.i8 != nullptr) {
                    pt.i8 = pto.i8 + p0 * pts_ds->ld8 * static_cast<std::size_t>(pts_ds->i8_slices);  // p0 is a multiple of 128 rows: whole boxes
                    pt.scale = pto.scale + p0;
                }
                if (pto.hi != nullptr) {
                    pt.hi = pto.hi + p0 * pts_ds->ld;
                    pt.lo = pto.lo + p0 * pts_ds->ld;
                }
            } else {
                const int buf = static_cast<int>(batch_index & 1);
                if (batch_index >= 2) { PB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_computed[buf], 0)); }  // buffer free again
                upload_rows<T>(ctx, stage_X[buf], sv->ld, pts_host, p0, p0 + mb, ctx->copy_stream);
                PB_CUDA(cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
                PB_CUDA(cudaStreamWaitEvent(st, ctx->ev_copied[buf], 0));
                ctx->tm.h2d_bytes += static_cast<double>(mb * sv->d * sizeof(T));
                if (kernel == pb::K_RBF) {
                    const std::size_t total = mb * sv->ld;
                    const unsigned grid = static_cast<unsigned>(std::min<std::size_t>((total + 255) / 256, static_cast<std::size_t>(ctx->num_sms) * 32));
                    pb::center_rows_kernel<T><<<grid, 256, 0, st>>>(stage_X[buf], stage_X[buf], mb, static_cast<std::uint32_t>(sv->ld), mean);  // in place
                    pb::row_norms_kernel<T><<<static_cast<unsigned>((mb + 7) / 8), 256, 0, st>>>(stage_X[buf], mb, static_cast<std::uint32_t>(sv->ld), stage_sq[buf], nullptr);
                    PB_CUDA(cudaGetLastError());
                    ctx->tm.kernel_launches += 2;
                }
                pt.X = stage_X[buf];
                pt.sq = stage_sq[buf];
                if (tiles && is_i8(impl)) {
                    if (auto_i8) { PB_CUDA(cudaMemsetAsync(bad_d, 0, sizeof(int), st)); }
                    run_split_i8<T>(ctx, i8_slices_for<T>(impl), stage_X[buf], mb, sv->d, sv->ld, stage_i8[buf], stage_i8[buf], TILE, ld8, stage_sc[buf], auto_i8 ? bad_d : nullptr, nullptr,
                                    i8_slab_for<T>(impl), st);
                    pt.i8 = pt.i8b = stage_i8[buf];
                    pt.scale = stage_sc[buf];
                    pt.ld8 = static_cast<std::uint32_t>(ld8);
                    if (auto_i8) {
                        // the same dynamic-range guard the resident operands get: one 4-byte read-back per batch (the next batch's upload is
                        // already in flight on the copy stream, so the pipeline keeps running)
                        int *h = static_cast<int *>(ctx->pinned) + 512;
                        PB_CUDA(cudaMemcpyAsync(h, bad_d, sizeof(int), cudaMemcpyDeviceToHost, st));
                        PB_CUDA(cudaStreamSynchronize(st));
                        if (*h != 0) {
                            batch_impl = 2;
                            ctx->tm.fallback_batches++;
                            if (!have_fallback) {
                                svo_fallback = prepare_operand<T>(ctx, sv, centre, 2, kernel == pb::K_RBF);
                                if (sizeof(T) == 4) { alloc_split_stages(n_stage); }
                                have_fallback = true;
                            }
                        }
                    }
                }
                if constexpr (sizeof(T) == 4) {
                    if (tiles && batch_impl >= 2 && !is_i8(batch_impl)) {
                        const std::size_t total = mb * sv->ld;
                        const unsigned grid = static_cast<unsigned>(std::min<std::size_t>((total + 255) / 256, static_cast<std::size_t>(ctx->num_sms) * 32));
                        pb::split_tf32_kernel<<<grid, 256, 0, st>>>(stage_X[buf], stage_hi[buf], stage_lo[buf], total);
                        PB_CUDA(cudaGetLastError());
                        ctx->tm.kernel_launches++;
                        pt.hi = stage_hi[buf];
                        pt.lo = stage_lo[buf];
                    }
                }
            }
            predict_rows_device<T>(ctx, sv, batch_impl == impl ? svo : svo_fallback, alpha_d, w_d, shift_rho, pt, mb, kp, batch_impl, out_d + (p0 - s0));
            if (pts_ds == nullptr) { PB_CUDA(cudaEventRecord(ctx->ev_computed[batch_index & 1], st)); }
        }
        PB_CUDA(cudaMemcpyAsync(out + s0, out_d, ms * sizeof(T), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        ctx->tm.d2h_bytes += static_cast<double>(ms * sizeof(T));
    }
    PB_CUDA(cudaStreamSynchronize(st));

    // one process per GPU: every rank returns all m values (sum of the zero-extended shares; 8 bytes per point over NVLink)
    if (ctx->world > 1 && ctx->comm != nullptr && !ctx->in_process_group()) {
        for (std::size_t s0 = 0; s0 < m; s0 += SUPER_BATCH) {
            const std::size_t ms = std::min(SUPER_BATCH, m - s0);
            T *buf = workspace<T>(ctx, ctx_t::WS_OUT, std::min(m, SUPER_BATCH));
            PB_CUDA(cudaMemsetAsync(buf, 0, ms * sizeof(T), st));
            const std::size_t a = std::max(lo, s0), b = std::min(hi, s0 + ms);
            if (b > a) { PB_CUDA(cudaMemcpyAsync(buf + (a - s0), out + a, (b - a) * sizeof(T), cudaMemcpyHostToDevice, st)); }
            all_reduce_sum(ctx, buf, ms);
            PB_CUDA(cudaMemcpyAsync(out + s0, buf, ms * sizeof(T), cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
        }
    }
    ctx->tm.matvec_tile_ms = ctx->tile_timer.total_ms();
}

template <typename T>
void predict_common(plssvm_b200_ctx *ctx, plssvm_b200_dataset *sv, const T *alpha, const T rho, T *w_inout, int *w_valid, plssvm_b200_dataset *pts_ds, const host_matrix<T> &pts_host,
                    const std::size_t m, const int kernel, const int degree, const T gamma, const T coef0, T *out, const bool with_rho) {
    check_group_dataset(ctx, sv, sizeof(T), "support vector");
    if (pts_ds != nullptr) { check_group_dataset(ctx, pts_ds, sizeof(T), "predict points"); }
    const bool fill_w = kernel == pb::K_LINEAR && w_valid != nullptr && *w_valid == 0 && w_inout != nullptr;
    for_each_rank(ctx, [&](plssvm_b200_ctx *c, const int g) {
        predict_rank<T>(c, member_of(sv, g), alpha, rho, w_inout, w_valid, pts_ds != nullptr ? member_of(pts_ds, g) : nullptr, pts_host, m, kernel, degree, gamma, coef0, out, with_rho);
    });
    if (fill_w) { *w_valid = 1; }
}

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return PLSSVM_B200_OK;
    } catch (const api_error &e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return PLSSVM_B200_ERR_INTERNAL;
    }
}

// kernel-granular helpers ------------------------------------------------------------------------------------------------------
template <typename T>
void api_q_kernel(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const int kernel, const int degree, const T gamma, const T coef0, T *q_out, T *k_last) {
    check_group_dataset(ctx, X, sizeof(T), "training");
    PB_REQUIRE(X->N >= 2 && q_out != nullptr, "q_kernel needs at least two data points and an output buffer");
    validate_kernel_args(kernel, static_cast<double>(gamma));
    PB_CUDA(cudaSetDevice(ctx->device));
    dbuf<T> q_full(ctx, X->N);
    run_q_kernel<T>(ctx, X, KernelParams<T>{ kernel, degree, gamma, coef0 }, q_full.p);
    PB_CUDA(cudaMemcpyAsync(q_out, q_full.p, (X->N - 1) * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    T last{};
    PB_CUDA(cudaMemcpyAsync(&last, q_full.p + (X->N - 1), sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (k_last != nullptr) { *k_last = last; }
}

template <typename T>
void api_matvec_rank(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const T *q, const T *v, const T QA_cost, const T cost_inv, const T add, const int kernel, const int degree,
                     const T gamma, const T coef0, T *ret_inout, const bool download) {
    PB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const std::uint32_t n = static_cast<std::uint32_t>(X->N - 1);
    dbuf<T> q_d(ctx, n), v_d(ctx, n), ret_d(ctx, n), out_d(ctx, n), qa_d(ctx, 1);
    PB_CUDA(cudaMemcpyAsync(q_d.p, q, n * sizeof(T), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemcpyAsync(v_d.p, v, n * sizeof(T), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemcpyAsync(ret_d.p, ret_inout, n * sizeof(T), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemcpyAsync(qa_d.p, &QA_cost, sizeof(T), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaStreamSynchronize(st));  // the inputs are read before any rank of a group writes the result back
    {
        matvec_plan<T> mv(ctx, X, KernelParams<T>{ kernel, degree, gamma, coef0 }, q_d.p, qa_d.p, cost_inv, nullptr);
        ctx->tm.matvec_flops = static_cast<double>(X->d) * static_cast<double>(n) * (static_cast<double>(n) + 1.0);
        mv.run(v_d.p, out_d.p);
        pb::axpy_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(ret_d.p, out_d.p, add, n);
        PB_CUDA(cudaGetLastError());
        ctx->tm.kernel_launches++;
        PB_CUDA(cudaStreamSynchronize(st));
    }
    ctx->tm.matvec_tile_ms = ctx->tile_timer.total_ms();
    ctx->tm.matvec_ms = ctx->matvec_timer.total_ms();
    if (download) {
        PB_CUDA(cudaMemcpyAsync(ret_inout, ret_d.p, n * sizeof(T), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
    }
}

template <typename T>
void api_matvec(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const T *q, const T *v, const T QA_cost, const T cost_inv, const T add, const int kernel, const int degree,
                const T gamma, const T coef0, T *ret_inout) {
    check_group_dataset(ctx, X, sizeof(T), "training");
    PB_REQUIRE(X->N >= 2, "The data must contain at least two data points!");
    PB_REQUIRE(q != nullptr && v != nullptr && ret_inout != nullptr, "q, v and ret must not be NULL");
    PB_REQUIRE(add == T(1) || add == T(-1), "add must either be -1.0 or 1.0, but is " + std::to_string(add) + "!");
    PB_REQUIRE(cost_inv != T(0), "cost must not be 0.0 since it is 1 / plssvm::cost!");
    validate_kernel_args(kernel, static_cast<double>(gamma));
    const std::size_t G = std::max<std::size_t>(1, ctx->members.size());
    if (G == 1) {
        api_matvec_rank<T>(ctx, X, q, v, QA_cost, cost_inv, add, kernel, degree, gamma, coef0, ret_inout, true);
        return;
    }
    // device group: every rank reads the caller's `ret`, so the result goes through a private copy
    const std::size_t n = X->N - 1;
    std::vector<T> ret_in(ret_inout, ret_inout + n), ret_out(n);
    for_each_rank(ctx, [&](plssvm_b200_ctx *c, const int g) {
        std::vector<T> scratch;
        T *dst = ret_out.data();
        if (g != 0) {
            scratch = ret_in;
            dst = scratch.data();
        } else {
            std::copy(ret_in.begin(), ret_in.end(), ret_out.begin());
        }
        api_matvec_rank<T>(c, member_of(X, g), q, v, QA_cost, cost_inv, add, kernel, degree, gamma, coef0, dst, g == 0);
    });
    std::copy(ret_out.begin(), ret_out.end(), ret_inout);
}

template <typename T>
void api_w_kernel(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const T *alpha, T *w_out) {
    check_group_dataset(ctx, SV, sizeof(T), "support vector");
    PB_REQUIRE(alpha != nullptr && w_out != nullptr, "alpha and w_out must not be NULL");
    PB_CUDA(cudaSetDevice(ctx->device));
    dbuf<T> alpha_d(ctx, SV->N), w_d(ctx, SV->ld);
    PB_CUDA(cudaMemcpyAsync(alpha_d.p, alpha, SV->N * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    run_w_kernel<T>(ctx, SV, alpha_d.p, w_d.p);
    PB_CUDA(cudaMemcpyAsync(w_out, w_d.p, SV->d * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
}

plssvm_b200_ctx *create_device_context(const int device) {
    PB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        throw api_error(PLSSVM_B200_ERR_CUDA, std::string("plssvm_b200 targets sm_100a (B200) only, found '") + prop.name + "' with compute capability " +
                                                  std::to_string(prop.major) + "." + std::to_string(prop.minor));
    }
    auto ctx = std::make_unique<plssvm_b200_ctx>();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    int cluster_launch = 0;  // CTA pairs need thread-block clusters and two SMs; without them the automatic choice stays on single CTAs
    PB_CUDA(cudaDeviceGetAttribute(&cluster_launch, cudaDevAttrClusterLaunch, device));
    ctx->pairs_ok = (cluster_launch != 0 && ctx->num_sms >= 2) ? 1 : 0;
    ctx->leader = ctx.get();
    PB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreate(&ctx->ev_loop0));
    PB_CUDA(cudaEventCreate(&ctx->ev_loop1));
    PB_CUDA(cudaMallocHost(&ctx->pinned, 4096));
    PB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&ctx->ev_computed[i], cudaEventDisableTiming));
    }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres{};
    PB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) { throw api_error(PLSSVM_B200_ERR_CUDA, "driver does not export cuTensorMapEncodeTiled"); }
    ctx->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    ctx->tm.n_devices = 1;
    return ctx.release();
}

void destroy_device_context(plssvm_b200_ctx *ctx) {
    if (ctx == nullptr) { return; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm != nullptr) { nccl_api::get().CommDestroy(ctx->comm); }
    cudaEventDestroy(ctx->ev_loop0);
    cudaEventDestroy(ctx->ev_loop1);
    cudaFreeHost(ctx->pinned);
    for (int i = 0; i < plssvm_b200_ctx::WS_COUNT; ++i) { cudaFree(ctx->ws_ptr[i]); }
    pool_trim(ctx);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(ctx->ev_copied[i]);
        cudaEventDestroy(ctx->ev_computed[i]);
        if (ctx->ring[i] != nullptr) {
            cudaFreeHost(ctx->ring[i]);
            cudaEventDestroy(ctx->ev_ring[i]);
        }
    }
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

template <typename F>
void for_all_members(plssvm_b200_ctx *ctx, F &&f) {
    if (ctx->members.size() > 1) {
        for (auto *m : ctx->members) { f(m); }
    } else {
        f(ctx);
    }
}

}  // namespace

// ================================================================================================================================
//                                                           C ABI
// ================================================================================================================================
struct plssvm_b200_cg {
    plssvm_b200_ctx *ctx = nullptr;
    int elem_size = 0;
    std::vector<std::unique_ptr<cg_session_base>> ranks;  // one session per device of the context
};

namespace {

// frees a CG session handle (idempotent through the handle registry)
void cg_release(plssvm_b200_cg *cg) {
    if (!handle_remove(cg)) { return; }
    auto &live = cg->ctx->live_sessions;
    live.erase(std::remove(live.begin(), live.end(), cg), live.end());
    for (auto &s : cg->ranks) {
        if (s != nullptr) {
            cudaSetDevice(s->ctx->device);
            cudaStreamSynchronize(s->ctx->stream);
            s.reset();
        }
    }
    delete cg;
}

}  // namespace

extern "C" {

const char *plssvm_b200_last_error(void) { return g_last_error.c_str(); }

int plssvm_b200_device_count(int *count) {
    return guarded([&] {
        PB_REQUIRE(count != nullptr, "count is NULL");
        PB_CUDA(cudaGetDeviceCount(count));
    });
}

int plssvm_b200_has_experimental(void) { return EXPERIMENTAL ? 1 : 0; }

int plssvm_b200_create(const int *device_ids, int n_dev, plssvm_b200_ctx **out) {
    return guarded([&] {
        PB_REQUIRE(out != nullptr, "out is NULL");
        int count = 0;
        PB_CUDA(cudaGetDeviceCount(&count));
        if (count == 0) { throw api_error(PLSSVM_B200_ERR_CUDA, "CUDA backend selected but no CUDA devices were found!"); }
        std::vector<int> devs;
        if (device_ids != nullptr) {
            PB_REQUIRE(n_dev >= 1, "n_dev must be at least 1 when device_ids is given");
            devs.assign(device_ids, device_ids + n_dev);
        } else {
            const int n = n_dev <= 0 ? count : n_dev;
            for (int i = 0; i < n; ++i) { devs.push_back(i); }
        }
        PB_REQUIRE(devs.size() <= 64, "at most 64 devices per context");
        for (std::size_t i = 0; i < devs.size(); ++i) {
            PB_REQUIRE(devs[i] >= 0 && devs[i] < count, "invalid device index " + std::to_string(devs[i]));
            for (std::size_t j = 0; j < i; ++j) { PB_REQUIRE(devs[j] != devs[i], "device " + std::to_string(devs[i]) + " listed twice"); }
        }
        std::vector<plssvm_b200_ctx *> members;
        try {
            for (const int dev : devs) { members.push_back(create_device_context(dev)); }
            if (members.size() > 1) {
                const nccl_api &nccl = nccl_api::get();
                std::vector<nccl_api::comm_t> comms(members.size(), nullptr);
                nccl.check(nccl.CommInitAll(comms.data(), static_cast<int>(devs.size()), devs.data()), "ncclCommInitAll");
                for (std::size_t g = 0; g < members.size(); ++g) {
                    members[g]->comm = comms[g];
                    members[g]->rank = static_cast<int>(g);
                    members[g]->world = static_cast<int>(members.size());
                    members[g]->leader = members[0];
                    members[g]->tm.n_devices = static_cast<int>(members.size());
                }
                members[0]->members = members;
                // NCCL connects its channels lazily, at the first collective (hundreds of milliseconds): do that here, like the device
                // initialisation of the reference's constructor (csvm.cu:48-86), not inside the first fit / predict call
                for_each_rank(members[0], [&](plssvm_b200_ctx *c, const int) {
                    double *buf = workspace<double>(c, plssvm_b200_ctx::WS_MISC, 64);
                    PB_CUDA(cudaMemsetAsync(buf, 0, 64 * sizeof(double), c->stream));
                    all_reduce_sum(c, buf, 16);
                    nccl.check(nccl.AllGather(buf + c->rank, buf, 1, NCCL_FLOAT64, c->comm, c->stream), "ncclAllGather");
                    PB_CUDA(cudaStreamSynchronize(c->stream));
                });
            }
        } catch (...) {
            for (auto *m : members) { destroy_device_context(m); }
            throw;
        }
        PB_CUDA(cudaSetDevice(members[0]->device));
        handle_add(members[0]);
        *out = members[0];
    });
}

int plssvm_b200_destroy(plssvm_b200_ctx *ctx) {
    return guarded([&] {
        if (!handle_live(ctx)) { return; }  // NULL or already destroyed
        // sessions and data sets created from this context die with it (their handles become no-ops)
        while (!ctx->live_sessions.empty()) { cg_release(ctx->live_sessions.back()); }
        while (!ctx->live_datasets.empty()) { dataset_destroy(ctx->live_datasets.back()); }
        handle_remove(ctx);
        const std::vector<plssvm_b200_ctx *> members = ctx->members;
        if (members.size() > 1) {
            for (auto *m : members) { destroy_device_context(m); }
        } else {
            destroy_device_context(ctx);
        }
    });
}

int plssvm_b200_num_devices(const plssvm_b200_ctx *ctx, int *count) {
    return guarded([&] {
        PB_REQUIRE(handle_live(ctx) && count != nullptr, "ctx (NULL or destroyed) or count is NULL");
        *count = static_cast<int>(std::max<std::size_t>(1, ctx->members.size()));
    });
}

int plssvm_b200_set_option(plssvm_b200_ctx *ctx, const char *key, long long value) {
    return guarded([&] {
        PB_REQUIRE(handle_live(ctx) && key != nullptr, "ctx (NULL or destroyed) or key is NULL");
        const std::string k(key);
        if (k == "impl") {
            const bool known = value == 0 || value == 1 || value == 2 || value == 6 || value == 7 || value == 10 || (EXPERIMENTAL && (value == 4 || value == 5 || value == 8 || value == 9 || value == 11));
            PB_REQUIRE(known, std::string("impl must be 0 (auto), 1 (simt), 2 (floating-point tensor tiles), 6 (int8-slice tcgen05 tiles), 7 (int8-slice tiles with the exact-input "
                                          "slice count: fp32 4 instead of 3 slices) or 10 (int8-slice tiles on CTA pairs, cta_group::2 with the wide-N instructions)") +
                                  (EXPERIMENTAL ? "; experimental: 4 (fp32: CTA-pair 3xTF32), 5 (fp32: 128x256 3xTF32), 8 (int8-slice tiles, 2 x 2 CTA clusters with TMA multicast), 9 (fp32: "
                                                  "int8-slice tiles on CTA pairs, cta_group::2), 11 (int8-slice tiles, clusters of two CTAs sharing the A planes through TMA multicast)"
                                                : "; 4 / 5 / 8 / 9 / 11 need a build with -DPLSSVM_B200_EXPERIMENTAL"));
        } else if (k == "check_interval") {
            PB_REQUIRE(value >= 0 && value <= 1000000, "check_interval out of range");
        } else if (k == "max_ctas") {
            PB_REQUIRE(value >= 0 && value <= 4096, "max_ctas out of range");
        } else if (k == "balance_interval") {
            PB_REQUIRE(value >= 1 && value <= 1000000, "balance_interval out of range");
        } else if (k == "virtual_world" || k == "virtual_rank" || k == "virtual_skew") {
            PB_REQUIRE(ctx->comm == nullptr && ctx->members.size() <= 1, "virtual ranks are a single-device testing aid: the context must not have a communicator");
            PB_REQUIRE(value >= 0 && value <= 64, k + " out of range");
            if (k == "virtual_world") {
                ctx->world = std::max(1, static_cast<int>(value));
                ctx->rank = std::min(ctx->rank, ctx->world - 1);
            } else if (k == "virtual_rank") {
                PB_REQUIRE(value < ctx->world, "virtual_rank must be below virtual_world");
                ctx->rank = static_cast<int>(value);
            } else {
                ctx->virtual_skew = static_cast<int>(value);
            }
            return;
        } else if (k != "verbose" && k != "ignore_convergence" && k != "linear_factorized" && k != "balance" && k != "shard_upload" && k != "fp32_fast_drain" && k != "tile_stats" && k != "i8_a_via_tmem" && k != "fp32_pair") {
            throw api_error(PLSSVM_B200_ERR_INVALID, "unknown option '" + k + "'");
        }
        for_all_members(ctx, [&](plssvm_b200_ctx *c) {
            if (k == "impl") {
                c->impl = static_cast<int>(value);
            } else if (k == "check_interval") {
                c->check_interval = static_cast<int>(value);
            } else if (k == "verbose") {
                c->verbose = value != 0;
            } else if (k == "ignore_convergence") {
                c->ignore_convergence = value != 0;
            } else if (k == "max_ctas") {
                c->max_ctas = static_cast<int>(value);
            } else if (k == "linear_factorized") {
                c->linear_factorized = value != 0;
            } else if (k == "balance") {
                c->balance = value != 0;
            } else if (k == "balance_interval") {
                c->balance_interval = static_cast<int>(value);
            } else if (k == "shard_upload") {
                c->shard_upload = value != 0;
            } else if (k == "fp32_fast_drain") {
                c->fp32_fast_drain = value != 0;
            } else if (k == "tile_stats") {
                c->tile_stats = static_cast<int>(value);  // 1: counters into the timings; 2: also one JSON line per tile launch on stderr
            } else if (k == "i8_a_via_tmem") {
                c->i8_a_via_tmem = value != 0;
            } else if (k == "fp32_pair") {
                c->fp32_pair = value != 0;
            }
        });
    });
}

int plssvm_b200_get_timings(const plssvm_b200_ctx *ctx, plssvm_b200_timings *out) {
    return guarded([&] {
        PB_REQUIRE(handle_live(ctx) && out != nullptr, "ctx (NULL or destroyed) or out is NULL");
        *out = ctx->tm;
        // device group: device times are the maximum over the devices, byte and launch counts the sums
        for (std::size_t g = 1; g < ctx->members.size(); ++g) {
            const plssvm_b200_timings &t = ctx->members[g]->tm;
            out->cg_loop_ms = std::max(out->cg_loop_ms, t.cg_loop_ms);
            out->matvec_ms = std::max(out->matvec_ms, t.matvec_ms);
            out->matvec_tile_ms = std::max(out->matvec_tile_ms, t.matvec_tile_ms);
            out->kernel_launches += t.kernel_launches;
            out->h2d_bytes += t.h2d_bytes;
            out->d2h_bytes += t.d2h_bytes;
            out->fallback_batches += t.fallback_batches;
        }
        out->n_devices = ctx->world;
    });
}

int plssvm_b200_last_trace(const plssvm_b200_ctx *ctx, double *out, size_t capacity, size_t *count) {
    return guarded([&] {
        PB_REQUIRE(handle_live(ctx) && count != nullptr && (out != nullptr || capacity == 0), "ctx (NULL or destroyed), out or count is NULL");
        const std::size_t n = std::min(capacity, ctx->last_trace.size());
        std::copy(ctx->last_trace.begin(), ctx->last_trace.begin() + static_cast<std::ptrdiff_t>(n), out);
        *count = n;
    });
}

int plssvm_b200_comm_unique_id(void *id128) {
    return guarded([&] {
        PB_REQUIRE(id128 != nullptr, "id buffer is NULL");
        const nccl_api &nccl = nccl_api::get();
        nccl_api::unique_id id{};
        nccl.check(nccl.GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id128, &id, sizeof(id));
    });
}

int plssvm_b200_comm_init(plssvm_b200_ctx *ctx, int rank, int world_size, const void *id128) {
    return guarded([&] {
        PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");
        PB_REQUIRE(ctx->members.size() <= 1, "plssvm_b200_comm_init is for single-device contexts (one process per GPU); this context already drives several devices");
        PB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "invalid rank / world size");
        if (world_size == 1) {
            ctx->rank = 0;
            ctx->world = 1;
            return;
        }
        PB_REQUIRE(id128 != nullptr, "id buffer is NULL");
        PB_CUDA(cudaSetDevice(ctx->device));
        const nccl_api &nccl = nccl_api::get();
        nccl_api::unique_id id{};
        std::memcpy(&id, id128, sizeof(id));
        nccl.check(nccl.CommInitRank(&ctx->comm, world_size, id, rank), "ncclCommInitRank");
        ctx->rank = rank;
        ctx->world = world_size;
        ctx->tm.n_devices = world_size;
        double *buf = workspace<double>(ctx, plssvm_b200_ctx::WS_MISC, 64);  // first collectives: NCCL connects its channels here, not inside the first solve
        PB_CUDA(cudaMemsetAsync(buf, 0, 64 * sizeof(double), ctx->stream));
        all_reduce_sum(ctx, buf, 16);
        nccl.check(nccl.AllGather(buf + rank, buf, 1, NCCL_FLOAT64, ctx->comm, ctx->stream), "ncclAllGather");
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

uint64_t plssvm_b200_tile_size(void) { return static_cast<uint64_t>(pb::TILE); }
uint64_t plssvm_b200_tri_num_tiles(uint64_t tiles_per_side) { return pb::tri_num_tiles(tiles_per_side); }
uint64_t plssvm_b200_tri_encode(uint64_t tiles_per_side, uint64_t I, uint64_t J) { return pb::tri_encode(tiles_per_side, I, J); }
void plssvm_b200_tri_decode(uint64_t tiles_per_side, uint64_t L, uint32_t *I, uint32_t *J) { pb::tri_decode(tiles_per_side, L, *I, *J); }
void plssvm_b200_rank_range(uint64_t total, int rank, int world_size, uint64_t *lo, uint64_t *hi) { pb::rank_range(total, rank, world_size, *lo, *hi); }
void plssvm_b200_weighted_range(uint64_t total, int rank, int world_size, const double *weights, uint64_t *lo, uint64_t *hi) {
    pb::weighted_range(total, rank, world_size, weights, *lo, *hi);
}
uint64_t plssvm_b200_i8_plane_offset(uint64_t row, uint32_t feature, uint32_t plane, uint32_t planes, uint32_t box_rows, uint32_t slabs, uint32_t slab_bytes) {
    return static_cast<uint64_t>(pb::i8_boxed_offset(static_cast<std::size_t>(row), feature, plane, planes, box_rows, slabs, slab_bytes));
}

int plssvm_b200_dataset_destroy(plssvm_b200_dataset *ds) {
    return guarded([&] { dataset_destroy(ds); });
}

int plssvm_b200_cg_step(plssvm_b200_cg *cg, uint64_t iterations, uint64_t *iterations_done, int *converged) {
    return guarded([&] {
        PB_REQUIRE(handle_live(cg) && !cg->ranks.empty(), "cg session is NULL or has been finished / aborted / destroyed with its context");
        for_each_rank(cg->ctx, [&](plssvm_b200_ctx *, const int g) {
            if (cg->elem_size == 8) {
                static_cast<cg_session<double> *>(cg->ranks[g].get())->step(iterations);
            } else {
                static_cast<cg_session<float> *>(cg->ranks[g].get())->step(iterations);
            }
        });
        if (cg->elem_size == 8) {
            auto *s = static_cast<cg_session<double> *>(cg->ranks[0].get());
            if (iterations_done != nullptr) { *iterations_done = s->last.iter; }
            if (converged != nullptr) { *converged = s->converged ? 1 : 0; }
        } else {
            auto *s = static_cast<cg_session<float> *>(cg->ranks[0].get());
            if (iterations_done != nullptr) { *iterations_done = s->last.iter; }
            if (converged != nullptr) { *converged = s->converged ? 1 : 0; }
        }
    });
}

int plssvm_b200_cg_abort(plssvm_b200_cg *cg) {
    return guarded([&] { cg_release(cg); });
}

#define PB_INSTANTIATE(SUF, T)                                                                                                                                                   \
    int plssvm_b200_dataset_create_##SUF(plssvm_b200_ctx *ctx, const T *X, size_t N, size_t d, int src_on_device, plssvm_b200_dataset **out) {                                 \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(out != nullptr, "out is NULL");                                                                                                                           \
            PB_REQUIRE(X != nullptr, "The data must not be empty!");                                                                                                             \
            *out = src_on_device != 0 ? dataset_create<T>(ctx, host_matrix<T>{}, X, N, d) : dataset_create<T>(ctx, host_matrix<T>{ X, nullptr, d }, nullptr, N, d);              \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_dataset_create_rows_##SUF(plssvm_b200_ctx *ctx, const T *const *rows, size_t N, size_t d, plssvm_b200_dataset **out) {                                     \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(out != nullptr, "out is NULL");                                                                                                                           \
            PB_REQUIRE(rows != nullptr, "The data must not be empty!");                                                                                                          \
            *out = dataset_create<T>(ctx, host_matrix<T>{ nullptr, rows, d }, nullptr, N, d);                                                                                    \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_cg_begin_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const T *y, int kernel, int degree, T gamma, T coef0, T cost, T eps, plssvm_b200_cg **out) { \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx) && out != nullptr, "context (NULL or destroyed) or out is NULL");                                                                                              \
            check_group_dataset(ctx, X, sizeof(T), "training");                                                                                                                  \
            for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                  \
            auto holder = std::make_unique<plssvm_b200_cg>();                                                                                                                    \
            holder->ctx = ctx;                                                                                                                                                   \
            holder->elem_size = static_cast<int>(sizeof(T));                                                                                                                     \
            holder->ranks.resize(std::max<std::size_t>(1, ctx->members.size()));                                                                                                 \
            for_each_rank(ctx, [&](plssvm_b200_ctx *c, const int g) {                                                                                                            \
                holder->ranks[g] = std::make_unique<cg_session<T>>(c, member_of(X, g), y, kernel, degree, gamma, coef0, cost, eps);                                              \
                PB_CUDA(cudaStreamSynchronize(c->stream));                                                                                                                       \
            });                                                                                                                                                                  \
            ctx->live_sessions.push_back(holder.get());                                                                                                                          \
            handle_add(holder.get());                                                                                                                                            \
            *out = holder.release();                                                                                                                                             \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_cg_trace_##SUF(plssvm_b200_cg *cg, T *out, size_t capacity, size_t *count) {                                                                                \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(cg) && !cg->ranks.empty() && out != nullptr && count != nullptr, "cg session (NULL or no longer alive), out or count is NULL");               \
            PB_REQUIRE(cg->elem_size == static_cast<int>(sizeof(T)), "cg session has the wrong real_type");                                                                      \
            *count = static_cast<cg_session<T> *>(cg->ranks[0].get())->get_trace(out, capacity);                                                                                 \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_cg_finish_##SUF(plssvm_b200_cg *cg, T *alpha_out, T *rho_out, uint64_t *iters_out, T *residual_out) {                                                       \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(cg) && !cg->ranks.empty(), "cg session is NULL or has been finished / aborted / destroyed with its context");                                 \
            struct releaser {                                                                                                                                                    \
                plssvm_b200_cg *h;                                                                                                                                               \
                ~releaser() { cg_release(h); } /* the session is released on every path */                                                                                       \
            } guard{ cg };                                                                                                                                                       \
            PB_REQUIRE(cg->elem_size == static_cast<int>(sizeof(T)), "cg session has the wrong real_type");                                                                      \
            PB_REQUIRE(alpha_out != nullptr && rho_out != nullptr, "alpha_out and rho_out must not be NULL");                                                                    \
            for_each_rank(cg->ctx, [&](plssvm_b200_ctx *, const int g) {                                                                                                         \
                static_cast<cg_session<T> *>(cg->ranks[g].get())->finish(alpha_out, rho_out, iters_out, residual_out, g == 0);                                                   \
            });                                                                                                                                                                  \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_solve_dataset_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const T *y, int kernel, int degree, T gamma, T coef0, T cost, T eps, uint64_t max_iter,   \
                                        T *alpha_out, T *rho_out, uint64_t *iters_out, T *residual_out) {                                                                       \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                       \
            const host_timer ht;                                                                                                                                                 \
            for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                  \
            solve_dataset<T>(ctx, X, y, kernel, degree, gamma, coef0, cost, eps, max_iter, alpha_out, rho_out, iters_out, residual_out);                                         \
            ctx->tm.total_ms = ht.ms();                                                                                                                                          \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    static void solve_host_##SUF(plssvm_b200_ctx *ctx, const host_matrix<T> &X, size_t N, size_t d, const T *y, int kernel, int degree, T gamma, T coef0, T cost, T eps,        \
                                 uint64_t max_iter, T *alpha_out, T *rho_out, uint64_t *iters_out, T *residual_out) {                                                            \
        PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                           \
        PB_REQUIRE(X.valid(), "The data must not be empty!");                                                                                                                    \
        const host_timer ht;                                                                                                                                                     \
        for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                      \
        plssvm_b200_dataset *ds = dataset_create<T>(ctx, X, nullptr, N, d);                                                                                                      \
        try {                                                                                                                                                                    \
            solve_dataset<T>(ctx, ds, y, kernel, degree, gamma, coef0, cost, eps, max_iter, alpha_out, rho_out, iters_out, residual_out);                                        \
        } catch (...) {                                                                                                                                                          \
            dataset_destroy(ds);                                                                                                                                                 \
            throw;                                                                                                                                                               \
        }                                                                                                                                                                        \
        dataset_destroy(ds);                                                                                                                                                     \
        ctx->tm.total_ms = ht.ms();                                                                                                                                              \
    }                                                                                                                                                                            \
    int plssvm_b200_solve_##SUF(plssvm_b200_ctx *ctx, const T *X, size_t N, size_t d, const T *y, int kernel, int degree, T gamma, T coef0, T cost, T eps, uint64_t max_iter,   \
                                T *alpha_out, T *rho_out, uint64_t *iters_out, T *residual_out) {                                                                               \
        return guarded([&] {                                                                                                                                                     \
            solve_host_##SUF(ctx, host_matrix<T>{ X, nullptr, d }, N, d, y, kernel, degree, gamma, coef0, cost, eps, max_iter, alpha_out, rho_out, iters_out, residual_out);     \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_solve_rows_##SUF(plssvm_b200_ctx *ctx, const T *const *rows, size_t N, size_t d, const T *y, int kernel, int degree, T gamma, T coef0, T cost, T eps,       \
                                     uint64_t max_iter, T *alpha_out, T *rho_out, uint64_t *iters_out, T *residual_out) {                                                       \
        return guarded([&] {                                                                                                                                                     \
            solve_host_##SUF(ctx, host_matrix<T>{ nullptr, rows, d }, N, d, y, kernel, degree, gamma, coef0, cost, eps, max_iter, alpha_out, rho_out, iters_out, residual_out);  \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_predict_dataset_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const T *alpha, T rho, T *w_inout, int *w_valid, plssvm_b200_dataset *points,          \
                                          int kernel, int degree, T gamma, T coef0, T *out) {                                                                                   \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx) && handle_live(points), "context or points data set is NULL or has been destroyed");                                                                                        \
            const host_timer ht;                                                                                                                                                 \
            for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                  \
            predict_common<T>(ctx, SV, alpha, rho, w_inout, w_valid, points, host_matrix<T>{}, points->N, kernel, degree, gamma, coef0, out, true);                              \
            ctx->tm.total_ms = ht.ms();                                                                                                                                          \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    static void predict_host_##SUF(plssvm_b200_ctx *ctx, const host_matrix<T> &SV, size_t n_sv, size_t d, const T *alpha, T rho, T *w_inout, int *w_valid,                      \
                                   const host_matrix<T> &points, size_t m, int kernel, int degree, T gamma, T coef0, T *out) {                                                   \
        PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                           \
        PB_REQUIRE(points.valid(), "The data points to predict must not be empty!");                                                                                             \
        const host_timer ht;                                                                                                                                                     \
        for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                      \
        plssvm_b200_dataset *ds = dataset_create<T>(ctx, SV, nullptr, n_sv, d);                                                                                                  \
        try {                                                                                                                                                                    \
            predict_common<T>(ctx, ds, alpha, rho, w_inout, w_valid, nullptr, points, m, kernel, degree, gamma, coef0, out, true);                                               \
        } catch (...) {                                                                                                                                                          \
            dataset_destroy(ds);                                                                                                                                                 \
            throw;                                                                                                                                                               \
        }                                                                                                                                                                        \
        dataset_destroy(ds);                                                                                                                                                     \
        ctx->tm.total_ms = ht.ms();                                                                                                                                              \
    }                                                                                                                                                                            \
    int plssvm_b200_predict_##SUF(plssvm_b200_ctx *ctx, const T *SV, size_t n_sv, size_t d, const T *alpha, T rho, T *w_inout, int *w_valid, const T *points, size_t m,          \
                                  int kernel, int degree, T gamma, T coef0, T *out) {                                                                                           \
        return guarded([&] {                                                                                                                                                     \
            predict_host_##SUF(ctx, host_matrix<T>{ SV, nullptr, d }, n_sv, d, alpha, rho, w_inout, w_valid, host_matrix<T>{ points, nullptr, d }, m, kernel, degree, gamma,     \
                               coef0, out);                                                                                                                                      \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_predict_rows_##SUF(plssvm_b200_ctx *ctx, const T *const *sv_rows, size_t n_sv, size_t d, const T *alpha, T rho, T *w_inout, int *w_valid,                  \
                                       const T *const *point_rows, size_t m, int kernel, int degree, T gamma, T coef0, T *out) {                                                \
        return guarded([&] {                                                                                                                                                     \
            predict_host_##SUF(ctx, host_matrix<T>{ nullptr, sv_rows, d }, n_sv, d, alpha, rho, w_inout, w_valid, host_matrix<T>{ nullptr, point_rows, d }, m, kernel, degree,   \
                               gamma, coef0, out);                                                                                                                               \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_q_kernel_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, int kernel, int degree, T gamma, T coef0, T *q_out, T *k_last) {                               \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                       \
            reset_timings(ctx);                                                                                                                                                  \
            api_q_kernel<T>(ctx, X, kernel, degree, gamma, coef0, q_out, k_last);                                                                                                \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_matvec_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const T *q, const T *v, T QA_cost, T cost_inv, T add, int kernel, int degree, T gamma, T coef0,  \
                                 T *ret_inout) {                                                                                                                                 \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                       \
            for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                  \
            api_matvec<T>(ctx, X, q, v, QA_cost, cost_inv, add, kernel, degree, gamma, coef0, ret_inout);                                                                        \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_w_kernel_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const T *alpha, T *w_out) {                                                                   \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx), "context is NULL or has been destroyed");                                                                                                                       \
            reset_timings(ctx);                                                                                                                                                  \
            api_w_kernel<T>(ctx, SV, alpha, w_out);                                                                                                                              \
        });                                                                                                                                                                      \
    }                                                                                                                                                                            \
    int plssvm_b200_predict_kernel_##SUF(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const T *alpha, plssvm_b200_dataset *points, int kernel, int degree, T gamma, T coef0,  \
                                         T *out) {                                                                                                                               \
        return guarded([&] {                                                                                                                                                     \
            PB_REQUIRE(handle_live(ctx) && handle_live(points), "context or points data set is NULL or has been destroyed");                                                                                        \
            PB_REQUIRE(kernel != PLSSVM_B200_KERNEL_LINEAR, "run_predict_kernel is only defined for the polynomial and rbf kernels (linear uses run_w_kernel)");                 \
            for_all_members(ctx, [](plssvm_b200_ctx *c) { reset_timings(c); });                                                                                                  \
            predict_common<T>(ctx, SV, alpha, T(0), nullptr, nullptr, points, host_matrix<T>{}, points->N, kernel, degree, gamma, coef0, out, false);                            \
        });                                                                                                                                                                      \
    }

PB_INSTANTIATE(f32, float)
PB_INSTANTIATE(f64, double)

}  // extern "C"
