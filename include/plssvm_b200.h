/*
 * plssvm_b200 — C ABI of the Blackwell-native LS-SVM compute backend (libplssvm_b200.so).
 *
 * This is the drop-in boundary for the reference's `detail::gpu_csvm` hot path (PLSSVM v2.0.0).  The reference has no
 * FFI: a backend is a C++ class overriding four protected virtuals of `plssvm::csvm` (include/plssvm/csvm.hpp:188-208).
 * Each entry point below names the reference interface it replaces; the C++ adaptor that turns these calls back into a
 * `plssvm::csvm` subclass is include/plssvm_b200/csvm.hpp, and INTEGRATION.md shows the reference-side registration.
 *
 * Conventions
 *   - plain pointers and sizes only; matrices are dense ROW-MAJOR (one data point per row, `d` contiguous features) —
 *     exactly the rows of the reference's `std::vector<std::vector<real_type>>`
 *   - `_f32` / `_f64` = the reference's `real_type` (float / double); all arithmetic runs in that type
 *   - kernel ids follow `plssvm::kernel_function_type` (include/plssvm/kernel_function_types.hpp:31-38)
 *   - every function returns 0 on success; otherwise `plssvm_b200_last_error()` holds the message the C++ adaptor
 *     rethrows as `plssvm::b200::backend_exception` (reference: CUDA/detail/utility.cu:17-21)
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with PLSSVM_B200_ERR_CUDA
 *   - not re-entrant per context (the reference's csvm is single-threaded by contract)
 */
#ifndef PLSSVM_B200_H_
#define PLSSVM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only the entry points below are exported */
#pragma GCC visibility push(default)

#define PLSSVM_B200_KERNEL_LINEAR 0
#define PLSSVM_B200_KERNEL_POLYNOMIAL 1
#define PLSSVM_B200_KERNEL_RBF 2

#define PLSSVM_B200_OK 0
#define PLSSVM_B200_ERR_INVALID 1 /* bad argument (the reference asserts: gpu_csvm.hpp:484-489, 664-672) */
#define PLSSVM_B200_ERR_CUDA 2    /* CUDA / NCCL runtime failure, or no device */
#define PLSSVM_B200_ERR_INTERNAL 3

typedef struct plssvm_b200_ctx plssvm_b200_ctx;         /* the devices of one process (or one rank of a multi-process run); replaces cuda::csvm::init (CUDA/csvm.cu:48-86) */
typedef struct plssvm_b200_dataset plssvm_b200_dataset; /* a dense matrix resident in HBM; replaces setup_data_on_device (gpu_csvm.hpp:302-346) */

/* device-side timings and CG statistics of the last solve / predict on a context (CUDA events on the compute stream); while a CG
 * session is open they accumulate over its lifetime.  The cg_* fields are what the reference reports through its logger and
 * performance tracker (gpu_csvm.hpp:569-571, 637-646: iterations, residuum, target residuum, average iteration time, epsilon). */
typedef struct plssvm_b200_timings {
    double total_ms;          /* whole call (upload + q + CG + download) measured on the host */
    double cg_loop_ms;        /* the CG iteration loop only (device events), excluding setup and the initial residual */
    double matvec_ms;         /* sum over all implicit matvec launches (tile kernel + partial reduction + all-reduce) */
    double matvec_tile_ms;    /* sum over the tile kernels alone (the dominant kernel) */
    uint64_t matvec_calls;    /* number of implicit matvecs (iterations + 1 + refreshes) */
    uint64_t kernel_launches; /* number of kernels of this library launched by the call */
    double matvec_flops;      /* algorithmic FLOPs of ONE matvec: d * n * (n + 1)   */
    double h2d_bytes;
    double d2h_bytes;
    int impl_used;            /* 1 = SIMT FMA tiles, 2 = floating-point tensor tiles (fp64: TMA + DMMA; fp32: TMA + tcgen05 3xTF32), 3 = factorised linear,
                               * 6 / 7 = int8-slice tcgen05 tiles (fp32: 3 / 4 digit planes), 10 = int8-slice tiles on CTA pairs (the fp32 default
                               * when the A operand has at least 256 rows); experimental builds only: 4 / 5 fp32 3xTF32 variants,
                               * 8 / 9 / 11 int8-slice variants on CTA clusters / CTA pairs (see "impl" below) */
    int n_devices;            /* devices (ranks) that took part in the call */
    uint64_t cg_iterations;     /* min(iter + 1, max_iter) as the reference reports it (gpu_csvm.hpp:639,646) */
    uint64_t cg_max_iterations;
    double cg_residuum;         /* final r.r */
    double cg_target_residuum;  /* eps^2 * r0.r0 */
    double cg_epsilon;
    double cg_avg_iteration_ms; /* host wall time of the iteration loop / iterations (what the reference's avg_iteration_time measures) */
    uint64_t rebalances;        /* several ranks: how often the tile shares were re-cut from the measured tile-kernel rates */
    uint64_t fallback_batches;  /* predict: host-staged batches that were re-run with the floating-point tensor tiles because of badly scaled rows */
    /* option "tile_stats" = 1 (int8-slice tiles only; synchronises after every tile launch): where the warp-specialised roles of the LAST tile
     * launch spent their time, as fractions of the MMA warp's loop, averaged over the CTAs */
    double tile_mma_wait_operands; /* MMA warp waiting for a ring stage to be filled by the TMA producer */
    double tile_mma_wait_drain;    /* MMA warp waiting for the epilogue to hand the TMEM accumulators back */
    double tile_producer_wait;     /* producer waiting for a free ring stage (= operands arrive faster than they are consumed) */
    double tile_epilogue_wait;     /* epilogue waiting for the accumulators of the next unit */
} plssvm_b200_timings;

/* ---- context ----------------------------------------------------------------------------------------------------
 * One context drives `n_dev` GPUs of this process, like the reference's CUDA backend which uses every visible device
 * (src/plssvm/backends/CUDA/csvm.cu:48-86; one host thread per device: include/plssvm/backends/gpu_csvm.hpp:331,369,521,574).
 * device_ids == NULL: devices 0 .. n_dev-1, and n_dev <= 0 then means "all visible devices".  With several devices every
 * data set is replicated (each device uploads 1 / n_dev of the rows over its own PCIe link, NCCL all-gather over NVLink), the
 * tiles of the implicit matvec are sharded over the devices (one ncclAllReduce of the result vector per matvec) and the test
 * points of a predict call are sharded by ranges with no collective.  The API is the same as for one device. */
int plssvm_b200_create(const int *device_ids, int n_dev, plssvm_b200_ctx **out);
int plssvm_b200_destroy(plssvm_b200_ctx *ctx);
/* number of devices the context drives */
int plssvm_b200_num_devices(const plssvm_b200_ctx *ctx, int *count);
/* message of the last failed call on this thread (valid until the next call) */
const char *plssvm_b200_last_error(void);
/* tuning / debugging knobs: "impl" — tile kernel of the implicit matvec / predict contraction: 0 auto (int8-slice tcgen05 tiles;
 * the floating-point tensor tiles for more than 16384 features or badly scaled rows), 1 SIMT FMA tiles, 2 floating-point tensor
 * tiles (fp64: TMA + DMMA, fp32: tcgen05 3xTF32), 6 int8 slices on tcgen05 kind::i8 with exact int32 accumulation (fp64: 7 slices =
 * 54 bits, fp32: 3 slices = 22 bits), 7 the same with 4 slices (30 bits) for fp32, 10 the kernel of 6 on CTA pairs (tcgen05.mma.cta_group::2,
 * tile_i8_pair.cuh; fp64: experimental builds only, otherwise it resolves to 6); only in builds with -DPLSSVM_B200_EXPERIMENTAL
 * (measured-but-not-faster variants kept for reference, bit-identical results): 4 / 5 fp32 3xTF32 variants (CTA pair / 128x256),
 * 8 / 9 / 11 variants of 6 (2 x 2 CTA clusters with TMA multicast / fp32 CTA pairs with cta_group::2 / clusters of two CTAs that share the A planes); "max_ctas" (debugging: cap the
 * number of persistent CTAs of the tile kernels, 0 = one per SM); "check_interval" (CG iterations between host polls),
 * "verbose" (0/1: per-iteration log lines like gpu_csvm.hpp:569-571), "linear_factorized" (0/1: for the linear kernel
 * evaluate Q~ v as X (X^T v) + rank-2 terms — two streaming passes over X, O(n d) instead of O(n^2 d); default 0 = the
 * implicit tiled formulation the reference uses), "ignore_convergence" (0/1, benchmarking only: the stopping test is
 * skipped so that exactly the requested number of CG iterations runs), "balance" (0/1, default 1; several devices / ranks:
 * re-cut the tile shares every "balance_interval" (default 8) iterations in proportion to the tile-kernel rates the ranks
 * measured — GPUs under a power cap do not run at the same clock; 0 = fixed equal shares, bit-reproducible run to run),
 * "shard_upload" (0/1, default 1; several devices / ranks: each uploads 1 / world of the rows, NCCL all-gather);
 * "tile_stats" (0/1, profiling: per-role wait-cycle counters of the int8-slice tile kernel, see plssvm_b200_timings; 2 = additionally, in a library built
 * with -DPB_TILE_STATS_FINE, one JSON line per tile launch on stderr with the split of the epilogue's time), "fp32_fast_drain" (0/1,
 * default 1; A/B switch of the fp32 epilogue: release TMEM before / after the fp64 -> fp32 conversion, bit-identical results),
 * "fp32_pair" (0/1, default 1; automatic kernel choice for fp32: the int8-slice tiles run on CTA pairs (impl 10: tcgen05.mma.cta_group::2, the two
 * tensor cores of a pair share the B operand) instead of single CTAs (impl 6) — bit-identical results, measured 4 - 5 % faster at C3),
 * "i8_a_via_tmem" (0/1, default 0; fp64 int8-slice tiles: the A digit planes that feed two MMAs per step are staged in tensor memory with tcgen05.cp
 * and read by TS-form MMAs — bit-identical, measured 4 % slower, kept as an experiment);
 * testing aid on ONE device: "virtual_world" = G, "virtual_rank" = g make the context compute rank g's share of a G-rank run without
 * a communicator — the matvec returns the partial result of rank g's tiles (the G partial results add up to the full product),
 * predict writes only rank g's range of points; "virtual_skew" = p makes the tile shares unequal (+- p / 2 percent) */
int plssvm_b200_set_option(plssvm_b200_ctx *ctx, const char *key, long long value);
int plssvm_b200_get_timings(const plssvm_b200_ctx *ctx, plssvm_b200_timings *out);
/* residual history of the last finished solve: out[k] = r.r after k iterations (k = 0: r0.r0), at most 4097 entries; what the
 * reference logs per iteration ("Start Iteration {} (max: {}) with current residuum {} ...", gpu_csvm.hpp:569-571) */
int plssvm_b200_last_trace(const plssvm_b200_ctx *ctx, double *out, size_t capacity, size_t *count);
int plssvm_b200_device_count(int *count);
/* 1 if the library was built with -DPLSSVM_B200_EXPERIMENTAL (impl 4 / 5 / 8 / 9 available) */
int plssvm_b200_has_experimental(void);

/* ---- multi-process multi-GPU (one single-device context per process, e.g. under torchrun): the same sharding as a
 * multi-device context, with the NCCL communicator built across processes.  Replaces the reference's host-staged
 * device_reduction (gpu_csvm.hpp:449-475).  `id` is NCCL's 128-byte unique id: rank 0 calls comm_unique_id, the launcher
 * broadcasts it (torch.distributed / MPI / a file), every rank calls comm_init.  Every rank then makes the same calls with
 * the same arguments; predict results are exchanged so that every rank returns all values. */
int plssvm_b200_comm_unique_id(void *id128);
int plssvm_b200_comm_init(plssvm_b200_ctx *ctx, int rank, int world_size, const void *id128);

/* host-only helpers (no CUDA needed): the banded tile schedule shared by kernels, ranks and tests */
uint64_t plssvm_b200_tile_size(void);
uint64_t plssvm_b200_tri_num_tiles(uint64_t tiles_per_side);
uint64_t plssvm_b200_tri_encode(uint64_t tiles_per_side, uint64_t I, uint64_t J);
void plssvm_b200_tri_decode(uint64_t tiles_per_side, uint64_t L, uint32_t *I, uint32_t *J);
void plssvm_b200_rank_range(uint64_t total, int rank, int world_size, uint64_t *lo, uint64_t *hi);
/* the same with shares proportional to `weights[world_size]` (rate-weighted tile shares, option "balance") */
void plssvm_b200_weighted_range(uint64_t total, int rank, int world_size, const double *weights, uint64_t *lo, uint64_t *hi);
/* byte offset of digit `plane` of element (row, feature) in the boxed, pre-swizzled layout of the int8 digit planes: boxes of
 * `box_rows` rows x `slab_bytes` features x `planes` planes, each box the shared-memory image tcgen05.mma reads — K-major rows of 64 bytes with
 * SWIZZLE_64B (fp64 kernel) or of 32 bytes with SWIZZLE_32B (fp32 kernel) */
uint64_t plssvm_b200_i8_plane_offset(uint64_t row, uint32_t feature, uint32_t plane, uint32_t planes, uint32_t box_rows, uint32_t slabs, uint32_t slab_bytes);

/* ---- datasets ----------------------------------------------------------------------------------------------------
 * Upload (or adopt from device memory when src_on_device != 0) a dense row-major N x d matrix.  The library keeps its
 * own padded copy (row pitch rounded up to 128 bytes, zero filled) plus the squared row norms. */
int plssvm_b200_dataset_create_f32(plssvm_b200_ctx *ctx, const float *X, size_t N, size_t d, int src_on_device, plssvm_b200_dataset **out);
int plssvm_b200_dataset_create_f64(plssvm_b200_ctx *ctx, const double *X, size_t N, size_t d, int src_on_device, plssvm_b200_dataset **out);
/* the same from N host row pointers (each d contiguous values): rows are staged through a pinned ring, no flat host copy */
int plssvm_b200_dataset_create_rows_f32(plssvm_b200_ctx *ctx, const float *const *rows, size_t N, size_t d, plssvm_b200_dataset **out);
int plssvm_b200_dataset_create_rows_f64(plssvm_b200_ctx *ctx, const double *const *rows, size_t N, size_t d, plssvm_b200_dataset **out);
int plssvm_b200_dataset_destroy(plssvm_b200_dataset *ds);

/* ---- csvm::solve_system_of_linear_equations (csvm.hpp:188-192; gpu_csvm.hpp:477-654) ------------------------------
 * Solves the reduced LS-SVM system with CG exactly as the reference does (x0 = 1, stop when r.r <= eps^2 r0.r0 or after
 * max_iter iterations, residual refresh every 50th iteration).  X: N x d host matrix, y: N labels as +-1.
 * Outputs: alpha[N] (last entry = -sum of the others), rho (= -bias), the iteration count min(iter + 1, max_iter), and
 * residual[2] = { final r.r, initial r0.r0 } (either may be NULL). */
int plssvm_b200_solve_f32(plssvm_b200_ctx *ctx, const float *X, size_t N, size_t d, const float *y, int kernel, int degree, float gamma, float coef0,
                          float cost, float eps, uint64_t max_iter, float *alpha_out, float *rho_out, uint64_t *iters_out, float *residual_out);
int plssvm_b200_solve_f64(plssvm_b200_ctx *ctx, const double *X, size_t N, size_t d, const double *y, int kernel, int degree, double gamma, double coef0,
                          double cost, double eps, uint64_t max_iter, double *alpha_out, double *rho_out, uint64_t *iters_out, double *residual_out);
/* same with the matrix given as N row pointers (each d contiguous values) — the rows of the reference's
 * std::vector<std::vector<real_type>> are staged straight into a pinned ring, no flat host copy (gpu_csvm.hpp:302-346 transposes
 * on the host and issues one blocking cudaMemcpy) */
int plssvm_b200_solve_rows_f32(plssvm_b200_ctx *ctx, const float *const *rows, size_t N, size_t d, const float *y, int kernel, int degree, float gamma, float coef0,
                               float cost, float eps, uint64_t max_iter, float *alpha_out, float *rho_out, uint64_t *iters_out, float *residual_out);
int plssvm_b200_solve_rows_f64(plssvm_b200_ctx *ctx, const double *const *rows, size_t N, size_t d, const double *y, int kernel, int degree, double gamma, double coef0,
                               double cost, double eps, uint64_t max_iter, double *alpha_out, double *rho_out, uint64_t *iters_out, double *residual_out);
/* same, on a matrix that is already resident in HBM (the timed region of the device-resident benchmark) */
int plssvm_b200_solve_dataset_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const float *y, int kernel, int degree, float gamma, float coef0,
                                  float cost, float eps, uint64_t max_iter, float *alpha_out, float *rho_out, uint64_t *iters_out, float *residual_out);
int plssvm_b200_solve_dataset_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const double *y, int kernel, int degree, double gamma, double coef0,
                                  double cost, double eps, uint64_t max_iter, double *alpha_out, double *rho_out, uint64_t *iters_out, double *residual_out);

/* The same solve as a session, for callers that want to drive / time the CG iterations themselves (bench.py times exactly
 * K iterations with it): begin = upload y, q-kernel, QA_cost, r0 = b~ - Q~ 1, d0 = r0 (gpu_csvm.hpp:505-554);
 * step = enqueue `iterations` CG iterations (gpu_csvm.hpp:568-636) and poll the device-side state once — iterations past
 * convergence are no-ops; finish = bias / alpha_N / download (gpu_csvm.hpp:649-653) and frees the session. */
typedef struct plssvm_b200_cg plssvm_b200_cg;
int plssvm_b200_cg_begin_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const float *y, int kernel, int degree, float gamma, float coef0, float cost, float eps,
                             plssvm_b200_cg **out);
int plssvm_b200_cg_begin_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const double *y, int kernel, int degree, double gamma, double coef0, double cost, double eps,
                             plssvm_b200_cg **out);
int plssvm_b200_cg_step(plssvm_b200_cg *cg, uint64_t iterations, uint64_t *iterations_done, int *converged);
/* residual history: out[k] = r.r after k iterations (k = 0 is r0.r0); *count entries written (<= capacity, <= 4097) */
int plssvm_b200_cg_trace_f32(plssvm_b200_cg *cg, float *out, size_t capacity, size_t *count);
int plssvm_b200_cg_trace_f64(plssvm_b200_cg *cg, double *out, size_t capacity, size_t *count);
int plssvm_b200_cg_finish_f32(plssvm_b200_cg *cg, float *alpha_out, float *rho_out, uint64_t *iters_out, float *residual_out);
int plssvm_b200_cg_finish_f64(plssvm_b200_cg *cg, double *alpha_out, double *rho_out, uint64_t *iters_out, double *residual_out);
int plssvm_b200_cg_abort(plssvm_b200_cg *cg);

/* ---- csvm::predict_values (csvm.hpp:204-208; gpu_csvm.hpp:656-730) --------------------------------------------------
 * out[p] = sum_i alpha_i k(sv_i, point_p) - rho.  Linear kernel: w = sum_i alpha_i sv_i is computed iff *w_valid == 0,
 * stored in w_inout[d] and *w_valid set to 1 (the reference's `w` cache, gpu_csvm.hpp:696-698); other kernels leave
 * w untouched.  Test points are streamed through HBM in batches, 64-bit indexing throughout. */
int plssvm_b200_predict_f32(plssvm_b200_ctx *ctx, const float *SV, size_t n_sv, size_t d, const float *alpha, float rho, float *w_inout, int *w_valid,
                            const float *points, size_t m, int kernel, int degree, float gamma, float coef0, float *out);
int plssvm_b200_predict_f64(plssvm_b200_ctx *ctx, const double *SV, size_t n_sv, size_t d, const double *alpha, double rho, double *w_inout, int *w_valid,
                            const double *points, size_t m, int kernel, int degree, double gamma, double coef0, double *out);
/* support vectors and points as row pointers (see plssvm_b200_solve_rows_*): test points are streamed through the pinned ring in batches */
int plssvm_b200_predict_rows_f32(plssvm_b200_ctx *ctx, const float *const *sv_rows, size_t n_sv, size_t d, const float *alpha, float rho, float *w_inout, int *w_valid,
                                 const float *const *point_rows, size_t m, int kernel, int degree, float gamma, float coef0, float *out);
int plssvm_b200_predict_rows_f64(plssvm_b200_ctx *ctx, const double *const *sv_rows, size_t n_sv, size_t d, const double *alpha, double rho, double *w_inout, int *w_valid,
                                 const double *const *point_rows, size_t m, int kernel, int degree, double gamma, double coef0, double *out);
int plssvm_b200_predict_dataset_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const float *alpha, float rho, float *w_inout, int *w_valid,
                                    plssvm_b200_dataset *points, int kernel, int degree, float gamma, float coef0, float *out);
int plssvm_b200_predict_dataset_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const double *alpha, double rho, double *w_inout, int *w_valid,
                                    plssvm_b200_dataset *points, int kernel, int degree, double gamma, double coef0, double *out);

/* ---- kernel-granular entry points = the reference's four run_*_kernel virtuals (gpu_csvm.hpp:208-277) -------------- */
/* run_q_kernel (csvm.cu:110-129): q[i] = k(x_i, x_{N-1}) for i < N-1; *k_last = k(x_{N-1}, x_{N-1}) (QA_cost = *k_last + 1/C) */
int plssvm_b200_q_kernel_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, int kernel, int degree, float gamma, float coef0, float *q_out, float *k_last);
int plssvm_b200_q_kernel_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, int kernel, int degree, double gamma, double coef0, double *q_out, double *k_last);
/* run_svm_kernel (csvm.cu:134-153): ret[0..N-2] += add * Q~ v with Q~_ij = k(x_i,x_j) + QA_cost - q_i - q_j + delta_ij * cost_inv,
 * add in {+1, -1}, cost_inv = 1 / C — the argument convention of device_kernel_* (svm_kernel.cu:17) */
int plssvm_b200_matvec_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const float *q, const float *v, float QA_cost, float cost_inv, float add,
                           int kernel, int degree, float gamma, float coef0, float *ret_inout);
int plssvm_b200_matvec_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *X, const double *q, const double *v, double QA_cost, double cost_inv, double add,
                           int kernel, int degree, double gamma, double coef0, double *ret_inout);
/* run_w_kernel (csvm.cu:158-165): w[f] = sum_i alpha_i SV[i][f] over ALL n_sv rows */
int plssvm_b200_w_kernel_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const float *alpha, float *w_out);
int plssvm_b200_w_kernel_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const double *alpha, double *w_out);
/* run_predict_kernel (csvm.cu:170-186): out[p] = sum_i alpha_i k(sv_i, point_p)  (no -rho; polynomial / rbf only) */
int plssvm_b200_predict_kernel_f32(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const float *alpha, plssvm_b200_dataset *points, int kernel, int degree,
                                   float gamma, float coef0, float *out);
int plssvm_b200_predict_kernel_f64(plssvm_b200_ctx *ctx, plssvm_b200_dataset *SV, const double *alpha, plssvm_b200_dataset *points, int kernel, int degree,
                                   double gamma, double coef0, double *out);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif
#endif /* PLSSVM_B200_H_ */
