// Tile kernel of the implicit kernel matrix on the 5th-generation tensor cores for BOTH real types: tcgen05.mma kind::i8 over
// int8 slices of X (an Ozaki-style error-free splitting), exact int32 accumulation in TMEM, fp64 recombination in the epilogue.
// fp64: S = 7 slices (54 fractional bits), 128 x 64 units; fp32: S = 3 slices (22 bits, the input precision of the 3xTF32 scheme) or,
// opt-in, S = 4 (30 bits — more than the 24 of the inputs), 128 x 128 units.
//
// Why: tcgen05.mma has no f64 kind and the FP64 pipes of B200 (DMMA == DFMA) stop at ~37 TFLOP/s; the int8 tensor pipe is
// ~120x faster.  Every row x_i is written once per data set (split_i8_kernel) as a fixed-point number relative to its own
// largest element:   x_ik = 2^(e_i - (8S-2)) * sum_p a_p(i,k) 2^(8p),   a_p in [-128, 127]  (balanced base-256 digits),
// i.e. 8S - 2 = 54 fractional bits for S = 7 slices.  Then
//     x_i . x_j = 2^(e_i + e_j - 12) * sum_{t'=0}^{S-1} 2^(-8 (S-1-t')) * ACC_t',      ACC_t' = sum_{p+q = t'+S-1} a_p(i,:) . a_q(j,:)
// where only the S most significant digit diagonals p + q >= S - 1 are kept (S (S+1) / 2 = 28 int8 products instead of S^2;
// the dropped ones are below 2^-51 of |x_i|_max |x_j|_max d).  Each ACC_t' is an EXACT integer (|ACC| < 2^31 for d <= 16384),
// so the only roundings are the S fp64 FMAs that recombine the diagonals: measured error 7e-18 |x_i||x_j| at d = 4096 —
// below that of a native fp64 dot product (4e-17).
//
// Tensor-core mapping (what makes it fast): the S accumulators of a 128 x 64 output block sit side by side in TMEM
// (column t' * 64), and the B slices sit side by side in shared memory in q order, so ONE tcgen05.mma with N = 64 * m multiplies
// slice A_p with the m consecutive slices B_q .. B_(q+m-1) and lands in the m consecutive accumulators t' .. t'+m-1:
// 10 wide MMAs (N up to 256) per 32-feature step instead of 28 narrow ones, which cuts the shared-memory operand reads
// from 168 KB to 96 KB per step (107 B/clk at the tensor pipe's pace; with the 42 KB the producer writes into the ring per step the
// shared-memory port sees 154 B/clk against the 128 B/clk it delivers — that port, not L2 or the tensor pipe, bounds the kernel: DESIGN.md §3.0).
//   * warp 0: producer — per 64-feature slab contiguous bulk copies (cp.async.bulk) out of pre-swizzled operand boxes: A = 128 rows x S planes
//     (56 KB, one copy), B = 64 rows x S planes (28 KB, one 4 KB piece per plane of the column block's box); 2-stage ring (see split_i8_kernel
//     for the layout in HBM)
//   * warp 1: allocates all 512 TMEM columns (S x 64 int32 accumulator columns) and issues the MMAs from one elected lane
//   * warps 2-9: epilogue, one accumulator row and 32 columns per thread: tcgen05.ld, int32 -> fp64 without I2F (exponent
//     trick), Horner recombination, hand TMEM back to the MMA warp, then kernel function, QA_cost - q_i - q_j (+ 1/C on the
//     diagonal), v-weighted row sums and mirrored column sums exactly as in the other tile kernels
// A logical 128 x 128 tile of the schedule (tile_order.hpp) is processed as two 128 x 64 units, so tile ownership, the partial
// buffer and the fixed-order reduction are shared with the DMMA kernel.
// Replaces device_kernel_{linear,polynomial,rbf}<double> (reference svm_kernel.cu:17-222) and device_kernel_predict_*<double>.
#pragma once

#include "common.cuh"

#include <type_traits>
#include "tile_dmma.cuh"  // mbarrier / TMA helpers
#include "tile_tf32.cuh"  // tcgen05 helpers
#include "tile_tf32_2sm.cuh"  // cluster helpers

// How the single-thread roles are entered (see elect_one_sync below): 1 = through elect.sync (tight issue code), 0 = `lane == 0`, -1 = per real
// type (the default).  Measured on one B200, same box, sustained (profiles/r02/ab_elect_roles.txt): the producer gains a little either way; the
// MMA issuer gains 2 - 4 % in the fp32 kernel (4 MMAs per step, the issue path mattered) but LOSES 5 % in the fp64 kernel under the 1 kW cap
// (10 MMAs queued back to back draw more power, the SM clock settles at 1420 instead of 1500 MHz for the same work per clock).
#ifndef PB_ELECT_PRODUCER
    #define PB_ELECT_PRODUCER 1
#endif
#ifndef PB_ELECT_MMA
    #define PB_ELECT_MMA -1
#endif
// how the B planes of one A plane are split into N <= 256 instructions (see the MMA issuer): 1 = as evenly as possible (measured + 0.6 % at C2 on one
// box, bit-identical: profiles/r02/ab_pair_kernel.txt), 0 = greedily
#ifndef PB_I8_EVEN_CHUNKS
    #define PB_I8_EVEN_CHUNKS 1
#endif

namespace pb {

constexpr int I8_BK = 64;                                  // bytes (= features) per slab of the fp64 kernel: one SWIZZLE_64B row, two K = 32 MMA steps
// Slab width of the fp32 kernel (64, or 32 = SWIZZLE_32B rows, one K = 32 step per stage, 9 x 24 KB instead of 4 x 48 KB stages).  The 32-byte
// layout was built to test whether the MMA issuer's operand waits (13.5 % at C3) were TMA latency with too few bytes in flight: they are not —
// with twice the stages the waits go UP (22.6 %: one full / empty handshake and one tcgen05.commit per K step) and the kernel is 1.3 % slower
// (422.6 vs 428.0 TFLOP/s sustained, 485.8 vs 489.3 burst, same box; profiles/r02/ab_fp32_slab_width.txt).  Kept selectable; default 64.
#ifndef PB_I8_BK_F32
    #define PB_I8_BK_F32 64
#endif
// the same for the fp64 kernel (64: 2 x 84 KB stages; 32: 5 x 42 KB)
#ifndef PB_I8_BK_F64
    #define PB_I8_BK_F64 64
#endif
constexpr int I8_THREADS = 320;                            // producer warp, MMA warp, 8 epilogue warps
constexpr int I8_EPI_THREADS = 256;
constexpr std::uint32_t I8_TMEM_COLS = 512;
constexpr std::uint32_t I8_MAX_FEATURES = 16384;           // <= 7 products of <= 2^14 per feature and diagonal stay below 2^31

// per real type: default number of slices S (8 S - 2 fractional bits relative to the row maximum), unit width NH (S * NH <= 512 TMEM
// columns) and the dynamic-range window of the automatic kernel choice (split_i8_kernel).
//   fp64: S = 7 -> 54 bits (an fp64 input has 53)
//   fp32: S = 3 -> 22 bits, the input precision of the 3xTF32 scheme (hi + lo = 2 x 11 bits) at 6 instead of 10 int8 products;
//         S_EXACT = 4 -> 30 bits (every fp32 input within 2^6 of its row maximum is represented exactly), opt-in
template <typename T>
struct I8;
template <>
struct I8<double> {
    static constexpr int S = 7, S_EXACT = 7, NH = 64, AUTO_RANGE = 20, BK = PB_I8_BK_F64;
};
template <>
struct I8<float> {
    static constexpr int S = 3, S_EXACT = 4, NH = 128, AUTO_RANGE = 10, BK = PB_I8_BK_F32;
};
template <typename T, int S_>
struct I8Layout {
    static constexpr int S = S_, NH = I8<T>::NH, BK = I8<T>::BK;
    static constexpr int UNITS = TILE / NH;                     // units per 128 x 128 tile of the schedule
    static constexpr int A_SLICE = TILE * BK;                   // fp64: 8 KiB, fp32: 4 KiB
    static constexpr int B_SLICE = NH * BK;
    static constexpr int A_BYTES = S * A_SLICE;
    static constexpr int B_BYTES = S * B_SLICE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // fp64: 84 KiB, fp32: 48 KiB
    static constexpr int VEC_BYTES = (4 * TILE + 4 * NH + 4 * NH + TILE) * static_cast<int>(sizeof(T));  // row vectors, column vectors, column sums, row sums
    static constexpr int STAGES = (227 * 1024 - 1024 - VEC_BYTES - 256) / STAGE_BYTES;                   // fp64: 2, fp32: 4 (S = 3) / 3 (S = 4)
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + VEC_BYTES + (2 * STAGES + 2) * 8 + 16;
    static constexpr int SLICES_PER_MMA = 256 / NH;             // B slices one N <= 256 instruction covers
    static constexpr int CPT = NH / 2;                          // columns per epilogue thread
    static_assert(S * NH <= 512, "accumulators must fit into TMEM");
    static_assert(STAGE_BYTES % 1024 == 0 && STAGES >= 2 && (BK == 32 || BK == 64), "stage layout");
    static_assert(VEC_BYTES % 8 == 0 && CPT % 32 == 0, "epilogue layout");
};

// ---- operand preparation: rows -> S int8 digit planes + per-row scale ---------------------------------------------------------
// Layout of the planes in HBM ("boxed", pre-swizzled): the operand boxes the tile kernel stages are stored as CONTIGUOUS chunks that already
// are the shared-memory image tcgen05.mma expects (K-major, 64-byte rows, SWIZZLE_64B), so the producer fetches them with plain 1-D bulk
// copies (cp.async.bulk, whole 128-byte lines) instead of 2-D / 3-D tensor boxes whose 64-byte rows halve TMA's request efficiency:
//     box(R, ks) = all S planes of the BR rows [R BR, (R+1) BR) and the 64 features [64 ks, 64 ks + 64):  S x BR x 64 bytes, plane-major;
//     offset(p, r, k) = (((r / BR) num_slabs + k / 64) S + p) BR 64 + (r % BR) 64 + ((((k % 64) / 16) ^ (((r % BR) / 2) % 4)) 16) + k % 16
// ONE copy with BR = 128 serves both operands: the A operand of a tile is a whole box, the B operand of a 128 x NH unit is the row range
// [h NH, (h + 1) NH) of every plane of a box — S contiguous pieces of NH x 64 bytes (fp64: NH = 64, seven 4 KiB pieces; fp32: NH = 128, the whole
// box).  planes_b / br_b: an optional second copy in boxes of br_b rows, written only for the experimental CTA-pair kernel of tile_i8_2sm.cuh.
// Rows are padded to a multiple of 128 and features to a multiple of 64 with zero digits (written by this kernel: launch it over the padded rows).
// rscale[row] = 2^(e_row - 6); one warp per row.
// The products are accurate to ~2^-(8S-2) sqrt(d) |x_i| |x_j| whatever the data (the fixed-point grid is relative to the row maximum,
// which is at most the row norm), but elements far below their row's maximum keep fewer significant bits of their own.
// bad_rows (optional) counts the rows where more than 1 / 16 of the non-zero elements lie more than 2^AUTO_RANGE below the
// largest one; the automatic kernel choice falls back to the floating-point tensor tiles (DMMA / 3xTF32) for such badly scaled data.
// Rows containing inf / NaN get a NaN scale, so they poison their results exactly like native floating-point arithmetic would.
__host__ __device__ __forceinline__ std::size_t i8_boxed_offset(const std::size_t r, const std::uint32_t k, const std::uint32_t p, const std::uint32_t S, const std::uint32_t BR,
                                                                 const std::uint32_t num_slabs, const std::uint32_t BK = 64) {
    // K-major rows of BK bytes; the 16-byte chunks of a row are XOR-swizzled with the address bits the hardware swizzle mode uses:
    // SWIZZLE_64B (BK = 64): bits 4-5 ^= bits 7-8 = (row / 2) % 4;  SWIZZLE_32B (BK = 32): bit 4 ^= bit 7 = (row / 4) % 2
    const std::uint32_t rr = static_cast<std::uint32_t>(r % BR), kk = k % BK;
    const std::uint32_t sw = BK == 64 ? ((rr >> 1) & 3u) : ((rr >> 2) & 1u);
    return (((r / BR) * num_slabs + k / BK) * S + p) * (static_cast<std::size_t>(BR) * BK) + rr * BK + ((((kk >> 4) ^ sw) << 4) | (kk & 15u));
}

// `mean` (optional, ld entries, pad columns zero): the digits are those of x - mean, rounded once to T — the rbf kernel is evaluated on data
// centred at the feature means of the training set / support vectors (exact translation invariance).
template <typename T, int S>
__global__ void __launch_bounds__(256) split_i8_kernel(const T *__restrict__ X, const std::size_t rows, const std::uint32_t d, const std::uint32_t ld,
                                                       std::int8_t *__restrict__ planes_a, std::int8_t *__restrict__ planes_b, const std::uint32_t br_b,
                                                       const std::uint32_t num_slabs, T *__restrict__ rscale, int *__restrict__ bad_rows, const T *__restrict__ mean,
                                                       const std::uint32_t slab_w /* bytes (= features) per slab: 64 or 32 */) {
    const std::size_t row = static_cast<std::size_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (rows + TILE - 1) / TILE * TILE) { return; }
    const bool pad_row = row >= rows;  // padding rows of the last 128-row box: all digits zero
    const int lane = threadIdx.x & 31;
    const T *x = X + (pad_row ? 0 : row) * ld;
    double mx = 0.0, poison = 0.0;
    for (std::uint32_t k = lane; k < (pad_row ? 0u : d); k += 32) {
        const double ax = fabs(static_cast<double>(mean != nullptr ? x[k] - mean[k] : x[k]));
        mx = fmax(mx, ax);
        poison += ax * 0.0;  // NaN iff the row holds an inf or a NaN
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        poison += __shfl_xor_sync(0xffffffffu, poison, o);
    }
    const bool bad = poison != 0.0 || mx > 1.0e300 || pad_row;  // (poison is 0 or NaN)
    if (bad) { mx = 0.0; }
    int e = 0;
    if (mx > 0.0) { (void) frexp(mx, &e); }  // mx = m 2^e, m in [0.5, 1)  =>  |x_k| < 2^e
    e = e < -900 ? -900 : e;
    if (sizeof(T) == 4) { e = e < -100 ? -100 : e; }  // keep the scale a normal float
    const double to_fixed = bad ? 0.0 : ldexp(1.0, (8 * S - 2) - e);
    const double small = ldexp(1.0, e - I8<T>::AUTO_RANGE);
    if (lane == 0 && !pad_row) { rscale[row] = bad ? static_cast<T>(__longlong_as_double(0x7ff8000000000000ll)) : static_cast<T>(ldexp(1.0, e - 6)); }
    unsigned n_nonzero = 0, n_small = 0;
    for (std::uint32_t k0 = 4u * lane; k0 < slab_w * num_slabs; k0 += 128u) {
        long long v[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double xv = (k0 + j < d && !bad) ? static_cast<double>(mean != nullptr ? x[k0 + j] - mean[k0 + j] : x[k0 + j]) : 0.0;
            n_nonzero += xv != 0.0 ? 1u : 0u;
            n_small += (xv != 0.0 && fabs(xv) < small) ? 1u : 0u;
            v[j] = __double2ll_rn(xv * to_fixed);
        }
        #pragma unroll
        for (int p = 0; p < S; ++p) {
            std::uint32_t word = 0;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long a = (p == S - 1) ? v[j] : static_cast<long long>(static_cast<signed char>(v[j] & 0xFF));  // balanced digit in [-128, 127]
                word |= (static_cast<std::uint32_t>(a) & 0xFFu) << (8 * j);
                v[j] = (v[j] - a) >> 8;  // exact
            }
            *reinterpret_cast<std::uint32_t *>(planes_a + i8_boxed_offset(row, k0, p, S, TILE, num_slabs, slab_w)) = word;
            if (planes_b != planes_a) { *reinterpret_cast<std::uint32_t *>(planes_b + i8_boxed_offset(row, k0, p, S, br_b, num_slabs, slab_w)) = word; }
        }
    }
    if (bad_rows != nullptr && !pad_row) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_nonzero += __shfl_xor_sync(0xffffffffu, n_nonzero, o);
            n_small += __shfl_xor_sync(0xffffffffu, n_small, o);
        }
        if (lane == 0 && 16u * n_small > n_nonzero) { atomicAdd(bad_rows, 1); }
    }
}

// ---- tcgen05 kind::i8 helpers -----------------------------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major operand, 64-byte swizzle, 8-row groups 512 bytes apart
__device__ __forceinline__ std::uint64_t umma_desc_sw64(const std::uint32_t smem_addr) {
    return static_cast<std::uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<std::uint64_t>(1) << 16) | (static_cast<std::uint64_t>(512 >> 4) << 32) |
           (static_cast<std::uint64_t>(1) << 46) | (static_cast<std::uint64_t>(4) << 61);
}
// the same for rows of BK bytes: 32-byte swizzle (layout type 6), 8-row groups 256 bytes apart
template <int BK>
__device__ __forceinline__ std::uint64_t umma_desc_kmajor(const std::uint32_t smem_addr) {
    if constexpr (BK == 64) {
        return umma_desc_sw64(smem_addr);
    } else {
        return static_cast<std::uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<std::uint64_t>(1) << 16) | (static_cast<std::uint64_t>(256 >> 4) << 32) |
               (static_cast<std::uint64_t>(1) << 46) | (static_cast<std::uint64_t>(6) << 61);
    }
}
// instruction descriptor: D = S32, A = B = signed int8, both K-major, M = 128, N = n
__host__ __device__ constexpr std::uint32_t i8_idesc(const std::uint32_t n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | (static_cast<std::uint32_t>(TILE >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(const std::uint32_t tmem_d, const std::uint64_t adesc, const std::uint64_t bdesc, const std::uint32_t idesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: D = S32, A = B = signed int8, both K-major, M = 256 (pair), N = n
__host__ __device__ constexpr std::uint32_t i8_idesc_pair(const std::uint32_t n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24); }
// the same product issued for a CTA pair (cta_group::2, M = 256): each CTA supplies its 128 rows of A and HALF of the N rows of B
__device__ __forceinline__ void umma_i8_2sm(const std::uint32_t tmem_d, const std::uint64_t adesc, const std::uint64_t bdesc, const std::uint32_t idesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (TS form): `tmem_a` = 8 columns holding 128 rows x 32 bytes
__device__ __forceinline__ void umma_i8_ts(const std::uint32_t tmem_d, const std::uint32_t tmem_a, const std::uint64_t bdesc, const std::uint32_t idesc, const std::uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory: 128 rows x 256 bits (one K = 32 step of an int8 A operand) described by a matrix descriptor
__device__ __forceinline__ void tmem_cp_128x256b(const std::uint32_t tmem_dst, const std::uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
// One lane of a converged warp (elect.sync).  The single-thread roles branch on THIS predicate, not on `lane == 0`: ptxas then knows that exactly one
// thread runs the region and issues the tcgen05 / bulk-copy instructions (which take uniform registers) directly; behind `lane == 0` it wraps every
// one of them in an ELECT / BRA.U.ANY loop (~6 extra instructions per MMA — enough to make the issuing thread the bottleneck of a 10-MMA step).
__device__ __forceinline__ bool elect_one_sync() {
    std::uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0u;
}
template <bool ELECT>
__device__ __forceinline__ bool role_entered(const int lane) {
    if constexpr (ELECT) {
        return elect_one_sync();
    } else {
        return lane == 0;
    }
}
// contiguous global -> shared bulk copy (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_load(const std::uint32_t dst, const void *src, const std::uint32_t bytes, const std::uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same copy delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in `mask`
__device__ __forceinline__ void bulk_load_mc(const std::uint32_t dst, const void *src, const std::uint32_t bytes, const std::uint32_t bar, const std::uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(const std::uint32_t bar, const std::uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(const std::uint32_t taddr, std::uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// 32 columns of this thread's TMEM lane, no wait (several loads are put in flight before one tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(const std::uint32_t taddr, std::uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// exact int32 -> fp64 on the FP64 add pipe: (2^52 + 2^31 + a) - (2^52 + 2^31)
__device__ __forceinline__ double i32_to_f64(const std::uint32_t a) { return __hiloint2double(0x43300000, static_cast<int>(a ^ 0x80000000u)) - 4503601774854144.0; }

// ---- the epilogue of one unit, shared by the single-CTA kernel below and the CTA-pair kernel (tile_i8_pair.cuh) ---------------------------------
// Called by the 8 epilogue warps (warps 2..9 of the CTA; warp w owns TMEM lanes 32 (w % 4) .. + 31 and columns CPT ch .. + CPT - 1 of the unit) for unit h
// of tile (I, J): stages the row / column vectors, waits for the S accumulators, drains them (tcgen05.ld, int32 -> real), hands TMEM back to the MMA
// warp (PAIR: by a remote arrive on the leader CTA's barrier), evaluates the kernel function and accumulates / stores the v-weighted row sums and the
// mirrored column sums.  `valid` = false: a padding / strictly-upper tile of a super-tile (computed, never stored).
template <typename T, int S, int KERNEL, int MODE, bool PAIR>
__device__ __forceinline__ void i8_epilogue_unit(const TileParams<T> &p, T *s_row, T *s_col, T *s_colsum, T *s_rowsum, const std::uint32_t tmem_base, const std::uint32_t tfull,
                                                 const std::uint32_t tempty, const std::uint32_t unit_iter, const std::uint32_t I, const std::uint32_t J, const int h,
                                                 const bool valid, const bool diag, const T qa, T &rowacc) {
    constexpr int NH = I8<T>::NH, UNITS = TILE / NH, CPT = NH / 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3;
    const int ch = (warp - 2) >> 2;        // column half of the unit
    const int row = quarter * 32 + lane;   // accumulator row of this thread
    const int et = tid - 64;               // 0..255 among the epilogue threads
    const std::uint32_t row0 = I * TILE;
    // profiling (option "tile_stats"): where ONE epilogue thread spends its cycles — stats[4] waiting for the accumulators and, in builds with
    // -DPB_TILE_STATS_FINE, [5] drain (tcgen05.ld + fold / Horner + conversion), [7] vector loads, named barriers, partial stores; the remainder of the MMA
    // issuer's loop time is the kernel function + sums.  Differences of clock64() are accumulated as "- start" / "+ end" (mod 2^64), so that no time
    // stamp stays live across the register-heavy phases.
    auto stamp = [&](const int minus, const int plus) {
#ifdef PB_TILE_STATS_FINE  // (the extra stamps cost the fp32 kernels 1 - 2 %: profiling builds only, -DPB_TILE_STATS_FINE)
        if (p.stats != nullptr && tid == 64) {
            const unsigned long long now = static_cast<unsigned long long>(clock64());
            if (minus >= 0) { atomicAdd(p.stats + blockIdx.x * 8 + minus, 0ull - now); }
            if (plus >= 0) { atomicAdd(p.stats + blockIdx.x * 8 + plus, now); }
        }
#else
        (void) minus;
        (void) plus;
#endif
    };
    auto hand_back = [&]() {
        if constexpr (PAIR) {
            mbar_arrive_cluster(tempty, 0u);
        } else {
            mbar_arrive(tempty);
        }
    };
    const std::uint32_t col0 = J * TILE + h * NH;
    stamp(7, -1);
    if (h == 0 && et < TILE) {
        const std::uint32_t gi = row0 + et;
        const bool oki = gi < p.n_rows;
        s_row[0 * TILE + et] = (MODE == MODE_SYM && oki) ? p.q[gi] : T(0);
        s_row[1 * TILE + et] = (MODE == MODE_SYM && oki) ? p.v[gi] : T(0);
        s_row[2 * TILE + et] = (KERNEL == K_RBF && oki) ? p.row_sq[gi] : T(0);
        s_row[3 * TILE + et] = oki ? p.A_scale[gi] : T(0);
    }
    if (et >= TILE && et < TILE + NH) {
        const int c = et - TILE;
        const std::uint32_t gj = col0 + c;
        const bool okj = gj < p.n_cols;
        s_col[0 * NH + c] = (MODE == MODE_SYM && okj) ? p.q[gj] : T(0);
        s_col[1 * NH + c] = okj ? p.v[gj] : T(0);
        s_col[2 * NH + c] = (KERNEL == K_RBF && okj) ? p.col_sq[gj] : T(0);
        s_col[3 * NH + c] = okj ? p.B_scale[gj] : T(0);
    }
    named_bar_sync(1, I8_EPI_THREADS);
    const T qi = s_row[0 * TILE + row], vi = s_row[1 * TILE + row], sqi = s_row[2 * TILE + row];
    // fp32, S = 3, d <= 1984: the two upper diagonals fit one int32 (|ACC_2| 2^8 + |ACC_1| <= d (2^20 + 2^14) < 2^31), so an element is two
    // words (fast drain below) and its recombination one IMAD + 3 instead of 5 fp64 operations; the sum is 256 x the value, folded
    // into the row scale.  Both forms are exact in fp64: bit-identical results.
    const bool fold3 = sizeof(T) == 4 && S == 3 && p.ld8 <= 1984u && p.slow_drain == 0;
    const T sci = fold3 ? s_row[3 * TILE + row] * T(0.00390625) : s_row[3 * TILE + row];

    stamp(-1, 7);  // vector loads + named barrier before the wait
    const long long c2 = (p.stats != nullptr && tid == 64) ? clock64() : 0;
    mbar_wait(tfull, unit_iter & 1u);
    if (p.stats != nullptr && tid == 64) { atomicAdd(p.stats + blockIdx.x * 8 + 4, static_cast<unsigned long long>(clock64() - c2)); }
    stamp(5, -1);
    tcgen05_fence_after();
    const std::uint32_t taddr = tmem_base + (static_cast<std::uint32_t>(quarter * 32) << 16) + static_cast<std::uint32_t>(ch * CPT);

    // phase 1: S int32 diagonals -> one value per element (Horner in fp64 from the least significant diagonal: every step exact
    // up to one rounding relative to the running sum)
    T a[CPT];
    bool released = false;
    if constexpr (sizeof(T) == 4 && S == 3 && CPT == 64) {
        if (fold3) {
            // fp32 fast drain (measured with tools/tmem_probe: the fp64 -> fp32 conversion runs at ~12 elements / clk / SM and a
            // tcgen05.ld + wait round trip costs a few hundred cycles, so converting while the accumulators are still held kept the
            // tensor pipe idle for ~1/3 of a d = 1024 unit).  Here the raw int32 diagonals of all 64 columns are pulled into
            // registers with two rounds of three 32-column loads, folded to two words per element by integer arithmetic, and TMEM
            // goes back to the MMA warp BEFORE any floating-point work; the conversion overlaps the next unit's MMAs.
            std::uint32_t hi[CPT], lo[CPT];
            #pragma unroll
            for (int half = 0; half < 2; ++half) {
                std::uint32_t r0[32], r1[32], r2[32];
                tmem_ld_32x32b_x32_nowait(taddr + static_cast<std::uint32_t>(0 * NH + half * 32), r0);
                tmem_ld_32x32b_x32_nowait(taddr + static_cast<std::uint32_t>(1 * NH + half * 32), r1);
                tmem_ld_32x32b_x32_nowait(taddr + static_cast<std::uint32_t>(2 * NH + half * 32), r2);
                tmem_ld_wait();
                #pragma unroll
                for (int j = 0; j < 32; ++j) {
                    hi[half * 32 + j] = r2[j] * 256u + r1[j];
                    lo[half * 32 + j] = r0[j];
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { hand_back(); }
            released = true;
            #pragma unroll
            for (int j = 0; j < CPT; ++j) { a[j] = static_cast<T>(fma(i32_to_f64(lo[j]), 0.00390625, i32_to_f64(hi[j]))); }
        }
    }
    if (!released) {
        #pragma unroll
        for (int g = 0; g < CPT / 8; ++g) {
            std::uint32_t r[S][8];
            #pragma unroll
            for (int t = 0; t < S; ++t) { tmem_ld_32x32b_x8(taddr + static_cast<std::uint32_t>(t * NH + g * 8), r[t]); }
            tmem_ld_wait();
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                double s = i32_to_f64(r[0][j]);
                #pragma unroll
                for (int t = 1; t < S; ++t) { s = fma(s, 0.00390625, i32_to_f64(r[t][j])); }
                a[g * 8 + j] = static_cast<T>(s);
            }
        }
        // all of this warp's accumulator reads are done: hand TMEM back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { hand_back(); }
    }

    stamp(-1, 5);  // accumulators read (and converted)
    // phase 2: kernel function and the weighted sums.  The loop is instantiated per (polynomial degree, diagonal tile) so that the power
    // is straight-line code and off-diagonal tiles carry no diagonal test (at d = 1024 the fp32 epilogue, not the tensor pipe, paces
    // the kernel: profiles/r02/tile_role_stats.jsonl); the column vectors come out of shared memory as 16-byte vectors.  Operation
    // order and rounding are unchanged.
    auto phase2 = [&](auto deg_tag, auto diag_tag) {
        constexpr int DEG = decltype(deg_tag)::value;
        constexpr bool DIAG = decltype(diag_tag)::value;
        constexpr int V = 16 / static_cast<int>(sizeof(T));  // elements per 16-byte shared-memory load
        struct alignas(16) vec {
            T x[V];
        };
        #pragma unroll
        for (int j0 = 0; j0 < CPT; j0 += V) {
            const int c0 = ch * CPT + j0;
            const vec scj = *reinterpret_cast<const vec *>(s_col + 3 * NH + c0);
            const vec vj = *reinterpret_cast<const vec *>(s_col + 1 * NH + c0);
            vec sqj{}, qj{};
            if constexpr (KERNEL == K_RBF) { sqj = *reinterpret_cast<const vec *>(s_col + 2 * NH + c0); }
            if constexpr (MODE == MODE_SYM) { qj = *reinterpret_cast<const vec *>(s_col + 0 * NH + c0); }
            #pragma unroll
            for (int u = 0; u < V; ++u) {
                const int j = j0 + u;
                const T dot = a[j] * (sci * scj.x[u]);
                const T kv = kernel_from_dot<KERNEL, T, DEG>(dot, sqi, sqj.x[u], p.kp);
                T t = kv;
                if constexpr (MODE == MODE_SYM) {
                    t = kv + qa - qi - qj.x[u];
                    if constexpr (DIAG) {
                        if (row == h * NH + c0 + u) { t += p.cost_inv; }
                    }
                }
                rowacc = pb_fma(t, vj.x[u], rowacc);
                a[j] = t * vi;  // mirrored contribution of this row to column c0 + u
            }
        }
    };
    auto phase2_deg = [&](auto deg_tag) {
        if (diag) {
            phase2(deg_tag, std::true_type{});
        } else {
            phase2(deg_tag, std::false_type{});
        }
    };
    if constexpr (KERNEL == K_POLYNOMIAL) {
        switch (p.kp.degree) {  // CTA-uniform
            case 2: phase2_deg(std::integral_constant<int, 2>{}); break;
            case 3: phase2_deg(std::integral_constant<int, 3>{}); break;
            default: phase2_deg(std::integral_constant<int, 0>{}); break;
        }
    } else {
        phase2_deg(std::integral_constant<int, 0>{});
    }
    if constexpr (MODE == MODE_SYM) {
        if (!diag) {  // CTA-uniform
            // butterfly per 32 columns: after 5 halving steps lane c holds the sum over the warp's 32 rows of column c
            #pragma unroll
            for (int cc = 0; cc < CPT / 32; ++cc) {
                #pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const bool upper = (lane & step) != 0;
                    #pragma unroll
                    for (int k = 0; k < step; ++k) {
                        const T send = upper ? a[cc * 32 + k] : a[cc * 32 + k + step];
                        const T keep = upper ? a[cc * 32 + k + step] : a[cc * 32 + k];
                        a[cc * 32 + k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                    }
                }
                s_colsum[quarter * NH + ch * CPT + cc * 32 + lane] = a[cc * 32];
            }
        }
    }
    stamp(7, -1);
    if (h == UNITS - 1 && ch == 1) { s_rowsum[row] = rowacc; }
    named_bar_sync(1, I8_EPI_THREADS);  // column sums of the four row quarters / row sums of the second column half visible
    if constexpr (MODE == MODE_SYM) {
        if (!diag && valid && et < NH) {
            const T s = ((s_colsum[et] + s_colsum[NH + et]) + s_colsum[2 * NH + et]) + s_colsum[3 * NH + et];
            const std::size_t mslot = static_cast<std::size_t>(J) * p.T_cols + I;
            p.partial[mslot * TILE + h * NH + et] = (col0 + et < p.n_cols) ? s : T(0);
        }
    }
    if (h == UNITS - 1 && ch == 0 && valid) {
        const std::size_t slot = static_cast<std::size_t>(I) * p.T_cols + J;
        p.partial[slot * TILE + row] = (row0 + row < p.n_rows) ? rowacc + s_rowsum[row] : T(0);
    }
    named_bar_sync(1, I8_EPI_THREADS);  // s_col / s_colsum / s_rowsum consumed before the next unit overwrites them
    stamp(-1, 7);  // named barriers + partial stores after the sums
}

// CL = 1: one CTA per tile.  CL = 4: a cluster of 2 x 2 CTAs per 256 x 256 super-tile (launched with cluster dimension 4; the tile range is in
// super-tiles like the CTA-pair 3xTF32 kernel's): CTA (r, c) computes tile (2 I2 + r, 2 J2 + c); the two CTAs of a cluster row need the same A
// planes and the two of a cluster column the same B planes, so every CTA fetches only one half of the planes of its A block and of its B block and
// TMA multicasts it to its mate — half the L2 -> SM traffic per CTA (the single-CTA kernel runs at 92 % of the L2 throughput cap).  A stage is
// refilled once the MMA warps of all three CTAs that write into or read from it have released it (multicast tcgen05.commit, count 3).
// CL = 2: a cluster of two CTAs per super-tile that share only the A planes: CTA c computes the tiles (2 I2, 2 J2 + c) and (2 I2 + 1, 2 J2 + c) one after
// the other, each CTA fetches one half of the A planes of the current row block and multicasts it to its mate (a third less L2 -> SM traffic per CTA), the B
// planes are fetched per CTA as on single CTAs; a stage is refilled once both MMA warps have released it (count 2).
template <typename T, int S_, int KERNEL, int MODE, int CL, bool TS = false>
__global__ void __launch_bounds__(I8_THREADS, 1)  // (10 warps = 3 on one SM sub-partition: 170 registers per thread at most)
tile_kernel_i8(const TileParams<T> p) {
    using L8 = I8Layout<T, S_>;
    static_assert(CL == 1 || CL == 2 || CL == 4, "cluster size");
    constexpr int TPW = CL == 2 ? 2 : 1;  // tiles per work item and CTA
    constexpr int S = L8::S, NH = L8::NH, UNITS = L8::UNITS, STAGES = L8::STAGES;
    extern __shared__ unsigned char smem_raw[];
    if (p.done != nullptr && *p.done != 0) { return; }

    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *stages = smem;
    T *s_row = reinterpret_cast<T *>(smem + STAGES * L8::STAGE_BYTES);            // [4][TILE]: q_i, v_i, sq_i, scale_i
    T *s_col = s_row + 4 * TILE;                                                   // [4][NH]: q_j, v_j, sq_j, scale_j
    T *s_colsum = s_col + 4 * NH;                                                  // [4][NH]
    T *s_rowsum = s_colsum + 4 * NH;                                               // [TILE]
    std::uint64_t *bars = reinterpret_cast<std::uint64_t *>(s_rowsum + TILE);      // full[STAGES], empty[STAGES], tmem_full, tmem_empty
    std::uint32_t *tmem_slot = reinterpret_cast<std::uint32_t *>(bars + 2 * STAGES + 2);
    const std::uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const std::uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const std::uint32_t num_slabs = p.ld8 / L8::BK;
    std::uint32_t crank = 0;
    if constexpr (CL != 1) { crank = cluster_ctarank(); }
    const std::uint32_t cr = crank >> 1, cc = crank & 1u;  // CL = 4: position of this CTA inside the 2 x 2 cluster; CL = 2: cc = its tile column, cr = 0
    const std::uint64_t work_first = blockIdx.x / CL, work_stride = gridDim.x / CL;
    const std::uint32_t S_rows = (p.T_rows + 1) >> 1, S_cols = (p.T_cols + 1) >> 1;
    // work item L -> tile (I, J) of this CTA; `valid` = false for the padding / strictly-upper tile of a super-tile (computed, never stored)
    // (rr: CL = 2 only, the tile row inside the super-tile this CTA is working on)
    auto decode = [&](const std::uint64_t L, const std::uint32_t rr, std::uint32_t &I, std::uint32_t &J) -> bool {
        if constexpr (CL == 1) {
            if constexpr (MODE == MODE_SYM) {
                tri_decode(p.T_rows, L, I, J);
            } else {
                rect_decode(p.T_rows, p.T_cols, L, I, J);
            }
            return true;
        } else {
            std::uint32_t I2, J2;
            if constexpr (MODE == MODE_SYM) {
                tri_decode(S_rows, L, I2, J2);
            } else {
                rect_decode(S_rows, S_cols, L, I2, J2);
            }
            I = 2 * I2 + (CL == 4 ? cr : rr);
            J = 2 * J2 + cc;
            return I < p.T_rows && J < p.T_cols && (MODE == MODE_RECT || J <= I);
        }
    };

    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CL == 4 ? 3 : CL);  // CL = 4: this CTA's MMA warp, its row mate's and its column mate's; CL = 2: both MMA warps
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, I8_EPI_THREADS / 32);  // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(I8_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    if constexpr (CL != 1) { cluster_sync_all(); }  // the mbarriers of all CTAs of the cluster exist before anyone multicasts into them
    __syncthreads();
    tcgen05_fence_after();
    const std::uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (role_entered<PB_ELECT_PRODUCER != 0>(lane)) {
            std::uint32_t stage = 0, phase = 0;
            long long w_empty = 0;  // cycles spent waiting for a free ring stage (p.stats)
            const std::uint16_t mask_a = static_cast<std::uint16_t>(3u << (2 * cr));                // the two CTAs of this cluster row (CL = 2: both CTAs)
            const std::uint16_t mask_b = static_cast<std::uint16_t>((1u << cc) | (1u << (2 + cc)));  // the two CTAs of this cluster column (CL = 4)
            for (std::uint64_t L = p.tile_lo + work_first; L < p.tile_hi; L += work_stride) {
              for (std::uint32_t rr = 0; rr < TPW; ++rr) {
                std::uint32_t I, J;
                decode(L, rr, I, J);
                // padding tiles of a super-tile (CL = 4, odd tile counts) load the last valid block instead of running past the buffers
                const std::uint32_t Il = I < p.T_rows ? I : p.T_rows - 1, Jl = J < p.T_cols ? J : p.T_cols - 1;
                for (int h = 0; h < UNITS; ++h) {
                    // ONE copy of the digit planes serves both operands: boxes of 128 rows, plane-major.  The B operand of a unit is the
                    // row half h of every plane of column block J's box — S pieces of NH x 64 bytes (whole 128-byte lines) when NH < 128
                    const std::int8_t *src_a = p.A_i8 + static_cast<std::size_t>(Il) * num_slabs * L8::A_BYTES;
                    const std::int8_t *src_b = p.B_i8 + static_cast<std::size_t>(Jl) * num_slabs * L8::A_BYTES + h * L8::B_SLICE;
                    for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                        const long long c0 = p.stats != nullptr ? clock64() : 0;
                        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                        if (p.stats != nullptr) { w_empty += clock64() - c0; }
                        const std::uint32_t dst = smem_u32(stages + stage * L8::STAGE_BYTES);
                        const std::uint32_t bar = full0 + 8 * stage;
                        mbar_arrive_expect_tx(bar, L8::STAGE_BYTES);
                        const std::int8_t *box_a = src_a + static_cast<std::size_t>(ks) * L8::A_BYTES, *box_b = src_b + static_cast<std::size_t>(ks) * L8::A_BYTES;
                        if constexpr (CL == 1) {
                            bulk_load(dst, box_a, L8::A_BYTES, bar);
                            if constexpr (UNITS == 1) {
                                bulk_load(dst + L8::A_BYTES, box_b, L8::B_BYTES, bar);
                            } else {
                                #pragma unroll
                                for (int pl = 0; pl < S; ++pl) { bulk_load(dst + L8::A_BYTES + pl * L8::B_SLICE, box_b + pl * L8::A_SLICE, L8::B_SLICE, bar); }
                            }
                        } else {
                            // this CTA fetches one contiguous half of the planes of its row block / column block for itself and its mate
                            constexpr int S0 = (S + 1) / 2;
                            const int pa0 = cc == 0 ? 0 : S0, pa1 = cc == 0 ? S0 : S, pb0 = cr == 0 ? 0 : S0, pb1 = cr == 0 ? S0 : S;
                            bulk_load_mc(dst + pa0 * L8::A_SLICE, box_a + pa0 * L8::A_SLICE, static_cast<std::uint32_t>((pa1 - pa0) * L8::A_SLICE), bar, mask_a);
                            if constexpr (CL == 2) {  // B per CTA, as on single CTAs
                                if constexpr (UNITS == 1) {
                                    bulk_load(dst + L8::A_BYTES, box_b, L8::B_BYTES, bar);
                                } else {
                                    #pragma unroll
                                    for (int pl = 0; pl < S; ++pl) { bulk_load(dst + L8::A_BYTES + pl * L8::B_SLICE, box_b + pl * L8::A_SLICE, L8::B_SLICE, bar); }
                                }
                            } else if constexpr (UNITS == 1) {
                                bulk_load_mc(dst + L8::A_BYTES + pb0 * L8::B_SLICE, box_b + pb0 * L8::B_SLICE, static_cast<std::uint32_t>((pb1 - pb0) * L8::B_SLICE), bar, mask_b);
                            } else {
                                for (int pl = pb0; pl < pb1; ++pl) { bulk_load_mc(dst + L8::A_BYTES + pl * L8::B_SLICE, box_b + pl * L8::A_SLICE, L8::B_SLICE, bar, mask_b); }
                            }
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
              }
            }
            if (p.stats != nullptr) { p.stats[blockIdx.x * 8 + 3] = static_cast<unsigned long long>(w_empty); }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (role_entered<(PB_ELECT_MMA < 0 ? sizeof(T) == 4 : PB_ELECT_MMA != 0)>(lane)) {
            std::uint32_t stage = 0, phase = 0, unit_iter = 0;
            std::uint32_t kstep_parity = 0;
            long long w_full = 0, w_tempty = 0;  // cycles waiting for operands / for the epilogue to hand the accumulators back (p.stats)
            const long long c_begin = p.stats != nullptr ? clock64() : 0;
            const std::uint16_t mask_rel = CL == 2 ? static_cast<std::uint16_t>(3u)
                                                   : static_cast<std::uint16_t>((1u << crank) | (1u << (crank ^ 1u)) | (1u << (crank ^ 2u)));  // this CTA, its row mate, its column mate
            for (std::uint64_t L = p.tile_lo + work_first; L < p.tile_hi; L += work_stride) {
                for (int h = 0; h < TPW * UNITS; ++h, ++unit_iter) {  // (CL = 2: the units of both tiles of the work item)
                    const long long c0 = p.stats != nullptr ? clock64() : 0;
                    mbar_wait(tempty, (unit_iter & 1u) ^ 1u);  // epilogue has drained the accumulators of the previous unit
                    if (p.stats != nullptr) { w_tempty += clock64() - c0; }
                    tcgen05_fence_after();
                    for (std::uint32_t ks = 0; ks < num_slabs; ++ks) {
                        const long long c1 = p.stats != nullptr ? clock64() : 0;
                        mbar_wait(full0 + 8 * stage, phase);
                        if (p.stats != nullptr) { w_full += clock64() - c1; }
                        tcgen05_fence_after();
                        const std::uint32_t base = smem_u32(stages + stage * L8::STAGE_BYTES);
                        const std::uint64_t d_a = umma_desc_kmajor<L8::BK>(base), d_b = umma_desc_kmajor<L8::BK>(base + L8::A_BYTES);
                        #pragma unroll
                        for (std::uint32_t k = 0; k < L8::BK / 32; ++k) {
                            const std::uint64_t koff = static_cast<std::uint64_t>((k * 32) >> 4);  // 32 bytes per K = 32 step inside the swizzle atom
                            const bool first = (ks | k) == 0u;
                            // Experimental (option "i8_a_via_tmem", fp64 only): the A planes that feed TWO instructions per step (p = 4, 5, 6: 5 - 7 B planes
                            // = two N <= 256 chunks) are copied shared -> tensor memory once (tcgen05.cp, 24 of the 64 spare columns, double-buffered
                            // over the steps) and read from there by both instructions — one shared-memory read of the plane instead of two.
                            constexpr bool a_tmem = TS;
                            std::uint32_t a_buf = 0;
                            if constexpr (TS) {
                                static_assert(!TS || (sizeof(T) == 8 && S == 7 && CL == 1), "A planes through tensor memory: fp64 kernel only");
                                a_buf = tmem_base + static_cast<std::uint32_t>(S * NH) + (kstep_parity ? 24u : 0u);
                                kstep_parity ^= 1u;
                                #pragma unroll
                                for (int pl = S - 1; pl >= 4; --pl) {
                                    tmem_cp_128x256b(a_buf + static_cast<std::uint32_t>((pl - 4) * 8), d_a + koff + static_cast<std::uint64_t>((pl * L8::A_SLICE) >> 4));
                                }
                            }
                            // slice A_p times the slices B_q, q = S-1-p .. S-1, lands in the accumulators t' = 0 .. p (N <= 256 per instruction)
                            #pragma unroll
                            for (int pp = S - 1; pp >= 0; --pp) {
                                const int q_lo = S - 1 - pp, cnt = pp + 1;
                                // the cnt planes go into ceil(cnt / SLICES_PER_MMA) instructions: PB_I8_EVEN_CHUNKS = 0 fills them greedily (fp64: 4 + 3, 4 + 2, 4 + 1),
                                // 1 splits them as evenly as possible (4 + 3, 3 + 3, 3 + 2: no N = 64 instruction — the shape that starves on the shared-memory port — next to a full one)
                                constexpr int SPM = L8::SLICES_PER_MMA;
                                const int nchunks = (cnt + SPM - 1) / SPM, base_sz = PB_I8_EVEN_CHUNKS ? cnt / nchunks : SPM, rem = PB_I8_EVEN_CHUNKS ? cnt % nchunks : 0;
                                #pragma unroll
                                for (int c = 0; c < nchunks; ++c) {
                                    const int start = c * base_sz + (c < rem ? c : rem);
                                    const int nsl = PB_I8_EVEN_CHUNKS ? base_sz + (c < rem ? 1 : 0) : (cnt - start < SPM ? cnt - start : SPM);
                                    const std::uint32_t d_acc = tmem_base + static_cast<std::uint32_t>(start * NH);
                                    const std::uint64_t bdesc = d_b + koff + static_cast<std::uint64_t>(((q_lo + start) * L8::B_SLICE) >> 4);
                                    const std::uint32_t idesc = i8_idesc(static_cast<std::uint32_t>(nsl * NH)), acc = (first && pp == S - 1) ? 0u : 1u;
                                    if (a_tmem && pp >= 4) {  // (compile-time after unrolling)
                                        umma_i8_ts(d_acc, a_buf + static_cast<std::uint32_t>((pp - 4) * 8), bdesc, idesc, acc);
                                    } else {
                                        umma_i8(d_acc, d_a + koff + static_cast<std::uint64_t>((pp * L8::A_SLICE) >> 4), bdesc, idesc, acc);
                                    }
                                }
                            }
                        }
                        // smem stage reusable once these MMAs have read it
                        if constexpr (CL == 1) {
                            umma_commit(empty0 + 8 * stage);
                        } else {
                            umma_commit_mc(empty0 + 8 * stage, mask_rel);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit(tfull);  // all S accumulators of this unit complete
                }
            }
            if (p.stats != nullptr) {
                p.stats[blockIdx.x * 8 + 0] = static_cast<unsigned long long>(clock64() - c_begin);
                p.stats[blockIdx.x * 8 + 1] = static_cast<unsigned long long>(w_full);
                p.stats[blockIdx.x * 8 + 2] = static_cast<unsigned long long>(w_tempty);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..9; warp w owns TMEM lanes 32 (w % 4) .. + 31 and columns CPT ch .. + CPT - 1 of the unit =====
        std::uint32_t unit_iter = 0;
        for (std::uint64_t L = p.tile_lo + work_first; L < p.tile_hi; L += work_stride) {
            for (std::uint32_t rr = 0; rr < TPW; ++rr) {
                std::uint32_t I, J;
                const bool valid = decode(L, rr, I, J);
                const bool diag = (MODE == MODE_SYM) && (I == J);
                const T qa = (MODE == MODE_SYM) ? *p.QA_cost : T(0);
                T rowacc = T(0);
                for (int h = 0; h < UNITS; ++h, ++unit_iter) {
                    i8_epilogue_unit<T, S, KERNEL, MODE, false>(p, s_row, s_col, s_colsum, s_rowsum, tmem_base, tfull, tempty, unit_iter, I, J, h, valid, diag, qa, rowacc);
                }
            }
        }
    }

    // teardown: everyone done with TMEM before the allocating warp frees it (CL = 4: and with each other's shared memory and mbarriers)
    tcgen05_fence_before();
    if constexpr (CL != 1) { cluster_sync_all(); }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

}  // namespace pb
