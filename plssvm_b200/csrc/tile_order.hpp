// Tile schedules for the implicit kernel matrix (host + device).
//
// The n x n matrix Q~ is cut into TILE x TILE tiles.  The training matvec only visits the lower triangle (I >= J) and
// mirrors every off-diagonal tile (the reference does the same per 96x96 block: svm_kernel.cu:26,74,85); prediction
// visits a rectangle (points x support vectors).  Tiles are enumerated in a *banded* order: bands of BAND tile rows,
// column-major inside a band, so the ~148 tiles that are in flight at any time share ~BAND row blocks and ~148/BAND
// column blocks of X in L2.  A rank of a multi-GPU run owns one contiguous range of this order, which balances the
// triangle exactly — the functions below are the single source of truth for "who computes tile (I, J)".
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
    #define PB_HD __host__ __device__ __forceinline__
#else
    #define PB_HD inline
#endif

namespace pb {

constexpr int TILE = 128;  // rows / columns of the implicit matrix per tile
constexpr int BAND = 12;   // tile rows per band

PB_HD std::uint64_t tri_num_tiles(const std::uint64_t T) { return T * (T + 1) / 2; }

// (I, J) with I >= J  ->  position in the banded lower-triangle order
PB_HD std::uint64_t tri_encode(const std::uint64_t T, const std::uint64_t I, const std::uint64_t J) {
    const std::uint64_t R0 = (I / BAND) * BAND;
    const std::uint64_t hh = (T - R0 < static_cast<std::uint64_t>(BAND)) ? (T - R0) : static_cast<std::uint64_t>(BAND);
    const std::uint64_t base = R0 * (R0 + 1) / 2;
    if (J < R0) {
        return base + J * hh + (I - R0);
    }
    const std::uint64_t jp = J - R0;
    // columns 0..jp-1 of the band's triangular end hold hh, hh-1, ... tiles
    return base + R0 * hh + jp * hh - jp * (jp - 1) / 2 - jp + (I - R0);
}

PB_HD void tri_decode(const std::uint64_t T, const std::uint64_t L, std::uint32_t &I, std::uint32_t &J) {
    // band: largest R0 = BAND * b with R0 (R0 + 1) / 2 <= L
    std::uint64_t R = static_cast<std::uint64_t>((sqrt(8.0 * static_cast<double>(L) + 1.0) - 1.0) * 0.5);
    while (R * (R + 1) / 2 > L) { --R; }
    while ((R + 1) * (R + 2) / 2 <= L) { ++R; }
    const std::uint64_t R0 = (R / BAND) * BAND;
    const std::uint64_t hh = (T - R0 < static_cast<std::uint64_t>(BAND)) ? (T - R0) : static_cast<std::uint64_t>(BAND);
    std::uint64_t rem = L - R0 * (R0 + 1) / 2;
    if (rem < R0 * hh) {
        J = static_cast<std::uint32_t>(rem / hh);
        I = static_cast<std::uint32_t>(R0 + rem % hh);
        return;
    }
    rem -= R0 * hh;
    std::uint64_t jp = 0;
    while (rem >= hh - jp) {
        rem -= hh - jp;
        ++jp;
    }
    J = static_cast<std::uint32_t>(R0 + jp);
    I = static_cast<std::uint32_t>(R0 + jp + rem);
}

// rectangle Tr x Tc (prediction): bands of BAND tile rows, column-major inside a band
PB_HD std::uint64_t rect_encode(const std::uint64_t Tr, const std::uint64_t Tc, const std::uint64_t I, const std::uint64_t J) {
    const std::uint64_t R0 = (I / BAND) * BAND;
    const std::uint64_t hh = (Tr - R0 < static_cast<std::uint64_t>(BAND)) ? (Tr - R0) : static_cast<std::uint64_t>(BAND);
    return R0 * Tc + J * hh + (I - R0);
}

PB_HD void rect_decode(const std::uint64_t Tr, const std::uint64_t Tc, const std::uint64_t L, std::uint32_t &I, std::uint32_t &J) {
    const std::uint64_t R0 = (L / (static_cast<std::uint64_t>(BAND) * Tc)) * BAND;
    const std::uint64_t hh = (Tr - R0 < static_cast<std::uint64_t>(BAND)) ? (Tr - R0) : static_cast<std::uint64_t>(BAND);
    const std::uint64_t rem = L - R0 * Tc;
    J = static_cast<std::uint32_t>(rem / hh);
    I = static_cast<std::uint32_t>(R0 + rem % hh);
}

// contiguous share of `total` tiles owned by `rank` of `world`
PB_HD void rank_range(const std::uint64_t total, const int rank, const int world, std::uint64_t &lo, std::uint64_t &hi) {
    lo = total * static_cast<std::uint64_t>(rank) / static_cast<std::uint64_t>(world);
    hi = total * (static_cast<std::uint64_t>(rank) + 1) / static_cast<std::uint64_t>(world);
}


// the same with shares proportional to weights[0 .. world) (rate-weighted shares: GPUs under a power cap run at different clocks);
// cut points are rounded down, so the ranges tile [0, total) exactly for any weights; non-positive weights count as equal shares
inline void weighted_range(const std::uint64_t total, const int rank, const int world, const double *weights, std::uint64_t &lo, std::uint64_t &hi) {
    double sum = 0.0;
    bool ok = weights != nullptr;
    for (int g = 0; ok && g < world; ++g) {
        ok = weights[g] > 0.0 && weights[g] < 1e300;
        sum += weights[g];
    }
    bool equal = ok;
    for (int g = 1; equal && g < world; ++g) { equal = weights[g] == weights[0]; }
    if (!ok || equal) {  // equal weights: exactly the integer split
        rank_range(total, rank, world, lo, hi);
        return;
    }
    double before = 0.0;
    for (int g = 0; g < rank; ++g) { before += weights[g]; }
    auto cut = [&](const double cum) {
        const double c = static_cast<double>(total) * (cum / sum);
        const std::uint64_t v = static_cast<std::uint64_t>(c < 0.0 ? 0.0 : c);
        return v > total ? total : v;
    };
    lo = rank == 0 ? 0 : cut(before);
    hi = rank == world - 1 ? total : cut(before + weights[rank]);
    if (hi < lo) { hi = lo; }
}

}  // namespace pb
