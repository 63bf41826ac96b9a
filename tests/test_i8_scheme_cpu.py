"""CPU restatement of the int8-slice contraction of plssvm_b200/csrc/tile_i8.cuh (split_i8_kernel + the digit-diagonal products), in exact
integer arithmetic with numpy — pins the ARITHMETIC of the default tile kernel without a GPU:

  * the balanced base-256 digits reproduce the fixed-point value exactly and stay in int8,
  * keeping the S most significant digit diagonals (S (S + 1) / 2 products) with S = 7 is at least as accurate as a native fp64 dot product,
  * S = 3 (fp32 default) is at the accuracy of fp32 FMA accumulation, S = 4 below it,
  * no int32 accumulator overflows for d <= 16,384 (I8_MAX_FEATURES).

The GPU tests (tests/test_gpu_parity.py) check the kernel itself against the oracle; this file documents why those tolerances hold."""
import numpy as np
import pytest


def split(x, S):
    """Rows -> S digit planes (int64 holding int8 values) and the row exponent e (|x_k| < 2^e), as split_i8_kernel does."""
    mx = np.abs(x).max(axis=1)
    _, e = np.frexp(mx)
    fixed = np.rint(np.ldexp(x, (8 * S - 2) - e[:, None])).astype(np.int64)
    value = fixed.copy()
    planes = []
    for _ in range(S - 1):
        a = (fixed & 0xFF).astype(np.uint8).view(np.int8).astype(np.int64)  # low byte read as signed = balanced digit
        planes.append(a)
        fixed = (fixed - a) >> 8
    planes.append(fixed)
    return np.stack(planes), e, value


def sliced_dot(A, B, S, drop=()):
    pa, ea, _ = split(A, S)
    pb_, eb, _ = split(B, S)
    acc = [np.zeros((A.shape[0], B.shape[0]), dtype=np.int64) for _ in range(S)]
    for p in range(S):
        for q in range(S):
            t = p + q - (S - 1)
            if t >= 0 and (p, q) not in drop:  # the S most significant digit diagonals only
                acc[t] += pa[p] @ pb_[q].T
    s = acc[0].astype(np.float64)
    for t in range(1, S):
        s = s * 2.0 ** -8 + acc[t].astype(np.float64)  # the kernel uses one fp64 FMA per diagonal
    return s * np.ldexp(1.0, ea - 6)[:, None] * np.ldexp(1.0, eb - 6)[None, :], acc


def data(d, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (24, d)).astype(dtype)
    B = rng.uniform(-1, 1, (20, d)).astype(dtype)
    A[3] *= dtype(1e-7)
    A[5, ::2] *= dtype(1e-3)
    B[2] *= dtype(1e5)
    return A.astype(np.float64), B.astype(np.float64)


@pytest.mark.parametrize("S", [3, 4, 7])
def test_digits_are_int8_and_reconstruct_exactly(S):
    A, _ = data(300, 1)
    planes, e, value = split(A, S)
    assert planes.min() >= -128 and planes.max() <= 127
    recon = sum(planes[p] << (8 * p) for p in range(S))
    assert np.array_equal(recon, value)
    # the fixed-point value is the input rounded to 8 S - 2 bits relative to the row maximum
    err = np.abs(np.ldexp(value.astype(np.float64), e[:, None] - (8 * S - 2)) - A)
    assert np.all(err <= np.ldexp(1.0, e[:, None] - (8 * S - 1)))


@pytest.mark.parametrize("d", [64, 1000, 4096])
def test_seven_slices_are_at_least_as_accurate_as_fp64(d):
    A, B = data(d, 2)
    exact = A.astype(np.longdouble) @ B.astype(np.longdouble).T
    scale = np.sqrt((A * A).sum(1))[:, None] * np.sqrt((B * B).sum(1))[None, :]
    got, acc = sliced_dot(A, B, 7)
    err_i8 = float(np.max(np.abs(got - exact) / scale))
    err_f64 = float(np.max(np.abs(A @ B.T - exact) / scale))
    assert err_i8 <= max(err_f64, 2.0 ** -56)
    assert err_i8 < 2.0 ** -53
    assert max(int(np.abs(a).max()) for a in acc) < 2 ** 31


def test_all_28_products_are_needed_for_fp64_accuracy():
    """The product count of the fp64 kernel is minimal: leaving out even the two least significant products (lowest plane x highest plane)
    puts the result above the error of a native fp64 dot product."""
    A, B = data(4096, 2)
    exact = A.astype(np.longdouble) @ B.astype(np.longdouble).T
    scale = np.sqrt((A * A).sum(1))[:, None] * np.sqrt((B * B).sum(1))[None, :]
    err_f64 = float(np.max(np.abs(A @ B.T - exact) / scale))
    err_26 = float(np.max(np.abs(sliced_dot(A, B, 7, drop=((0, 6), (6, 0)))[0] - exact) / scale))
    assert err_26 > 4.0 * err_f64


def test_fp32_slice_counts():
    A, B = data(1024, 3, np.float32)
    exact = A.astype(np.longdouble) @ B.astype(np.longdouble).T
    scale = np.sqrt((A * A).sum(1))[:, None] * np.sqrt((B * B).sum(1))[None, :]
    fma32 = np.zeros((A.shape[0], B.shape[0]), np.float32)
    A32, B32 = A.astype(np.float32), B.astype(np.float32)
    for k in range(A.shape[1]):  # sequential fp32 accumulation, the reference's dot product (operators.hpp:117-126) up to FMA vs mul + add
        fma32 += A32[:, k:k + 1] * B32[None, :, k]
    err_fp32 = float(np.max(np.abs(fma32.astype(np.float64) - exact) / scale))
    err3 = float(np.max(np.abs(sliced_dot(A, B, 3)[0] - exact) / scale))
    err4 = float(np.max(np.abs(sliced_dot(A, B, 4)[0] - exact) / scale))
    assert err4 < err3 < 2.0 ** -21       # 22 / 30 fixed-point bits relative to the row maximum
    assert err3 < 4.0 * err_fp32          # three slices: the level of an fp32 FMA chain
    assert err4 < err_fp32                # four slices: below it


def test_int32_accumulators_cannot_overflow_at_the_feature_limit():
    # worst case: every digit at its extreme and all products of a diagonal with the same sign
    d, S = 16384, 7
    assert S * d * 128 * 128 < 2 ** 31
