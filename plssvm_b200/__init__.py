"""plssvm_b200 — Python binding (ctypes) of libplssvm_b200.so, the Blackwell-native LS-SVM compute backend.

The product is the C-ABI shared library (include/plssvm_b200.h); this module is the thin host-side mirror used by the
tests, ``bench.py`` and Python users.  It mirrors the reference's backend interface for the hot path:

* :class:`CSVM` ~ ``plssvm::csvm`` (include/plssvm/csvm.hpp:50-222): ``fit`` / ``predict`` / ``score`` plus the two
  virtuals a backend implements, ``solve_system_of_linear_equations`` and ``predict_values`` (csvm.hpp:188-208)
* :class:`Parameter` ~ ``plssvm::detail::parameter`` (parameter.hpp:105-266)
* the four kernel-granular calls ``run_q_kernel`` / ``run_svm_kernel`` / ``run_w_kernel`` / ``run_predict_kernel``
  (gpu_csvm.hpp:208-277)

There is no CPU fallback and nothing here imports the test oracle: without the CUDA library or a B200 the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import build as _build

__all__ = ["Backend", "Dataset", "CGSession", "CSVM", "Parameter", "Model", "BackendError", "lib_path", "load_library",
           "LINEAR", "POLYNOMIAL", "RBF", "kernel_id", "broadcast_bytes", "tile_size", "tri_num_tiles", "tri_encode", "tri_decode", "rank_range"]

LINEAR, POLYNOMIAL, RBF = 0, 1, 2
_KERNELS = {"linear": LINEAR, "polynomial": POLYNOMIAL, "poly": POLYNOMIAL, "rbf": RBF, 0: LINEAR, 1: POLYNOMIAL, 2: RBF}


def kernel_id(kernel) -> int:
    try:
        return _KERNELS[kernel]
    except KeyError:
        raise ValueError(f"unknown kernel function type {kernel!r}") from None


class BackendError(RuntimeError):
    """~ plssvm::b200::backend_exception (reference: cuda::backend_exception, CUDA/exceptions.hpp:26-34)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Timings(ctypes.Structure):
    _fields_ = [("total_ms", ctypes.c_double), ("cg_loop_ms", ctypes.c_double), ("matvec_ms", ctypes.c_double), ("matvec_tile_ms", ctypes.c_double),
                ("matvec_calls", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("matvec_flops", ctypes.c_double), ("h2d_bytes", ctypes.c_double),
                ("d2h_bytes", ctypes.c_double), ("impl_used", ctypes.c_int), ("n_devices", ctypes.c_int),
                ("cg_iterations", ctypes.c_uint64), ("cg_max_iterations", ctypes.c_uint64), ("cg_residuum", ctypes.c_double), ("cg_target_residuum", ctypes.c_double),
                ("cg_epsilon", ctypes.c_double), ("cg_avg_iteration_ms", ctypes.c_double), ("rebalances", ctypes.c_uint64), ("fallback_batches", ctypes.c_uint64),
                ("tile_mma_wait_operands", ctypes.c_double), ("tile_mma_wait_drain", ctypes.c_double), ("tile_producer_wait", ctypes.c_double), ("tile_epilogue_wait", ctypes.c_double)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


_LIB: Optional[ctypes.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """Loads libplssvm_b200.so (building it with nvcc first if it is missing).  Loading needs no GPU."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_build.LIB):
        if not build_if_missing:
            raise FileNotFoundError(f"{_build.LIB} not built; run `python -m plssvm_b200.build`")
        _build.build()
    lib = ctypes.CDLL(_build.LIB)
    vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
    lib.plssvm_b200_last_error.restype = ctypes.c_char_p
    lib.plssvm_b200_create.argtypes = [ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
    lib.plssvm_b200_num_devices.argtypes = [vp, ctypes.POINTER(i32)]
    lib.plssvm_b200_last_trace.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
    lib.plssvm_b200_weighted_range.restype = None
    lib.plssvm_b200_weighted_range.argtypes = [u64, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64), ctypes.POINTER(u64)]
    lib.plssvm_b200_destroy.argtypes = [vp]
    lib.plssvm_b200_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_longlong]
    lib.plssvm_b200_get_timings.argtypes = [vp, ctypes.POINTER(Timings)]
    lib.plssvm_b200_device_count.argtypes = [ctypes.POINTER(i32)]
    lib.plssvm_b200_comm_unique_id.argtypes = [vp]
    lib.plssvm_b200_comm_init.argtypes = [vp, i32, i32, vp]
    lib.plssvm_b200_tile_size.restype = u64
    lib.plssvm_b200_tri_num_tiles.restype = u64
    lib.plssvm_b200_tri_num_tiles.argtypes = [u64]
    lib.plssvm_b200_tri_encode.restype = u64
    lib.plssvm_b200_tri_encode.argtypes = [u64, u64, u64]
    lib.plssvm_b200_tri_decode.restype = None
    lib.plssvm_b200_tri_decode.argtypes = [u64, u64, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    lib.plssvm_b200_rank_range.restype = None
    lib.plssvm_b200_i8_plane_offset.restype = u64
    lib.plssvm_b200_i8_plane_offset.argtypes = [u64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    lib.plssvm_b200_rank_range.argtypes = [u64, i32, i32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    lib.plssvm_b200_dataset_destroy.argtypes = [vp]
    lib.plssvm_b200_cg_step.argtypes = [vp, u64, ctypes.POINTER(u64), ctypes.POINTER(i32)]
    lib.plssvm_b200_cg_abort.argtypes = [vp]
    for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
        getattr(lib, f"plssvm_b200_dataset_create_{suf}").argtypes = [vp, vp, sz, sz, i32, ctypes.POINTER(vp)]
        getattr(lib, f"plssvm_b200_dataset_create_rows_{suf}").argtypes = [vp, vp, sz, sz, ctypes.POINTER(vp)]
        getattr(lib, f"plssvm_b200_solve_{suf}").argtypes = [vp, vp, sz, sz, vp, i32, i32, ct, ct, ct, ct, u64, vp, vp, vp, vp]
        getattr(lib, f"plssvm_b200_solve_rows_{suf}").argtypes = [vp, vp, sz, sz, vp, i32, i32, ct, ct, ct, ct, u64, vp, vp, vp, vp]
        getattr(lib, f"plssvm_b200_predict_rows_{suf}").argtypes = [vp, vp, sz, sz, vp, ct, vp, vp, vp, sz, i32, i32, ct, ct, vp]
        getattr(lib, f"plssvm_b200_solve_dataset_{suf}").argtypes = [vp, vp, vp, i32, i32, ct, ct, ct, ct, u64, vp, vp, vp, vp]
        getattr(lib, f"plssvm_b200_cg_begin_{suf}").argtypes = [vp, vp, vp, i32, i32, ct, ct, ct, ct, ctypes.POINTER(vp)]
        getattr(lib, f"plssvm_b200_cg_finish_{suf}").argtypes = [vp, vp, vp, vp, vp]
        getattr(lib, f"plssvm_b200_cg_trace_{suf}").argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
        getattr(lib, f"plssvm_b200_predict_{suf}").argtypes = [vp, vp, sz, sz, vp, ct, vp, vp, vp, sz, i32, i32, ct, ct, vp]
        getattr(lib, f"plssvm_b200_predict_dataset_{suf}").argtypes = [vp, vp, vp, ct, vp, vp, vp, i32, i32, ct, ct, vp]
        getattr(lib, f"plssvm_b200_q_kernel_{suf}").argtypes = [vp, vp, i32, i32, ct, ct, vp, vp]
        getattr(lib, f"plssvm_b200_matvec_{suf}").argtypes = [vp, vp, vp, vp, ct, ct, ct, i32, i32, ct, ct, vp]
        getattr(lib, f"plssvm_b200_w_kernel_{suf}").argtypes = [vp, vp, vp, vp]
        getattr(lib, f"plssvm_b200_predict_kernel_{suf}").argtypes = [vp, vp, vp, vp, i32, i32, ct, ct, vp]
    _LIB = lib
    return lib


# every symbol include/plssvm_b200.h declares (tests/test_boundary.py checks the header against this list and the .so)
EXPORTED_SYMBOLS = [
    "plssvm_b200_create", "plssvm_b200_destroy", "plssvm_b200_num_devices", "plssvm_b200_last_error", "plssvm_b200_set_option", "plssvm_b200_get_timings", "plssvm_b200_device_count",
    "plssvm_b200_has_experimental", "plssvm_b200_last_trace", "plssvm_b200_comm_unique_id", "plssvm_b200_comm_init", "plssvm_b200_tile_size", "plssvm_b200_tri_num_tiles", "plssvm_b200_tri_encode",
    "plssvm_b200_tri_decode", "plssvm_b200_rank_range", "plssvm_b200_weighted_range", "plssvm_b200_i8_plane_offset", "plssvm_b200_dataset_destroy", "plssvm_b200_cg_step",
    "plssvm_b200_cg_abort",
] + [f"plssvm_b200_{name}_{suf}" for suf in ("f32", "f64")
     for name in ("dataset_create", "dataset_create_rows", "solve", "solve_rows", "solve_dataset", "cg_begin", "cg_finish", "cg_trace", "predict", "predict_rows", "predict_dataset",
                  "q_kernel", "matvec", "w_kernel", "predict_kernel")]


def _check(rc: int) -> None:
    if rc != 0:
        raise BackendError(rc, load_library().plssvm_b200_last_error().decode(errors="replace"))


# ---- host-only helpers: the banded tile schedule ------------------------------------------------------------------------------
def tile_size() -> int:
    return int(load_library().plssvm_b200_tile_size())


def tri_num_tiles(tiles_per_side: int) -> int:
    return int(load_library().plssvm_b200_tri_num_tiles(tiles_per_side))


def tri_encode(tiles_per_side: int, I: int, J: int) -> int:
    return int(load_library().plssvm_b200_tri_encode(tiles_per_side, I, J))


def tri_decode(tiles_per_side: int, L: int):
    I, J = ctypes.c_uint32(), ctypes.c_uint32()
    load_library().plssvm_b200_tri_decode(tiles_per_side, L, ctypes.byref(I), ctypes.byref(J))
    return int(I.value), int(J.value)


def rank_range(total: int, rank: int, world_size: int):
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    load_library().plssvm_b200_rank_range(total, rank, world_size, ctypes.byref(lo), ctypes.byref(hi))
    return int(lo.value), int(hi.value)


def weighted_range(total: int, rank: int, world_size: int, weights):
    """Rate-weighted contiguous share of `total` tiles (option "balance"); the ranges of all ranks tile [0, total) exactly."""
    lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
    w = (ctypes.c_double * world_size)(*[float(x) for x in weights])
    load_library().plssvm_b200_weighted_range(total, rank, world_size, w, ctypes.byref(lo), ctypes.byref(hi))
    return int(lo.value), int(hi.value)


def has_experimental() -> bool:
    """True if the library was built with -DPLSSVM_B200_EXPERIMENTAL (tile-kernel variants 4 / 5 / 8 / 9)."""
    return bool(load_library().plssvm_b200_has_experimental())


def _row_pointers(X: np.ndarray):
    """Array of row pointers into a C-contiguous 2-D array (the std::vector<std::vector<T>> view of the *_rows entry points)."""
    base, stride = X.ctypes.data, X.strides[0]
    return (ctypes.c_void_p * X.shape[0])(*[base + i * stride for i in range(X.shape[0])])


def i8_plane_offset(row: int, feature: int, plane: int, planes: int, box_rows: int, slabs: int, slab_bytes: int = 64) -> int:
    """Byte offset of one digit in the boxed, pre-swizzled plane layout of the int8-slice tile kernel: slabs of 64 features
    (SWIZZLE_64B rows, fp64 kernel) or of 32 features (SWIZZLE_32B rows, fp32 kernel)."""
    return int(load_library().plssvm_b200_i8_plane_offset(row, feature, plane, planes, box_rows, slabs, slab_bytes))


def broadcast_bytes(raw: bytes, size: int, src: int = 0, device: int = 0) -> bytes:
    """Ship `size` bytes from rank `src` to every rank of the default torch.distributed group (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(size, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy().tobytes()


# ---- buffers -----------------------------------------------------------------------------------------------------------------
def _suffix(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"real_type must be float32 or float64, got {dtype}")


def _host(a, dtype=None) -> np.ndarray:
    """C-contiguous numpy view of a host array (numpy array or CPU torch tensor, pinned or not)."""
    if hasattr(a, "numpy") and hasattr(a, "is_cuda"):
        if a.is_cuda:
            raise TypeError("expected a host buffer")
        a = a.numpy()
    a = np.asarray(a)
    if dtype is not None and a.dtype != np.dtype(dtype):
        a = a.astype(dtype)
    return np.ascontiguousarray(a)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class Dataset:
    """A dense row-major matrix resident in HBM (~ the reference's ``data_d`` device pointers, gpu_csvm.hpp:302-346)."""

    def __init__(self, backend: "Backend", X, *, device_ptr: Optional[int] = None, shape=None, dtype=None):
        self.backend = backend
        self._h = ctypes.c_void_p()
        lib = backend.lib
        if device_ptr is not None:
            self.N, self.d = int(shape[0]), int(shape[1])
            self.dtype = np.dtype(dtype)
            fn = getattr(lib, f"plssvm_b200_dataset_create_{_suffix(self.dtype)}")
            _check(fn(backend._h, ctypes.c_void_p(device_ptr), self.N, self.d, 1, ctypes.byref(self._h)))
        else:
            Xh = _host(X)
            if Xh.ndim != 2:
                raise ValueError("data must be a 2-D matrix (one data point per row)")
            self.N, self.d = Xh.shape
            self.dtype = Xh.dtype
            fn = getattr(lib, f"plssvm_b200_dataset_create_{_suffix(self.dtype)}")
            _check(fn(backend._h, _ptr(Xh), self.N, self.d, 0, ctypes.byref(self._h)))

    @classmethod
    def from_torch_cuda(cls, backend: "Backend", t) -> "Dataset":
        """Adopt (copy + pad) a contiguous CUDA torch tensor without a host round trip."""
        import torch
        assert t.is_cuda and t.is_contiguous() and t.dim() == 2
        torch.cuda.current_stream(t.device).synchronize()
        dt = {torch.float32: np.float32, torch.float64: np.float64}[t.dtype]
        return cls(backend, None, device_ptr=t.data_ptr(), shape=tuple(t.shape), dtype=dt)

    def close(self) -> None:
        if self._h:
            self.backend.lib.plssvm_b200_dataset_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Backend:
    """One context for the GPUs this process drives (~ ``cuda::csvm::init``, CUDA/csvm.cu:48-86).  ``Backend(i)``: device i;
    ``Backend(devices=[...])``: a device group — data sets replicated, matvec tiles and predict points sharded over the devices;
    ``Backend(devices="all")``: every visible device like the reference's backends.  Raises :class:`BackendError` without a B200."""

    def __init__(self, device: int = 0, *, devices=None):
        self.lib = load_library()
        self._h = ctypes.c_void_p()
        if devices is None:
            devices = [int(device)]
        if isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be a list of device indices or 'all'")
            _check(self.lib.plssvm_b200_create(None, 0, ctypes.byref(self._h)))
        else:
            ids = (ctypes.c_int * len(devices))(*[int(x) for x in devices])
            _check(self.lib.plssvm_b200_create(ids, len(devices), ctypes.byref(self._h)))
        n = ctypes.c_int()
        _check(self.lib.plssvm_b200_num_devices(self._h, ctypes.byref(n)))
        self.num_devices = int(n.value)
        self.device = int(device) if isinstance(devices, str) else int(devices[0])
        self.rank, self.world_size = 0, 1

    def close(self) -> None:
        if self._h:
            self.lib.plssvm_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int) -> None:
        _check(self.lib.plssvm_b200_set_option(self._h, key.encode(), int(value)))

    def timings(self) -> dict:
        t = Timings()
        _check(self.lib.plssvm_b200_get_timings(self._h, ctypes.byref(t)))
        return t.as_dict()

    def last_trace(self) -> np.ndarray:
        """Residual history r.r of the last finished solve (entry k: after k iterations)."""
        out = np.empty(4097, dtype=np.float64)
        count = ctypes.c_size_t()
        _check(self.lib.plssvm_b200_last_trace(self._h, _ptr(out), out.size, ctypes.byref(count)))
        return out[: count.value].copy()

    # -- multi-GPU ------------------------------------------------------------------------------------------------------------
    def init_comm_from_torch(self) -> None:
        """One process per GPU: broadcast NCCL's unique id over the already-initialised torch.distributed group."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        raw = b""
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            _check(self.lib.plssvm_b200_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
            raw = bytes(buf.raw)
        raw = broadcast_bytes(raw, 128, device=self.device)
        idbuf = ctypes.create_string_buffer(raw, 128)
        _check(self.lib.plssvm_b200_comm_init(self._h, rank, world, ctypes.cast(idbuf, ctypes.c_void_p)))
        self.rank, self.world_size = rank, world

    # -- datasets ------------------------------------------------------------------------------------------------------------
    def dataset(self, X) -> Dataset:
        if hasattr(X, "is_cuda") and X.is_cuda:
            return Dataset.from_torch_cuda(self, X)
        return Dataset(self, X)

    # -- csvm::solve_system_of_linear_equations ----------------------------------------------------------------------------------
    def solve(self, X, y, kernel, *, degree=3, gamma=None, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None):
        """X: host matrix (N x d) or a resident :class:`Dataset`.  Returns dict(alpha[N], rho, iterations, delta, delta0)."""
        k = kernel_id(kernel)
        if isinstance(X, Dataset):
            N, d, dtype = X.N, X.d, X.dtype
        else:
            X = _host(X)
            N, d = X.shape
            dtype = X.dtype
        suf = _suffix(dtype)
        yh = _host(y, dtype)
        if yh.shape != (N,):
            raise ValueError(f"The number of data points in the matrix A ({N}) and the values in the right hand side vector ({yh.size}) must be the same!")
        gamma = 1.0 / d if gamma is None else gamma  # csvm.hpp:304-307
        max_iter = N if max_iter is None else int(max_iter)  # csvm.hpp:269
        alpha = np.empty(N, dtype=dtype)
        rho = np.zeros(1, dtype=dtype)
        iters = np.zeros(1, dtype=np.uint64)
        res = np.zeros(2, dtype=dtype)
        if isinstance(X, Dataset):
            fn = getattr(self.lib, f"plssvm_b200_solve_dataset_{suf}")
            _check(fn(self._h, X._h, _ptr(yh), k, int(degree), gamma, coef0, cost, eps, max_iter, _ptr(alpha), _ptr(rho), _ptr(iters), _ptr(res)))
        else:
            fn = getattr(self.lib, f"plssvm_b200_solve_{suf}")
            _check(fn(self._h, _ptr(X), N, d, _ptr(yh), k, int(degree), gamma, coef0, cost, eps, max_iter, _ptr(alpha), _ptr(rho), _ptr(iters), _ptr(res)))
        return {"alpha": alpha, "rho": rho[0], "iterations": int(iters[0]), "delta": res[0], "delta0": res[1]}

    def solve_rows(self, X, y, kernel, *, degree=3, gamma=None, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None):
        """``solve`` through plssvm_b200_solve_rows_*: the matrix is handed over as row pointers (the reference's vector<vector<T>>)."""
        X = _host(X)
        N, d = X.shape
        suf = _suffix(X.dtype)
        yh = _host(y, X.dtype)
        gamma = 1.0 / d if gamma is None else gamma
        max_iter = N if max_iter is None else int(max_iter)
        alpha, rho, iters, res = np.empty(N, dtype=X.dtype), np.zeros(1, dtype=X.dtype), np.zeros(1, dtype=np.uint64), np.zeros(2, dtype=X.dtype)
        rows = _row_pointers(X)
        _check(getattr(self.lib, f"plssvm_b200_solve_rows_{suf}")(self._h, ctypes.cast(rows, ctypes.c_void_p), N, d, _ptr(yh), kernel_id(kernel), int(degree), gamma, coef0, cost, eps,
                                                                  max_iter, _ptr(alpha), _ptr(rho), _ptr(iters), _ptr(res)))
        return {"alpha": alpha, "rho": rho[0], "iterations": int(iters[0]), "delta": res[0], "delta0": res[1]}

    def predict_values_rows(self, SV, alpha, rho, points, kernel, *, degree=3, gamma=None, coef0=0.0):
        """``predict_values`` through plssvm_b200_predict_rows_* (support vectors and points as row pointers)."""
        SV = _host(SV)
        points = _host(points, SV.dtype)
        n_sv, d = SV.shape
        suf = _suffix(SV.dtype)
        alpha = _host(alpha, SV.dtype)
        gamma = 1.0 / d if gamma is None else gamma
        out = np.empty(points.shape[0], dtype=SV.dtype)
        w_buf, w_valid = np.zeros(d, dtype=SV.dtype), ctypes.c_int(0)
        sv_rows, pt_rows = _row_pointers(SV), _row_pointers(points)
        _check(getattr(self.lib, f"plssvm_b200_predict_rows_{suf}")(self._h, ctypes.cast(sv_rows, ctypes.c_void_p), n_sv, d, _ptr(alpha), rho, _ptr(w_buf),
                                                                    ctypes.cast(ctypes.byref(w_valid), ctypes.c_void_p), ctypes.cast(pt_rows, ctypes.c_void_p), points.shape[0],
                                                                    kernel_id(kernel), int(degree), gamma, coef0, _ptr(out)))
        return out, (w_buf if w_valid.value else None)

    def solve_traced(self, X, y, kernel, *, degree=3, gamma=None, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None, interval=8):
        """``solve`` through the session API, additionally returning the residual history under ``"trace"``."""
        ds = X if isinstance(X, Dataset) else self.dataset(X)
        max_iter = ds.N if max_iter is None else int(max_iter)
        cg = self.cg_begin(ds, y, kernel, degree=degree, gamma=gamma, coef0=coef0, cost=cost, eps=eps)
        done, conv = 0, False
        while done < max_iter and not conv:
            asked = min(interval, max_iter - done)
            now, conv = cg.step(asked)
            if not conv and now != done + asked:
                raise BackendError(3, "CG session lost iterations")
            done = now
        tr = cg.trace()
        res = cg.finish()
        res["trace"] = tr
        return res

    def cg_begin(self, X: "Dataset", y, kernel, *, degree=3, gamma=None, coef0=0.0, cost=1.0, eps=1e-3) -> "CGSession":
        """The same solve as a session whose iterations the caller drives (bench.py times exactly K of them)."""
        return CGSession(self, X, y, kernel, degree=degree, gamma=gamma, coef0=coef0, cost=cost, eps=eps)

    # -- csvm::predict_values ----------------------------------------------------------------------------------------------------
    def predict_values(self, SV, alpha, rho, points, kernel, *, degree=3, gamma=None, coef0=0.0, w=None):
        """Returns (values[m], w) — ``w`` is the linear-kernel normal vector (filled iff kernel is linear), else None."""
        k = kernel_id(kernel)
        sv_ds = isinstance(SV, Dataset)
        pt_ds = isinstance(points, Dataset)
        if sv_ds != pt_ds:
            raise TypeError("pass both the support vectors and the points either as host matrices or as Datasets")
        if sv_ds:
            n_sv, d, dtype = SV.N, SV.d, SV.dtype
            m = points.N
        else:
            SV = _host(SV)
            n_sv, d = SV.shape
            dtype = SV.dtype
            points = _host(points, dtype)
            if points.ndim != 2 or points.shape[1] != d:
                raise ValueError(f"The number of features in the support vectors ({d}) must be the same as in the data points to predict ({points.shape[-1]})!")
            m = points.shape[0]
        suf = _suffix(dtype)
        alpha = _host(alpha, dtype)
        if alpha.shape != (n_sv,):
            raise ValueError(f"The number of support vectors ({n_sv}) and number of weights ({alpha.size}) must be the same!")
        gamma = 1.0 / d if gamma is None else gamma
        out = np.empty(m, dtype=dtype)
        w_buf = np.zeros(d, dtype=dtype)
        w_valid = ctypes.c_int(0)
        if w is not None and len(w) > 0:
            w_buf[:] = _host(w, dtype)
            w_valid.value = 1
        if sv_ds:
            fn = getattr(self.lib, f"plssvm_b200_predict_dataset_{suf}")
            _check(fn(self._h, SV._h, _ptr(alpha), rho, _ptr(w_buf), ctypes.cast(ctypes.byref(w_valid), ctypes.c_void_p), points._h, k, int(degree), gamma, coef0, _ptr(out)))
        else:
            fn = getattr(self.lib, f"plssvm_b200_predict_{suf}")
            _check(fn(self._h, _ptr(SV), n_sv, d, _ptr(alpha), rho, _ptr(w_buf), ctypes.cast(ctypes.byref(w_valid), ctypes.c_void_p), _ptr(points), m, k, int(degree), gamma,
                      coef0, _ptr(out)))
        return out, (w_buf if w_valid.value else None)

    # -- the four run_*_kernel virtuals ---------------------------------------------------------------------------------------------
    def run_q_kernel(self, X: Dataset, kernel, *, degree=3, gamma=None, coef0=0.0):
        """Returns (q[N-1], k(x_last, x_last))."""
        suf = _suffix(X.dtype)
        gamma = 1.0 / X.d if gamma is None else gamma
        q = np.empty(X.N - 1, dtype=X.dtype)
        k_last = np.zeros(1, dtype=X.dtype)
        _check(getattr(self.lib, f"plssvm_b200_q_kernel_{suf}")(self._h, X._h, kernel_id(kernel), int(degree), gamma, coef0, _ptr(q), _ptr(k_last)))
        return q, k_last[0]

    def run_svm_kernel(self, X: Dataset, q, v, ret, QA_cost, cost_inv, add, kernel, *, degree=3, gamma=None, coef0=0.0) -> np.ndarray:
        """ret += add * Q~ v; returns the updated copy of ``ret`` (``cost_inv`` = 1 / C as the reference kernels take it)."""
        suf = _suffix(X.dtype)
        gamma = 1.0 / X.d if gamma is None else gamma
        q = _host(q, X.dtype)
        v = _host(v, X.dtype)
        out = np.array(_host(ret, X.dtype), copy=True)
        _check(getattr(self.lib, f"plssvm_b200_matvec_{suf}")(self._h, X._h, _ptr(q), _ptr(v), QA_cost, cost_inv, add, kernel_id(kernel), int(degree), gamma, coef0, _ptr(out)))
        return out

    def run_w_kernel(self, SV: Dataset, alpha) -> np.ndarray:
        suf = _suffix(SV.dtype)
        alpha = _host(alpha, SV.dtype)
        w = np.empty(SV.d, dtype=SV.dtype)
        _check(getattr(self.lib, f"plssvm_b200_w_kernel_{suf}")(self._h, SV._h, _ptr(alpha), _ptr(w)))
        return w

    def run_predict_kernel(self, SV: Dataset, alpha, points: Dataset, kernel, *, degree=3, gamma=None, coef0=0.0) -> np.ndarray:
        suf = _suffix(SV.dtype)
        gamma = 1.0 / SV.d if gamma is None else gamma
        alpha = _host(alpha, SV.dtype)
        out = np.empty(points.N, dtype=SV.dtype)
        _check(getattr(self.lib, f"plssvm_b200_predict_kernel_{suf}")(self._h, SV._h, _ptr(alpha), points._h, kernel_id(kernel), int(degree), gamma, coef0, _ptr(out)))
        return out


class CGSession:
    """begin / step / finish of one CG solve on a resident dataset (plssvm_b200_cg_*)."""

    def __init__(self, backend: Backend, X: Dataset, y, kernel, *, degree=3, gamma=None, coef0=0.0, cost=1.0, eps=1e-3):
        self.backend, self.X = backend, X
        self._h = ctypes.c_void_p()
        self.suf = _suffix(X.dtype)
        yh = _host(y, X.dtype)
        if yh.shape != (X.N,):
            raise ValueError("label vector has the wrong length")
        gamma = 1.0 / X.d if gamma is None else gamma
        _check(getattr(backend.lib, f"plssvm_b200_cg_begin_{self.suf}")(backend._h, X._h, _ptr(yh), kernel_id(kernel), int(degree), gamma, coef0, cost, eps, ctypes.byref(self._h)))

    def step(self, iterations: int):
        """Enqueue ``iterations`` CG iterations, wait for them, return (completed iterations, converged)."""
        done, conv = ctypes.c_uint64(), ctypes.c_int()
        _check(self.backend.lib.plssvm_b200_cg_step(self._h, int(iterations), ctypes.byref(done), ctypes.byref(conv)))
        return int(done.value), bool(conv.value)

    def trace(self) -> np.ndarray:
        """r.r after 0, 1, 2, ... completed iterations (the reference logs these per iteration, gpu_csvm.hpp:569-571)."""
        out = np.empty(4097, dtype=self.X.dtype)
        count = ctypes.c_size_t()
        _check(getattr(self.backend.lib, f"plssvm_b200_cg_trace_{self.suf}")(self._h, _ptr(out), out.size, ctypes.byref(count)))
        return out[: count.value].copy()

    def finish(self) -> dict:
        alpha = np.empty(self.X.N, dtype=self.X.dtype)
        rho = np.zeros(1, dtype=self.X.dtype)
        iters = np.zeros(1, dtype=np.uint64)
        res = np.zeros(2, dtype=self.X.dtype)
        h, self._h = self._h, ctypes.c_void_p()
        _check(getattr(self.backend.lib, f"plssvm_b200_cg_finish_{self.suf}")(h, _ptr(alpha), _ptr(rho), _ptr(iters), _ptr(res)))
        return {"alpha": alpha, "rho": rho[0], "iterations": int(iters[0]), "delta": res[0], "delta0": res[1]}

    def abort(self) -> None:
        if self._h:
            self.backend.lib.plssvm_b200_cg_abort(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass


# ---- the reference-facing interface ---------------------------------------------------------------------------------------------
@dataclass
class Parameter:
    """~ plssvm::detail::parameter (parameter.hpp:105-266); defaults from parameter.hpp:157-165."""
    kernel_type: object = "linear"
    degree: int = 3
    gamma: Optional[float] = None  # None -> 1 / num_features at fit time (csvm.hpp:304-307)
    coef0: float = 0.0
    cost: float = 1.0


@dataclass
class Model:
    """~ plssvm::model (model.hpp:49-167): params + support vectors + alpha + rho + cached w + the label pair."""
    params: Parameter
    support_vectors: np.ndarray
    alpha: np.ndarray
    rho: float
    labels: tuple  # (label mapped to -1, label mapped to +1): the smaller label maps to -1 (data_set.hpp:438-454)
    w: Optional[np.ndarray] = None
    iterations: int = 0


class CSVM:
    """~ plssvm::csvm with the b200 backend behind it: the same call sequence as csvm.hpp:263-375."""

    def __init__(self, params: Optional[Parameter] = None, *, device: int = 0, backend: Optional[Backend] = None, **named):
        self.params = params if params is not None else Parameter(**named)
        self.backend = backend if backend is not None else Backend(device)

    # the two virtuals (csvm.hpp:188-208)
    def solve_system_of_linear_equations(self, params: Parameter, A, b, eps, max_iter):
        r = self.backend.solve(A, b, params.kernel_type, degree=params.degree, gamma=params.gamma, coef0=params.coef0, cost=params.cost, eps=eps, max_iter=max_iter)
        return r["alpha"], r["rho"], r

    def predict_values(self, params: Parameter, support_vectors, alpha, rho, w, predict_points):
        return self.backend.predict_values(support_vectors, alpha, rho, predict_points, params.kernel_type, degree=params.degree, gamma=params.gamma, coef0=params.coef0, w=w)

    # csvm::fit (csvm.hpp:263-323)
    def fit(self, X, labels: Sequence, *, epsilon: float = 1e-3, max_iter: Optional[int] = None) -> Model:
        X = _host(X)
        labels = np.asarray(labels)
        if not (epsilon > 0):
            raise ValueError(f"epsilon must be greater than 0.0, but is {epsilon}!")
        if max_iter is not None and max_iter <= 0:
            raise ValueError(f"max_iter must be greater than 0, but is {max_iter}!")
        uniq = np.unique(labels)
        if len(uniq) != 2:
            raise ValueError(f"Currently only binary classification is supported, but {len(uniq)} different labels were given!")
        y = np.where(labels == uniq[0], -1.0, 1.0).astype(X.dtype)  # smaller label -> -1 (data_set.hpp:447-453)
        p = Parameter(**vars(self.params))
        if p.gamma is None:
            p.gamma = 1.0 / X.shape[1]
        alpha, rho, info = self.solve_system_of_linear_equations(p, X, y, epsilon, X.shape[0] if max_iter is None else max_iter)
        return Model(p, X, alpha, float(rho), (uniq[0], uniq[1]), None, info["iterations"])

    # csvm::predict (csvm.hpp:325-343): label = mapping(sign(value)), sign(0) = -1 (operators.hpp:178-181)
    def predict(self, model: Model, X) -> np.ndarray:
        X = _host(X, model.support_vectors.dtype)
        if X.shape[1] != model.support_vectors.shape[1]:
            raise ValueError(f"Number of features per data point ({X.shape[1]}) must match the number of features per support vector of the provided model ({model.support_vectors.shape[1]})!")
        values, w = self.predict_values(model.params, model.support_vectors, model.alpha, model.rho, model.w, X)
        if w is not None:
            model.w = w
        return np.where(values > 0, model.labels[1], model.labels[0])

    # csvm::score (csvm.hpp:345-375)
    def score(self, model: Model, X, labels) -> float:
        return float(np.mean(self.predict(model, X) == np.asarray(labels)))
