"""TEST INFRASTRUCTURE ONLY — ctypes loader for the parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package.  The product package ``plssvm_b200`` never does (tests/test_boundary.py checks that).

Two builds expose the same C ABI (oracle/oracle_capi.cpp):

* ``port``      — ``oracle/liboracle_port.so``: the reference's algorithm restated in ``oracle/lssvm_oracle.hpp``
* ``reference`` — ``oracle/_ref/liboracle_ref.so``: the reference's own OpenMP kernel translation units
  (``src/plssvm/backends/OpenMP/{svm_kernel,q_kernel}.cpp``) compiled in place from ``/root/reference`` and driven by
  the restated CG / predict loops (``OpenMP/csvm.cpp:71-183,188-227,255-280`` cannot be compiled offline: igor, fast_float).

Parity status: PINNED — tests/test_oracle_golden.py checks both builds against the reference's known-answer tests
(``tests/backends/generic_csvm_tests.hpp:99-137,149-195,197-247``) and the ``port`` build against the ``reference`` build.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LINEAR, POLYNOMIAL, RBF = 0, 1, 2
KERNEL_IDS = {"linear": LINEAR, "polynomial": POLYNOMIAL, "rbf": RBF}

_PATHS = {
    "port": os.path.join(_HERE, "liboracle_port.so"),
    "reference": os.path.join(_HERE, "_ref", "liboracle_ref.so"),
    "port_fast": os.path.join(_HERE, "liboracle_port_fast.so"),
    "reference_fast": os.path.join(_HERE, "_ref", "liboracle_ref_fast.so"),
}


def build(targets=("port", "ref")) -> None:
    """Compile the oracle (``make -C oracle port ref``).  ``ref`` is a no-op where /root/reference does not exist."""
    subprocess.run(["make", "-s", "-C", _HERE, *targets], check=True)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


class Oracle:
    """One loaded oracle build.  All arrays are C-contiguous numpy arrays of the build's dtype (float32 or float64)."""

    def __init__(self, kind: str = "port"):
        if kind not in _PATHS:
            raise ValueError(f"unknown oracle kind {kind!r}")
        if not available(kind) and kind == "port":
            build(("port",))
        if not available(kind):
            raise FileNotFoundError(f"oracle build {kind!r} not found at {_PATHS[kind]} (run `make -C oracle`)")
        self.kind = kind
        self.lib = ctypes.CDLL(_PATHS[kind])
        self.lib.oracle_kind.restype = ctypes.c_char_p
        self.lib.oracle_max_threads.restype = ctypes.c_int
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p = ctypes.c_void_p
            sz = ctypes.c_size_t
            i = ctypes.c_int
            getattr(self.lib, f"oracle_kernel_function_{suf}").restype = ct
            getattr(self.lib, f"oracle_kernel_function_{suf}").argtypes = [i, p, p, sz, i, ct, ct]
            getattr(self.lib, f"oracle_q_{suf}").restype = None
            getattr(self.lib, f"oracle_q_{suf}").argtypes = [i, p, sz, sz, i, ct, ct, p]
            getattr(self.lib, f"oracle_matvec_{suf}").restype = None
            getattr(self.lib, f"oracle_matvec_{suf}").argtypes = [i, p, sz, sz, p, p, p, ct, ct, ct, i, ct, ct]
            getattr(self.lib, f"oracle_solve_{suf}").restype = i
            getattr(self.lib, f"oracle_solve_{suf}").argtypes = [i, p, sz, sz, p, i, ct, ct, ct, ct, ctypes.c_uint64, p, p, p, p, p]
            getattr(self.lib, f"oracle_w_{suf}").restype = None
            getattr(self.lib, f"oracle_w_{suf}").argtypes = [p, sz, sz, p, p]
            getattr(self.lib, f"oracle_predict_{suf}").restype = None
            getattr(self.lib, f"oracle_predict_{suf}").argtypes = [i, p, sz, sz, p, ct, p, p, p, sz, i, ct, ct, p]

    # -- helpers ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _suf(a: np.ndarray) -> str:
        if a.dtype == np.float64:
            return "f64"
        if a.dtype == np.float32:
            return "f32"
        raise TypeError(f"unsupported dtype {a.dtype}")

    @staticmethod
    def _c(a: np.ndarray, dtype=None) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=dtype if dtype is not None else a.dtype)

    @staticmethod
    def _ptr(a: Optional[np.ndarray]):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def reported_kind(self) -> str:
        return self.lib.oracle_kind().decode()

    def set_threads(self, n: int) -> None:
        self.lib.oracle_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.lib.oracle_max_threads())

    # -- the path ------------------------------------------------------------------------------------------------------
    def kernel_function(self, kernel: int, x, y, degree=3, gamma=1.0, coef0=0.0):
        x = self._c(x)
        y = self._c(y, x.dtype)
        return getattr(self.lib, f"oracle_kernel_function_{self._suf(x)}")(kernel, self._ptr(x), self._ptr(y), x.size, degree, gamma, coef0)

    def q(self, kernel: int, X, degree=3, gamma=1.0, coef0=0.0) -> np.ndarray:
        X = self._c(X)
        N, d = X.shape
        q = np.empty(N - 1, dtype=X.dtype)
        getattr(self.lib, f"oracle_q_{self._suf(X)}")(kernel, self._ptr(X), N, d, degree, gamma, coef0, self._ptr(q))
        return q

    def matvec(self, kernel: int, X, q, v, ret, QA_cost, cost_inv, add, degree=3, gamma=1.0, coef0=0.0) -> np.ndarray:
        """ret += add * Q~ v  (``cost_inv`` = 1 / C, as the reference kernels take it).  Returns the updated copy of ``ret``."""
        X = self._c(X)
        N, d = X.shape
        q = self._c(q, X.dtype)
        v = self._c(v, X.dtype)
        out = np.array(ret, dtype=X.dtype, copy=True)
        getattr(self.lib, f"oracle_matvec_{self._suf(X)}")(kernel, self._ptr(X), N, d, self._ptr(q), self._ptr(v), self._ptr(out), QA_cost, cost_inv, add, degree, gamma, coef0)
        return out

    def solve(self, kernel: int, X, y, degree=3, gamma=1.0, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None, trace=False):
        X = self._c(X)
        N, d = X.shape
        y = self._c(y, X.dtype)
        max_iter = int(N if max_iter is None else max_iter)
        alpha = np.empty(N, dtype=X.dtype)
        rho = np.zeros(1, dtype=X.dtype)
        iters = np.zeros(1, dtype=np.uint64)
        delta = np.zeros(2, dtype=X.dtype)
        tr = np.full(max_iter + 1, np.nan, dtype=X.dtype) if trace else None
        rc = getattr(self.lib, f"oracle_solve_{self._suf(X)}")(kernel, self._ptr(X), N, d, self._ptr(y), degree, gamma, coef0, cost, eps, max_iter,
                                                               self._ptr(alpha), self._ptr(rho), self._ptr(iters), self._ptr(delta), self._ptr(tr))
        if rc != 0:
            raise ValueError("oracle_solve: invalid arguments")
        res = {"alpha": alpha, "rho": rho[0], "iterations": int(iters[0]), "delta": delta[0], "delta0": delta[1]}
        if trace:
            res["trace"] = tr[: int(iters[0]) + 1]
        return res

    def w(self, SV, alpha) -> np.ndarray:
        SV = self._c(SV)
        n_sv, d = SV.shape
        alpha = self._c(alpha, SV.dtype)
        w = np.empty(d, dtype=SV.dtype)
        getattr(self.lib, f"oracle_w_{self._suf(SV)}")(self._ptr(SV), n_sv, d, self._ptr(alpha), self._ptr(w))
        return w

    def predict(self, kernel: int, SV, alpha, rho, P, degree=3, gamma=1.0, coef0=0.0, w=None):
        """Returns (decision values, w) — ``w`` is filled iff the kernel is linear (csvm.cpp:204-207)."""
        SV = self._c(SV)
        n_sv, d = SV.shape
        P = self._c(P, SV.dtype)
        alpha = self._c(alpha, SV.dtype)
        out = np.empty(P.shape[0], dtype=SV.dtype)
        w_buf = np.zeros(d, dtype=SV.dtype)
        w_valid = ctypes.c_int(0)
        if w is not None and len(w) > 0:
            w_buf[:] = w
            w_valid.value = 1
        getattr(self.lib, f"oracle_predict_{self._suf(SV)}")(kernel, self._ptr(SV), n_sv, d, self._ptr(alpha), rho, self._ptr(w_buf), ctypes.byref(w_valid),
                                                             self._ptr(P), P.shape[0], degree, gamma, coef0, self._ptr(out))
        return out, (w_buf if w_valid.value else None)


def sign_labels(values: np.ndarray) -> np.ndarray:
    """csvm.hpp:337-340 + operators.hpp:178-181: label = value > 0 ? +1 : -1 (0 maps to -1)."""
    return np.where(values > 0, 1, -1).astype(np.int32)


class RefCuda:
    """The reference's OWN CUDA kernels (src/plssvm/backends/CUDA/svm_kernel.cu, compiled unchanged for sm_100 by `make -C oracle
    refcuda`) launched with the reference's layout and grid (oracle/ref_cuda_harness.cu): a second oracle on the GPU and the
    GPU-vs-GPU baseline of bench.py.  Test / baseline infrastructure only."""

    PATH = os.path.join(_HERE, "_ref", "libref_cuda.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = ctypes.CDLL(self.PATH)
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(self.lib, f"refcuda_matvec_{suf}").argtypes = [i32, vp, sz, sz, vp, vp, ct, ct, ct, i32, ct, ct, vp, vp, i32]

    def matvec(self, kernel: int, X_cuda, q, v, ret, QA_cost, cost_inv, add, degree=3, gamma=1.0, coef0=0.0, reps=1):
        """X_cuda: contiguous CUDA torch tensor (N x d).  Returns (ret + add * Q~ v, milliseconds per launch)."""
        import torch
        assert X_cuda.is_cuda and X_cuda.is_contiguous()
        torch.cuda.synchronize()
        N, d = X_cuda.shape
        dt = np.float64 if X_cuda.dtype == torch.float64 else np.float32
        suf = "f64" if dt == np.float64 else "f32"
        q = np.ascontiguousarray(q, dtype=dt)
        v = np.ascontiguousarray(v, dtype=dt)
        out = np.array(ret, dtype=dt, copy=True)
        ms = ctypes.c_float(0)
        rc = getattr(self.lib, f"refcuda_matvec_{suf}")(kernel, ctypes.c_void_p(X_cuda.data_ptr()), N, d, q.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p),
                                                        QA_cost, cost_inv, add, degree, gamma, coef0, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ms), reps)
        if rc != 0:
            raise RuntimeError(f"refcuda_matvec failed with code {rc}")
        return out, float(ms.value)


class Exact:
    """The extended-precision, deterministic restatement of the same algorithm (oracle/lssvm_exact.cpp -> liboracle_exact.so):
    compensated dot products (twice the working precision) + long double everywhere else.  Results come back as float64.
    It is the noise-free target: parity is judged by |repo - exact| against |reference - exact| (tests/parity.py)."""

    PATH = os.path.join(_HERE, "liboracle_exact.so")

    def __init__(self):
        if not os.path.exists(self.PATH):
            build(("exact",))
        self.lib = ctypes.CDLL(self.PATH)
        vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(self.lib, f"exact_solve_{suf}").argtypes = [i32, vp, sz, sz, vp, i32, ct, ct, ct, ct, u64, vp, vp, vp, vp]
            getattr(self.lib, f"exact_solve_{suf}").restype = i32
            getattr(self.lib, f"exact_matvec_{suf}").argtypes = [i32, vp, sz, sz, vp, vp, ct, ct, i32, ct, ct, vp, sz, vp]
            getattr(self.lib, f"exact_matvec_{suf}").restype = None
            getattr(self.lib, f"plain_matvec_{suf}").argtypes = [i32, vp, sz, sz, vp, vp, ct, ct, i32, ct, ct, vp, sz, vp]
            getattr(self.lib, f"plain_matvec_{suf}").restype = None
            getattr(self.lib, f"exact_q_{suf}").argtypes = [i32, vp, sz, sz, i32, ct, ct, vp]
            getattr(self.lib, f"exact_q_{suf}").restype = None
            getattr(self.lib, f"exact_predict_{suf}").argtypes = [i32, vp, sz, sz, vp, ct, vp, sz, i32, ct, ct, vp]
            getattr(self.lib, f"exact_predict_{suf}").restype = None

    @staticmethod
    def _prep(X):
        X = np.ascontiguousarray(X)
        if X.dtype not in (np.float32, np.float64):
            raise TypeError("real_type must be float32 or float64")
        return X, ("f32" if X.dtype == np.float32 else "f64")

    def solve(self, kernel: int, X, y, degree=3, gamma=1.0, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None):
        """CG in extended precision with the reference's stopping rule; pin `max_iter` to compare at an equal iteration count."""
        X, suf = self._prep(X)
        N, d = X.shape
        y = np.ascontiguousarray(y, dtype=X.dtype)
        max_iter = N if max_iter is None else int(max_iter)
        alpha, rho, iters, trace = np.empty(N), np.zeros(1), np.zeros(1, dtype=np.uint64), np.zeros(max_iter + 1)
        rc = getattr(self.lib, f"exact_solve_{suf}")(kernel, X.ctypes.data, N, d, y.ctypes.data, int(degree), gamma, coef0, cost, eps, max_iter, alpha.ctypes.data,
                                                     rho.ctypes.data, iters.ctypes.data, trace.ctypes.data)
        if rc != 0:
            raise ValueError("invalid arguments")
        return {"alpha": alpha, "rho": float(rho[0]), "iterations": int(iters[0]), "trace": trace[: int(iters[0]) + 1].copy()}

    def matvec(self, kernel: int, X, q, v, QA_cost, cost_inv, degree=3, gamma=1.0, coef0=0.0, rows=None) -> np.ndarray:
        """Rows `rows` (default: all) of Q~ v for q, v, QA_cost given in the real type of X (the run_svm_kernel argument convention, add = +1, ret = 0)."""
        X, suf = self._prep(X)
        N, d = X.shape
        q = np.ascontiguousarray(q, dtype=X.dtype)
        v = np.ascontiguousarray(v, dtype=X.dtype)
        if rows is None:
            out = np.empty(N - 1)
            getattr(self.lib, f"exact_matvec_{suf}")(kernel, X.ctypes.data, N, d, q.ctypes.data, v.ctypes.data, QA_cost, cost_inv, int(degree), gamma, coef0, None, 0, out.ctypes.data)
        else:
            rows = np.ascontiguousarray(rows, dtype=np.uint64)
            out = np.empty(rows.size)
            getattr(self.lib, f"exact_matvec_{suf}")(kernel, X.ctypes.data, N, d, q.ctypes.data, v.ctypes.data, QA_cost, cost_inv, int(degree), gamma, coef0, rows.ctypes.data,
                                                     rows.size, out.ctypes.data)
        return out

    def reference_arithmetic_matvec(self, kernel: int, X, q, v, QA_cost, cost_inv, rows, degree=3, gamma=1.0, coef0=0.0) -> np.ndarray:
        """The same rows in the reference's arithmetic (sequential FMA chains and sums in the real type of X): its rounding behaviour on sampled rows."""
        X, suf = self._prep(X)
        N, d = X.shape
        q = np.ascontiguousarray(q, dtype=X.dtype)
        v = np.ascontiguousarray(v, dtype=X.dtype)
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.empty(rows.size, dtype=X.dtype)
        getattr(self.lib, f"plain_matvec_{suf}")(kernel, X.ctypes.data, N, d, q.ctypes.data, v.ctypes.data, QA_cost, cost_inv, int(degree), gamma, coef0, rows.ctypes.data, rows.size,
                                                 out.ctypes.data)
        return out

    def q(self, kernel: int, X, degree=3, gamma=1.0, coef0=0.0) -> np.ndarray:
        """k(x_i, x_last) for ALL N rows (the last entry is k(x_last, x_last))."""
        X, suf = self._prep(X)
        out = np.empty(X.shape[0])
        getattr(self.lib, f"exact_q_{suf}")(kernel, X.ctypes.data, X.shape[0], X.shape[1], int(degree), gamma, coef0, out.ctypes.data)
        return out

    def predict(self, kernel: int, SV, alpha, rho, P, degree=3, gamma=1.0, coef0=0.0) -> np.ndarray:
        SV, suf = self._prep(SV)
        P = np.ascontiguousarray(P, dtype=SV.dtype)
        alpha = np.ascontiguousarray(alpha, dtype=SV.dtype)
        out = np.empty(P.shape[0])
        getattr(self.lib, f"exact_predict_{suf}")(kernel, SV.ctypes.data, SV.shape[0], SV.shape[1], alpha.ctypes.data, rho, P.ctypes.data, P.shape[0], int(degree), gamma, coef0,
                                                  out.ctypes.data)
        return out
