"""ctypes loader for integration/_ref/libplssvm_ref_bridge.so — the UNMODIFIED reference library (core + OpenMP backend,
compiled in place from /root/reference with offline shims, integration/Makefile) plus the real `plssvm::csvm` subclass of the
b200 backend.  Test infrastructure only.  Built in the build container; the .so travels to the GPU box."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "integration", "_ref", "libplssvm_ref_bridge.so")
OPENMP, B200 = 0, 1


def available() -> bool:
    return os.path.exists(PATH)


class RefBridge:
    def __init__(self):
        import plssvm_b200
        plssvm_b200.load_library()  # libplssvm_b200.so must be resolvable (rpath points at it)
        self.lib = ctypes.CDLL(PATH)
        vp, sz, i32, f64, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double, ctypes.c_ulonglong
        self.lib.refb_last_error.restype = ctypes.c_char_p
        self.lib.refb_fit_f64.argtypes = [i32, vp, sz, sz, vp, i32, i32, f64, f64, f64, f64, u64, vp, vp, ctypes.c_char_p]
        self.lib.refb_predict_f64.argtypes = [i32, ctypes.c_char_p, vp, sz, sz, vp, vp, vp]
        self.lib.refb_openmp_solve_f64.argtypes = [vp, sz, sz, vp, i32, i32, f64, f64, f64, f64, u64, vp, vp]
        self.lib.refb_openmp_solve_f32.argtypes = [vp, sz, sz, vp, i32, i32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, u64, vp, vp]
        self.lib.refb_openmp_predict_values_f64.argtypes = [vp, sz, sz, vp, f64, vp, sz, i32, i32, f64, f64, vp]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.refb_last_error().decode(errors="replace"))

    @staticmethod
    def _p(a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)

    def fit(self, backend, X, labels, kernel, degree=3, gamma=0.0, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None, model_path=None):
        """plssvm::csvm::fit of the chosen backend through the reference's public API; returns (alpha, rho)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        N, d = X.shape
        alpha = np.empty(N)
        rho = np.zeros(1)
        self._check(self.lib.refb_fit_f64(backend, self._p(X), N, d, self._p(labels), kernel, degree, gamma, coef0, cost, eps, N if max_iter is None else max_iter,
                                          self._p(alpha), self._p(rho), (model_path or "").encode()))
        return alpha, rho[0]

    def predict(self, backend, model_path, P, true_labels=None):
        """plssvm::csvm::predict (+ score) on a LIBSVM model file through the reference's public API."""
        P = np.ascontiguousarray(P, dtype=np.float64)
        m, d = P.shape
        out = np.empty(m, dtype=np.int32)
        score = np.zeros(1)
        tl = None if true_labels is None else np.ascontiguousarray(true_labels, dtype=np.int32)
        self._check(self.lib.refb_predict_f64(backend, model_path.encode(), self._p(P), m, d, self._p(tl), self._p(out), self._p(score)))
        return out, (score[0] if tl is not None else None)

    def openmp_solve(self, X, y, kernel, degree=3, gamma=0.0, coef0=0.0, cost=1.0, eps=1e-3, max_iter=None):
        """The reference's REAL openmp::csvm::solve_system_of_linear_equations (OpenMP/csvm.cpp:71-183)."""
        X = np.ascontiguousarray(X)
        N, d = X.shape
        y = np.ascontiguousarray(y, dtype=X.dtype)
        alpha = np.empty(N, dtype=X.dtype)
        rho = np.zeros(1, dtype=X.dtype)
        fn = self.lib.refb_openmp_solve_f64 if X.dtype == np.float64 else self.lib.refb_openmp_solve_f32
        self._check(fn(self._p(X), N, d, self._p(y), kernel, degree, gamma, coef0, cost, eps, N if max_iter is None else max_iter, self._p(alpha), self._p(rho)))
        return alpha, rho[0]

    def openmp_predict_values(self, SV, alpha, rho, P, kernel, degree=3, gamma=0.0, coef0=0.0):
        SV = np.ascontiguousarray(SV, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        out = np.empty(P.shape[0])
        self._check(self.lib.refb_openmp_predict_values_f64(self._p(SV), SV.shape[0], SV.shape[1], self._p(alpha), rho, self._p(P), P.shape[0], kernel, degree, gamma, coef0, self._p(out)))
        return out
