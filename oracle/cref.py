"""ctypes wrapper of oracle/pcd_ref.c (C/OpenMP restatement) -- TEST
INFRASTRUCTURE and the CPU arm of bench.py.  Mirrors oracle.petsc_algos.
PCDPreconditioner / fgmres on the same inputs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libpcd_ref.so")
_lib = None
_vp = C.c_void_p


def available():
    return os.path.exists(_PATH)


def build():
    subprocess.check_call(["make", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = C.CDLL(_PATH)
        _lib.ref_hier_create.restype = _vp
        _lib.ref_hier_create.argtypes = [C.c_int, C.c_int, C.c_double]
        _lib.ref_hier_set_level.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_double] + [_vp] * 6
        _lib.ref_hier_set_coarse.argtypes = [_vp, C.c_int, _vp]
        _lib.ref_hier_free.argtypes = [_vp]
        _lib.ref_vcycle.argtypes = [_vp, _vp, _vp]
        _lib.ref_pcd_create.restype = _vp
        _lib.ref_pcd_create.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib.ref_pcd_free.argtypes = [_vp]
        _lib.ref_pcd_set_matrix.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
        _lib.ref_pcd_set_inner.argtypes = [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _vp, _vp,
                                           _vp, _vp, C.c_int]
        _lib.ref_schur_apply.argtypes = [_vp, _vp, _vp]
        _lib.ref_pc_apply.argtypes = [_vp] * 5
        _lib.ref_system_matvec.argtypes = [_vp, _vp, _vp]
        _lib.ref_fgmres.restype = C.c_int
        _lib.ref_fgmres.argtypes = [_vp, _vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, _vp, _vp]
        _lib.ref_spmv.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp]
        _lib.ref_cheb_jacobi.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int,
                                         C.c_double, _vp, _vp, _vp, _vp]
        _lib.ref_aggregate_greedy.restype = C.c_int
        _lib.ref_aggregate_greedy.argtypes = [C.c_int, _vp, _vp, _vp, _vp]
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class _Csr:
    """Keeps the int32/float64 arrays of a scipy CSR alive for the C side."""

    def __init__(self, A):
        A = A.tocsr()
        self.shape = A.shape
        self.rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        self.ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.v = np.ascontiguousarray(A.data, dtype=np.float64)


def num_threads():
    return int(lib().ref_num_threads())


def aggregate_greedy(S):
    S = S.tocsr()
    sp_ = np.ascontiguousarray(S.indptr, dtype=np.int32)
    sc = np.ascontiguousarray(S.indices, dtype=np.int32)
    sv = np.ascontiguousarray(S.data, dtype=np.float64)
    agg = np.empty(S.shape[0], dtype=np.int64)
    nagg = lib().ref_aggregate_greedy(S.shape[0], _p(sp_), _p(sc), _p(sv), _p(agg))
    return agg, int(nagg)


class CHierarchy:
    """oracle.amg.Hierarchy mirrored into the C library."""

    def __init__(self, H):
        L = lib()
        self._keep = []
        self.h = L.ref_hier_create(len(H.levels), int(H.smooth_steps), float(H.eig_ratio))
        for l, lv in enumerate(H.levels):
            A = _Csr(lv.A)
            dinv = np.ascontiguousarray(lv.dinv, dtype=np.float64)
            if lv.P is not None:
                P, R = _Csr(lv.P), _Csr(lv.R)
                nc = P.shape[1]
                args = [_p(P.rp), _p(P.ci), _p(P.v), _p(R.rp), _p(R.ci), _p(R.v)]
                self._keep += [P, R]
            else:
                nc = 0
                args = [None] * 6
            self._keep += [A, dinv]
            L.ref_hier_set_level(self.h, l, A.shape[0], nc, _p(A.rp), _p(A.ci), _p(A.v), _p(dinv), float(lv.rho), *args)
        self.cinv = np.ascontiguousarray(H.coarse_inv, dtype=np.float64)
        L.ref_hier_set_coarse(self.h, self.cinv.shape[0], _p(self.cinv))
        self.n = H.levels[0].A.shape[0]

    def vcycle(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.n)
        lib().ref_vcycle(self.h, _p(b), _p(x))
        return x

    __call__ = vcycle

    def __del__(self):
        try:
            lib().ref_hier_free(self.h)
        except Exception:
            pass


class CPCD:
    """The block-triangular PCD preconditioner + FGMRES in C/OpenMP.
    ``mats``: dict name -> scipy CSR for A00 A01 A10 Ap Mp Kp [P00]."""

    IDS = {"A00": 0, "A01": 1, "A10": 2, "Ap": 3, "Mp": 4, "Kp": 5, "P00": 6}

    def __init__(self, mats, variant, bc_idx, bc_val, Hu, Hp, cheb_bounds, cheb_steps=5, ap_its=2, u_its=1):
        L = lib()
        self.n_u, self.n_p = mats["A00"].shape[0], mats["Mp"].shape[0]
        self.p = L.ref_pcd_create(self.n_u, self.n_p, 1 if variant == "BRM1" else 2)
        self._keep = []
        for name, A in mats.items():
            if A is None:
                continue
            a = _Csr(A)
            self._keep.append(a)
            L.ref_pcd_set_matrix(self.p, self.IDS[name], a.shape[0], a.shape[1], _p(a.rp), _p(a.ci), _p(a.v))
        self.dinv = np.ascontiguousarray(1.0 / mats["Mp"].diagonal())
        self.bc_idx = np.ascontiguousarray(bc_idx, dtype=np.int32)
        self.bc_val = np.ascontiguousarray(bc_val, dtype=np.float64)
        self.Hu = Hu if isinstance(Hu, CHierarchy) else CHierarchy(Hu)
        self.Hp = Hp if isinstance(Hp, CHierarchy) else CHierarchy(Hp)
        L.ref_pcd_set_inner(self.p, _p(self.dinv), float(cheb_bounds[0]), float(cheb_bounds[1]), cheb_steps, ap_its,
                            u_its, self.Hu.h, self.Hp.h, _p(self.bc_idx), _p(self.bc_val), self.bc_idx.size)

    @classmethod
    def from_problem(cls, prob, Hu, Hp):
        mats = {"A00": prob.A00, "A01": prob.A01, "A10": prob.A10, "Ap": prob.Ap, "Mp": prob.Mp, "Kp": prob.Kp,
                "P00": prob.P00}
        return cls(mats, prob.variant, prob.bc_idx, prob.bc_val, Hu, Hp, prob.cheb_bounds)

    def schur_apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n_p)
        lib().ref_schur_apply(self.p, _p(x), _p(y))
        return y

    def apply_split(self, xu, xp):
        xu = np.ascontiguousarray(xu, dtype=np.float64)
        xp = np.ascontiguousarray(xp, dtype=np.float64)
        yu, yp = np.empty(self.n_u), np.empty(self.n_p)
        lib().ref_pc_apply(self.p, _p(xu), _p(xp), _p(yu), _p(yp))
        return yu, yp

    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib().ref_system_matvec(self.p, _p(x), _p(y))
        return y

    def fgmres(self, b, rtol=1e-6, atol=1e-50, restart=150, max_it=10000):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        hist = np.zeros(max_it + 2)
        nap = C.c_int()
        its = lib().ref_fgmres(self.p, _p(b), _p(x), rtol, atol, restart, max_it, _p(hist), C.byref(nap))
        return x, int(its), hist[: its + 1].copy(), int(nap.value)

    def __del__(self):
        try:
            lib().ref_pcd_free(self.p)
        except Exception:
            pass
