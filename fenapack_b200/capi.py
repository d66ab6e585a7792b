"""ctypes binding of ``libfenapack_cuda.so`` (include/fenapack_cuda.h).

This is the whole FFI surface a FENaPack maintainer needs: the classes in
``fenapack_b200.preconditioners`` / ``field_split`` call nothing else.  No
PyTorch, no PETSc types -- numpy arrays (host) or raw device addresses.

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_LIB_NAME = "libfenapack_cuda.so"
_lib = None

# operator ids (enum fnp_operator)
MAT_A00, MAT_A01, MAT_A10, MAT_AP, MAT_MP, MAT_KP, MAT_P00, MAT_P01, MAT_A11 = range(9)
MAT_RP = 100      # derived operator of the PCDR variants (introspection / spmv only)
MAT_NAMES = {"A00": MAT_A00, "A01": MAT_A01, "A10": MAT_A10, "Ap": MAT_AP, "Mp": MAT_MP,
             "Kp": MAT_KP, "P00": MAT_P00, "P01": MAT_P01, "A11": MAT_A11}

ERR_ARG, ERR_CUDA, ERR_OPTION, ERR_NCCL, ERR_STATE, ERR_NUMERIC = -1, -2, -3, -4, -5, -6

_c_double_p = C.POINTER(C.c_double)
_c_int32_p = C.POINTER(C.c_int32)
_c_int64_p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol declared in include/fenapack_cuda.h
SIGNATURES = {
    "fnp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "fnp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "fnp_create_dist": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "fnp_destroy": (C.c_int, [C.c_void_p]),
    "fnp_last_error": (C.c_char_p, []),
    "fnp_version": (C.c_char_p, []),
    "fnp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnp_synchronize": (C.c_int, [C.c_void_p]),
    "fnp_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "fnp_set_layout": (C.c_int, [C.c_void_p] + [C.c_int64] * 6),
    "fnp_set_pattern": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fnp_set_values": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fnp_set_bc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "fnp_set_mu_diag": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnp_rp_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_set_index_sets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fnp_setup": (C.c_int, [C.c_void_p]),
    "fnp_spmv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_mp_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_ap_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_u_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_schur_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_pc_apply": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]),
    "fnp_solve": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, _c_int32_p, _c_double_p, _c_int32_p]),
    "fnp_solve_monolithic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _c_int32_p,
                                       _c_double_p, _c_int32_p]),
    "fnp_get_converged_reason": (C.c_int, [C.c_void_p, _c_int32_p]),
    "fnp_get_residual_history": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "fnp_operator_block_size": (C.c_int, [C.c_void_p, C.c_int, _c_int32_p]),
    "fnp_rp_info": (C.c_int, [C.c_void_p, _c_int64_p, _c_int64_p]),
    "fnp_rp_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fnp_amg_num_levels": (C.c_int, [C.c_void_p, C.c_int, _c_int32_p]),
    "fnp_amg_level_info": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _c_int64_p, _c_int64_p,
                                     _c_int64_p, _c_double_p]),
    "fnp_amg_level_get": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fnp_amg_coarse_inverse": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fnp_amg_vcycle": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "fnp_get_timer": (C.c_int, [C.c_void_p, C.c_char_p, _c_double_p, _c_int64_p]),
    "fnp_get_timer_bytes": (C.c_int, [C.c_void_p, C.c_char_p, _c_double_p]),
    "fnp_timer_names": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "fnp_reset_timers": (C.c_int, [C.c_void_p]),
    "fnp_kernel_launches": (C.c_int64, [C.c_void_p]),
    "fnp_event_tic": (C.c_int, [C.c_void_p]),
    "fnp_event_toc": (C.c_int, [C.c_void_p, _c_double_p]),
}


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C fenapack_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FenapackCudaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libfenapack_cuda error {code}: {message}")
        self.code = code


def _check(code):
    if code < 0:
        raise FenapackCudaError(code, load().fnp_last_error().decode())
    return code


def _ptr(a):
    """Address of a numpy array, or pass an integer device address through."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def merge_mpiaij(diag, offdiag, garray, col_start):
    """Local rows of a PETSc MPIAIJ matrix from its two sequential blocks, as the library wants them
    (global column ids, one CSR).  ``diag`` / ``offdiag`` = ``(indptr, indices, data-or-None)`` of the
    diagonal block (columns local to the owned column range starting at ``col_start``) and of the
    off-diagonal block (compressed columns, ``garray[j]`` = global id) -- what ``MatMPIAIJGetSeqAIJ``
    hands out.  Returns ``(indptr, indices, order)``: the merged pattern and the position of every merged
    entry in ``concatenate([diag.data, offdiag.data])``, so that a value refresh is
    ``concatenate([a_d, a_o])[order]`` without touching the pattern again (the MAT_REUSE_MATRIX path,
    field_split_backend.py:331-334)."""
    ia_d, ja_d = np.asarray(diag[0], dtype=np.int64), np.asarray(diag[1], dtype=np.int64)
    ia_o, ja_o = np.asarray(offdiag[0], dtype=np.int64), np.asarray(offdiag[1], dtype=np.int64)
    garray = np.asarray(garray, dtype=np.int64)
    n = ia_d.size - 1
    if ia_o.size - 1 != n:
        raise ValueError("diagonal and off-diagonal blocks must have the same number of rows")
    rows = np.concatenate([np.repeat(np.arange(n), np.diff(ia_d)), np.repeat(np.arange(n), np.diff(ia_o))])
    cols = np.concatenate([ja_d + int(col_start), garray[ja_o] if ja_o.size else ja_o])
    order = np.lexsort((cols, rows))                    # row major, ascending global column
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    if indptr[-1] >= 2 ** 31:
        raise ValueError("more than 2^31 local entries")
    return indptr.astype(np.int32), cols[order].astype(np.int32), order


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load().fnp_nccl_unique_id(buf))
    return buf.raw


class Context:
    """Owner of one ``fnp_context``: the device-resident operators, AMG
    hierarchies and Krylov basis of one PCD-preconditioned solver."""

    def __init__(self, device=0, nccl_id: bytes | None = None, rank=0, nranks=1):
        self._lib = load()
        self._h = C.c_void_p()
        if nccl_id is None:
            _check(self._lib.fnp_create(C.byref(self._h), int(device)))
        else:
            buf = C.create_string_buffer(nccl_id, 128)
            _check(self._lib.fnp_create_dist(C.byref(self._h), int(device), buf, int(rank), int(nranks)))
        self.n_u = self.n_p = 0
        self._shapes = {}

    # -- life cycle --------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fnp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(self._lib.fnp_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        _check(self._lib.fnp_synchronize(self._h))

    # -- options -------------------------------------------------------------
    def set_option(self, name, value):
        _check(self._lib.fnp_set_option(self._h, str(name).encode(), str(value).encode()))

    def set_options(self, opts: dict):
        for k, v in opts.items():
            self.set_option(k, v)

    # -- operators -----------------------------------------------------------
    def set_layout(self, n_u, n_p, u_begin=0, n_u_global=None, p_begin=0, n_p_global=None):
        n_u_global = n_u if n_u_global is None else n_u_global
        n_p_global = n_p if n_p_global is None else n_p_global
        _check(self._lib.fnp_set_layout(self._h, n_u, u_begin, n_u_global, n_p, p_begin, n_p_global))
        self.n_u, self.n_p = int(n_u), int(n_p)

    def set_pattern(self, which, rowptr, colidx):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        _check(self._lib.fnp_set_pattern(self._h, which, _ptr(rowptr), _ptr(colidx)))
        self._shapes[which] = (rowptr.size - 1, int(rowptr[-1]))

    def set_values(self, which, values):
        values = _f64(values)
        if which in self._shapes and values.size != self._shapes[which][1]:
            raise ValueError("value array does not match the pattern (same pattern, new values only)")
        _check(self._lib.fnp_set_values(self._h, which, _ptr(values)))

    def set_values_device(self, which, device_address):
        """Value refresh from an array that already lives in device memory (address as int)."""
        _check(self._lib.fnp_set_values(self._h, which, _ptr(int(device_address))))

    def set_matrix(self, which, A):
        """Upload a scipy CSR matrix (pattern + values)."""
        A = A.tocsr()
        self.set_pattern(which, A.indptr, A.indices)
        self.set_values(which, A.data)

    def set_matrix_mpiaij(self, which, diag, offdiag, garray, col_start):
        """Upload the local rows of an MPIAIJ matrix given as its diagonal / off-diagonal blocks
        (``merge_mpiaij``); the entry order is remembered for ``set_values_mpiaij``."""
        indptr, indices, order = merge_mpiaij(diag, offdiag, garray, col_start)
        if not hasattr(self, "_mpiaij_order"):
            self._mpiaij_order = {}
        self._mpiaij_order[which] = order
        self.set_pattern(which, indptr, indices)
        self.set_values_mpiaij(which, diag[2], offdiag[2])

    def set_values_mpiaij(self, which, diag_values, offdiag_values):
        """Value refresh of an operator uploaded with ``set_matrix_mpiaij`` (same pattern)."""
        vals = np.concatenate([_f64(diag_values).ravel(), _f64(offdiag_values).ravel()])
        self.set_values(which, vals[self._mpiaij_order[which]])

    def set_bc(self, idx, values):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        values = _f64(values)
        _check(self._lib.fnp_set_bc(self._h, _ptr(idx), _ptr(values), idx.size))

    def set_mu_diag(self, diag):
        diag = _f64(diag)
        if diag.size != self.n_u:
            raise ValueError("Mu diagonal must have one entry per local velocity dof")
        _check(self._lib.fnp_set_mu_diag(self._h, _ptr(diag)))

    def rp_solve(self, b):
        return self._solve1(self._lib.fnp_rp_solve, b, self.n_p)

    def set_index_sets(self, is_u, is_p):
        is_u = np.ascontiguousarray(is_u, dtype=np.int64)
        is_p = np.ascontiguousarray(is_p, dtype=np.int64)
        _check(self._lib.fnp_set_index_sets(self._h, _ptr(is_u), _ptr(is_p)))

    def setup(self):
        _check(self._lib.fnp_setup(self._h))

    # -- hot path (host numpy arrays) ------------------------------------------
    def spmv(self, which, x, nrows):
        x = _f64(x)
        y = np.empty(nrows)
        _check(self._lib.fnp_spmv(self._h, which, _ptr(x), _ptr(y), 0))
        return y

    def _solve1(self, fn, b, n):
        b = _f64(b)
        x = np.empty(n)
        _check(fn(self._h, _ptr(b), _ptr(x), 0))
        return x

    def mp_solve(self, b):
        return self._solve1(self._lib.fnp_mp_solve, b, self.n_p)

    def ap_solve(self, b):
        return self._solve1(self._lib.fnp_ap_solve, b, self.n_p)

    def u_solve(self, b):
        return self._solve1(self._lib.fnp_u_solve, b, self.n_u)

    def schur_apply(self, x_p, out=None):
        x_p = _f64(x_p)
        y = np.empty(self.n_p) if out is None else out
        _check(self._lib.fnp_schur_apply(self._h, _ptr(x_p), _ptr(y), 0))
        return y

    def pc_apply(self, x_u, x_p):
        x_u, x_p = _f64(x_u), _f64(x_p)
        y_u, y_p = np.empty(self.n_u), np.empty(self.n_p)
        _check(self._lib.fnp_pc_apply(self._h, _ptr(x_u), _ptr(x_p), _ptr(y_u), _ptr(y_p), 0))
        return y_u, y_p

    def solve(self, b_u, b_p, out=None):
        """FGMRES solve with host vectors.  `out=(x_u, x_p)` writes the solution into caller-owned
        (e.g. pinned) float64 arrays instead of fresh ones."""
        b_u, b_p = _f64(b_u), _f64(b_p)
        if out is None:
            x_u, x_p = np.empty(self.n_u), np.empty(self.n_p)
        else:
            x_u, x_p = out
            for v, n in ((x_u, self.n_u), (x_p, self.n_p)):
                if v.dtype != np.float64 or not v.flags.c_contiguous or v.size != n:
                    raise ValueError("out arrays must be contiguous float64 of the local block sizes")
        its, nap, rn = C.c_int32(), C.c_int32(), C.c_double()
        _check(self._lib.fnp_solve(self._h, _ptr(b_u), _ptr(b_p), _ptr(x_u), _ptr(x_p), 0,
                                   C.byref(its), C.byref(rn), C.byref(nap)))
        return x_u, x_p, its.value, rn.value, nap.value

    def solve_monolithic(self, b):
        b = _f64(b)
        x = np.empty_like(b)
        its, nap, rn = C.c_int32(), C.c_int32(), C.c_double()
        _check(self._lib.fnp_solve_monolithic(self._h, _ptr(b), _ptr(x), 0, C.byref(its), C.byref(rn), C.byref(nap)))
        return x, its.value, rn.value, nap.value

    # -- device-pointer variants (addresses as ints) ---------------------------
    def pc_apply_device(self, x_u, x_p, y_u, y_p):
        _check(self._lib.fnp_pc_apply(self._h, _ptr(x_u), _ptr(x_p), _ptr(y_u), _ptr(y_p), 1))

    def schur_apply_device(self, x_p, y_p):
        _check(self._lib.fnp_schur_apply(self._h, _ptr(x_p), _ptr(y_p), 1))

    def spmv_device(self, which, x, y):
        _check(self._lib.fnp_spmv(self._h, which, _ptr(x), _ptr(y), 1))

    def solve_device(self, b_u, b_p, x_u, x_p):
        its, nap, rn = C.c_int32(), C.c_int32(), C.c_double()
        _check(self._lib.fnp_solve(self._h, _ptr(b_u), _ptr(b_p), _ptr(x_u), _ptr(x_p), 1,
                                   C.byref(its), C.byref(rn), C.byref(nap)))
        return its.value, rn.value, nap.value

    def converged_reason(self):
        r = C.c_int32()
        _check(self._lib.fnp_get_converged_reason(self._h, C.byref(r)))
        return r.value

    def residual_history(self):
        buf = np.empty(20000)
        n = self._lib.fnp_get_residual_history(self._h, _ptr(buf), buf.size)
        return buf[:max(n, 0)].copy()

    # -- introspection -----------------------------------------------------------
    def amg_hierarchy(self, which):
        """Download the AMG hierarchy as scipy matrices: list of dicts with keys
        A, P, R, rho, plus the dense coarse inverse."""
        import scipy.sparse as sp
        nl = C.c_int32()
        _check(self._lib.fnp_amg_num_levels(self._h, which, C.byref(nl)))
        levels = []
        for l in range(nl.value):
            entry = {}
            for kind, key in ((0, "A"), (1, "P"), (2, "R")):
                if kind and l == nl.value - 1:
                    continue
                nr, nc, nnz, rho = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
                _check(self._lib.fnp_amg_level_info(self._h, which, l, kind, C.byref(nr), C.byref(nc),
                                                    C.byref(nnz), C.byref(rho)))
                rp = np.empty(nr.value + 1, dtype=np.int32)
                ci = np.empty(nnz.value, dtype=np.int32)
                va = np.empty(nnz.value)
                _check(self._lib.fnp_amg_level_get(self._h, which, l, kind, _ptr(rp), _ptr(ci), _ptr(va)))
                entry[key] = sp.csr_matrix((va, ci, rp), shape=(nr.value, nc.value))
                entry["rho"] = rho.value
            levels.append(entry)
        nc = levels[-1]["A"].shape[0]
        cinv = np.empty((nc, nc))
        _check(self._lib.fnp_amg_coarse_inverse(self._h, which, _ptr(cinv)))
        return levels, cinv

    def rp_local_rows(self, n_p_global):
        """This rank's rows of the derived PCDR operator Rp (scipy CSR, global column ids)."""
        import scipy.sparse as sp
        nr, nnz = C.c_int64(), C.c_int64()
        _check(self._lib.fnp_rp_info(self._h, C.byref(nr), C.byref(nnz)))
        rp = np.empty(nr.value + 1, dtype=np.int32)
        ci = np.empty(nnz.value, dtype=np.int32)
        va = np.empty(nnz.value)
        _check(self._lib.fnp_rp_get(self._h, _ptr(rp), _ptr(ci), _ptr(va)))
        return sp.csr_matrix((va, ci, rp), shape=(nr.value, n_p_global))

    def block_size(self, which):
        bs = C.c_int32()
        _check(self._lib.fnp_operator_block_size(self._h, which, C.byref(bs)))
        return bs.value

    def amg_vcycle(self, which, b):
        b = _f64(b)
        x = np.empty_like(b)
        _check(self._lib.fnp_amg_vcycle(self._h, which, _ptr(b), _ptr(x), 0))
        return x

    def timer(self, name):
        ms, calls = C.c_double(), C.c_int64()
        _check(self._lib.fnp_get_timer(self._h, name.encode(), C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

    def timer_bytes(self, name):
        b = C.c_double()
        _check(self._lib.fnp_get_timer_bytes(self._h, name.encode(), C.byref(b)))
        return b.value

    def timer_names(self):
        n = _check(self._lib.fnp_timer_names(self._h, None, 0))
        buf = C.create_string_buffer(n + 1)
        _check(self._lib.fnp_timer_names(self._h, buf, n + 1))
        return [t for t in buf.value.decode().split("\n") if t]

    def reset_timers(self):
        _check(self._lib.fnp_reset_timers(self._h))

    def kernel_launches(self):
        return int(self._lib.fnp_kernel_launches(self._h))

    def tic(self):
        _check(self._lib.fnp_event_tic(self._h))

    def toc(self):
        ms = C.c_double()
        _check(self._lib.fnp_event_toc(self._h, C.byref(ms)))
        return ms.value
