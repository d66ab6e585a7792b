"""CPU tests of host-side C++ logic of libfenapack_cuda that has no CUDA dependency: the headers are
compiled alone with g++ into a small harness (tests/host/) and driven through ctypes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import amg, petsc_algos as pa, problems

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "libplan_harness.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", out,
                    os.path.join(HERE, "host", "plan_harness.cpp")], check=True)
    return C.CDLL(out)


def _args(M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    rp, ci, va = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64)
    return M, (rp, ci, va)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_galerkin_plan_cxx_matches_oracle(harness):
    """build_galerkin_plan (csrc/galerkin_plan.hpp) against oracle.amg.galerkin_plan on every level
    transition of a BFS velocity hierarchy: same number of terms, W @ A.data = (P^T A P).data for new
    values of A."""
    p0, _ = problems.backward_facing_step(3, variant="BRM1")
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    p1, _ = problems.backward_facing_step(3, variant="BRM1", wind=x[:p0.n_u].reshape(-1, 2))
    H = amg.build_hierarchy(p0.A00[::2, :][:, ::2].tocsr())
    assert len(H.levels) >= 2
    harness.plan_build.restype = C.c_int64
    A_new = p1.A00[::2, :][:, ::2].tocsr()
    A_new.sort_indices()
    for k, lvl in enumerate(H.levels[:-1]):
        A, a = _args(lvl.A)
        P, p = _args(lvl.P)
        R, r = _args(lvl.P.T)
        pat, W_ref = amg.galerkin_plan(A, P)
        Ac, c = _args(pat)
        n, nc = A.shape[0], P.shape[1]
        terms = harness.plan_build(C.c_int64(n), C.c_int64(nc), *[_ptr(v) for v in a + p + r + c])
        assert terms == W_ref.nnz
        ptr, src, coef = np.empty(Ac.nnz + 1, dtype=np.int64), np.empty(terms, dtype=np.int32), np.empty(terms)
        harness.plan_copy(_ptr(ptr), _ptr(src), _ptr(coef))
        W = sp.csr_matrix((coef, src, ptr), shape=(Ac.nnz, A.nnz))
        vals = A_new.data if k == 0 else np.random.default_rng(k).standard_normal(A.nnz)
        assert np.array_equal(A_new.indices, A.indices) or k > 0
        got = W @ vals
        Anew = sp.csr_matrix((vals, A.indices, A.indptr), shape=A.shape)
        G = (P.T @ (Anew @ P)).tocsr()
        ref = np.array([G[i, j] for i, j in zip(np.repeat(np.arange(nc), np.diff(Ac.indptr)), Ac.indices)])
        assert np.allclose(got, ref, rtol=0, atol=1e-13 * max(1.0, abs(ref).max()))
        assert np.allclose(got, W_ref @ vals, rtol=0, atol=1e-13 * max(1.0, abs(ref).max()))
    # coarse rows that are not sorted by column (several ranks: local numbering [owned | ghost]): same plan,
    # rows of W follow the entry order given
    lvl = H.levels[0]
    A, a = _args(lvl.A)
    P, p = _args(lvl.P)
    R, r = _args(lvl.P.T)
    pat, W_ref = amg.galerkin_plan(A, P)
    rp = pat.indptr.astype(np.int32)
    perm = np.concatenate([np.arange(rp[i + 1] - 1, rp[i] - 1, -1) for i in range(pat.shape[0])])     # every row reversed
    ci_rev = np.ascontiguousarray(pat.indices[perm].astype(np.int32))
    va = np.ones(pat.nnz)
    terms = harness.plan_build(C.c_int64(A.shape[0]), C.c_int64(P.shape[1]), *[_ptr(v) for v in a + p + r + (rp, ci_rev, va)])
    assert terms == W_ref.nnz
    ptr, src, coef = np.empty(pat.nnz + 1, dtype=np.int64), np.empty(terms, dtype=np.int32), np.empty(terms)
    harness.plan_copy(_ptr(ptr), _ptr(src), _ptr(coef))
    W = sp.csr_matrix((coef, src, ptr), shape=(pat.nnz, A.nnz))
    vals = np.random.default_rng(5).standard_normal(A.nnz)
    assert np.allclose(W @ vals, (W_ref @ vals)[perm], rtol=0, atol=1e-13 * abs(W_ref @ vals).max())
    # a coarse pattern that misses product entries is reported, not silently accepted
    A, a = _args(H.levels[0].A)
    P, p = _args(H.levels[0].P)
    R, r = _args(H.levels[0].P.T)
    bad, c = _args(sp.identity(P.shape[1], format="csr"))
    assert harness.plan_build(C.c_int64(A.shape[0]), C.c_int64(P.shape[1]), *[_ptr(v) for v in a + p + r + c]) == -1


def test_mpiaij_block_merge():
    """capi.merge_mpiaij: the diagonal / off-diagonal SeqAIJ blocks of a PETSc MPIAIJ matrix (local column
    ids in the diagonal block, compressed columns + garray in the off-diagonal one) merged into the local
    rows with global, ascending column ids the library ingests; the returned order refreshes values
    without touching the pattern."""
    from fenapack_b200 import capi
    rng = np.random.default_rng(3)
    n_glob, r0, r1 = 60, 20, 45                     # this rank owns rows / columns [20, 45)
    A = sp.random(n_glob, n_glob, density=0.15, random_state=4, format="csr")
    A.data[:] = rng.standard_normal(A.nnz)
    loc = A[r0:r1, :].tocsr()
    loc.sort_indices()
    own = (loc.indices >= r0) & (loc.indices < r1)
    rows = np.repeat(np.arange(r1 - r0), np.diff(loc.indptr))

    def block(mask, colmap):
        M = sp.csr_matrix((loc.data[mask], (rows[mask], colmap(loc.indices[mask]))), shape=(r1 - r0, n_glob))
        M.sort_indices()
        return M
    D = block(own, lambda c: c - r0)
    garray = np.unique(loc.indices[~own])            # PETSc: sorted global ids of the compressed columns
    O = block(~own, lambda c: np.searchsorted(garray, c))
    indptr, indices, order = capi.merge_mpiaij((D.indptr, D.indices, D.data), (O.indptr, O.indices, O.data), garray, r0)
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    assert np.array_equal(indptr, loc.indptr) and np.array_equal(indices, loc.indices)
    assert np.array_equal(np.concatenate([D.data, O.data])[order], loc.data)
    # value refresh: new numbers on the same blocks
    loc2 = loc.copy()
    loc2.data = rng.standard_normal(loc.nnz)
    D2, O2 = loc2.data[own], loc2.data[~own]         # row-major inside each block == the blocks' CSR order
    assert np.array_equal(np.concatenate([D2, O2])[order], loc2.data)
    # an empty off-diagonal block (single rank) and mismatching blocks
    i2, j2, o2 = capi.merge_mpiaij((D.indptr, D.indices, D.data), (np.zeros(r1 - r0 + 1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0)), np.zeros(0, dtype=np.int64), r0)
    assert np.array_equal(i2, D.indptr) and np.array_equal(j2, D.indices + r0)
    with pytest.raises(ValueError):
        capi.merge_mpiaij((D.indptr, D.indices, D.data), (O.indptr[:-1], O.indices, O.data), garray, r0)
