"""The torch input generator of bench.py against the numpy oracle assembler
(same discretisation, different node numbering)."""
import numpy as np
import pytest
import scipy.sparse as sp

import bench_inputs as bi
from oracle import problems


def _perm_maps(space, prob_t):
    """oracle node -> lattice node, oracle vertex -> lattice vertex."""
    g = prob_t.g
    nx, ny, nz = g.n
    X = space.node_coords
    L = np.rint(X / (0.5 * g.h)).astype(np.int64)
    node = (L[:, 2] * (2 * ny + 1) + L[:, 1]) * (2 * nx + 1) + L[:, 0]
    Lv = np.rint(space.verts / g.h).astype(np.int64)
    vert = (Lv[:, 2] * (ny + 1) + Lv[:, 1]) * (nx + 1) + Lv[:, 0]
    return node, vert


@pytest.mark.parametrize("kind,variant", [("cavity", "BRM2"), ("channel", "BRM1")])
def test_generator_matches_oracle(kind, variant):
    if kind == "cavity":
        n = (3, 3, 3)
        ref, space = problems.lid_driven_cavity(3, dim=3, variant=variant, stabilise=False)
    else:
        n = (4, 2, 3)
        ref, space = problems.channel(4, 2, 3, variant=variant, stabilise=False)
    t = bi.OseenBoxProblem(*n, kind=kind, variant=variant, device="cpu")
    assert t.n_u == ref.n_u and t.n_p == ref.n_p
    node, vert = _perm_maps(space, t)
    pu = (3 * node[:, None] + np.arange(3)[None]).ravel()     # oracle u dof -> lattice u dof
    Pu = sp.csr_matrix((np.ones(pu.size), (pu, np.arange(pu.size))), shape=(pu.size, pu.size))
    Pp = sp.csr_matrix((np.ones(vert.size), (vert, np.arange(vert.size))), shape=(vert.size, vert.size))
    def close(a, b):
        d = abs(a - b)
        return (d.max() if d.nnz else 0.0) <= 1e-12 * abs(b).max()
    assert close(t.scipy("A00"), Pu @ ref.A00 @ Pu.T)
    assert close(t.scipy("A01"), Pu @ ref.A01 @ Pp.T)
    assert close(t.scipy("A10"), Pp @ ref.A10 @ Pu.T)
    assert close(t.scipy("Mp"), Pp @ ref.Mp @ Pp.T)
    assert close(t.scipy("Kp"), Pp @ ref.Kp @ Pp.T)
    assert close(t.scipy("Ap"), Pp @ ref.Ap @ Pp.T)
    assert np.allclose(t.b_u, Pu @ ref.b_u, atol=1e-13)
    assert np.allclose(t.b_p, Pp @ ref.b_p, atol=1e-13)
    assert sorted(t.bc_idx) == sorted(vert[ref.bc_idx])
    # the structural pattern is the same too (value-independent)
    assert t.scipy("A00").nnz == ref.A00.nnz and t.scipy("A10").nnz == ref.A10.nnz


def test_row_partition_is_a_row_slice():
    full = bi.OseenBoxProblem(3, 2, 4, kind="channel", variant="BRM1", device="cpu")
    parts = [bi.OseenBoxProblem(3, 2, 4, kind="channel", variant="BRM1", device="cpu", rank=r, nranks=3)
             for r in range(3)]
    assert sum(p.n_u for p in parts) == full.n_u and sum(p.n_p for p in parts) == full.n_p
    for name in ("A00", "A01", "A10", "Mp", "Kp", "Ap"):
        stacked = sp.vstack([p.scipy(name) for p in parts]).tocsr()
        d = abs(stacked - full.scipy(name))
        assert (d.max() if d.nnz else 0.0) <= 1e-13, name
        assert stacked.nnz == full.scipy(name).nnz
    assert np.allclose(np.concatenate([p.b_u for p in parts]), full.b_u, atol=1e-14)
    assert np.allclose(np.concatenate([p.b_p for p in parts]), full.b_p, atol=1e-14)


def test_slab_wise_generation_equals_one_shot():
    """bench_inputs.generate builds a rank's rows in z-slabs (bounds the generator's device memory at
    128^3) and chains them: same patterns, values equal to rounding (the summation order of duplicate
    contributions follows the chunking), same right-hand sides and PCD Dirichlet set -- also for the
    scalar operator S00 handed to the oracle's hierarchy set-up, with A00 = S00 (x) I_3."""
    import scipy.sparse as sp
    for world in (1, 2):
        for r in range(world):
            a = bi.OseenBoxProblem(6, 6, 6, rank=r, nranks=world)
            b = bi.generate(6, 6, 6, rank=r, nranks=world, cells_per_slab=70)
            assert type(b).__name__ == "MergedProblem"
            assert (a.u_begin, a.n_u, a.p_begin, a.n_p) == (b.u_begin, b.n_u, b.p_begin, b.n_p)
            for name in bi.MergedProblem.OPS:
                x, y = getattr(a, name), getattr(b, name)
                assert x[0].dtype == y[0].dtype and np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]), name
                assert np.allclose(x[2], y[2], rtol=0, atol=1e-14 * np.abs(x[2]).max()), name
            assert np.array_equal(a.bc_idx, b.bc_idx) and np.allclose(a.b_u, b.b_u, atol=1e-15) and np.allclose(a.b_p, b.b_p, atol=1e-15)
    full = bi.generate(6, 6, 6, cells_per_slab=70)
    S, A = full.scipy("S00"), full.scipy("A00")
    d = abs(sp.kron(S, sp.identity(3), format="csr") - A)
    assert (d.max() if d.nnz else 0.0) == 0.0
