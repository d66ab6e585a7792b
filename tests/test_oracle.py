"""CPU tests of the oracle (the checker itself).  The reference pins no numbers on
this path (oracle/__init__.py), so the PETSc-owned part of the restatement is pinned by the
self-consistency properties of SURVEY.md section 8c and by the committed fixtures; the part the
reference's own Python owns is pinned against its code in tests/test_reference_golden.py."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import amg, cref, fem, petsc_algos as pa, problems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bfs_l2_pcd.npz")


def test_bfs_dof_counts_match_the_reference_documentation():
    # level 4 = 25 987 dofs: demo/unsteady-navier-stokes-pcd/documentation.rst:137
    counts = [problems.bfs_space(l).nu_dofs + problems.bfs_space(l).n1 for l in range(5)]
    assert counts == [142, 479, 1747, 6659, 25987]


def test_lshape_mesh_data_equals_the_reference_xml():
    path = "/root/reference/demo/data/mesh_lshape.xml"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    v, c = fem.load_dolfin_xml(path)
    assert np.array_equal(v, fem.LSHAPE_VERTS) and np.array_equal(c, fem.LSHAPE_CELLS)


@pytest.mark.parametrize("dim", [2, 3])
def test_taylor_hood_reproduces_poiseuille_flow(dim):
    if dim == 2:
        prob, space = problems.channel(8, 4, nu=0.1, length=2.0, stabilise=False)
    else:
        prob, space = problems.channel(4, 3, 3, nu=0.1, length=2.0, stabilise=False)
    x = spla.spsolve(prob.system_matrix().tocsc(), prob.rhs())
    X = space.node_coords
    u = x[:prob.n_u].reshape(-1, dim)
    exact = 4 * X[:, 1] * (1 - X[:, 1]) * (4 * X[:, 2] * (1 - X[:, 2]) if dim == 3 else 1.0)
    if dim == 2:        # exact in 2D (quadratic profile in P2); in 3D the profile is quartic
        assert np.abs(u[:, 0] - exact).max() < 1e-12 and np.abs(u[:, 1:]).max() < 1e-12
        assert np.abs(x[prob.n_u:] - 8 * 0.1 * (2.0 - space.verts[:, 0])).max() < 1e-11
    else:
        assert np.abs(u[:, 0] - exact).max() < 0.1


def test_quadrature_integrates_polynomials_exactly():
    from math import factorial
    for d in (1, 2, 3):
        lam, w = fem.simplex_quadrature(d, 3)
        # int over the unit simplex of prod lam_i^a_i = d! prod a_i! / (d + sum a)!  (weights sum to 1)
        for a in ([2, 1] + [0] * (d - 1), [1, 1] + [1] * (d - 1), [5] + [0] * d):
            a = (a + [0] * (d + 1))[:d + 1]
            num = np.sum(w * np.prod(lam ** np.array(a), axis=1))
            exact = factorial(d) * np.prod([factorial(k) for k in a]) / factorial(d + sum(a))
            assert abs(num - exact) < 1e-14


def test_chebyshev_residual_polynomial_is_the_optimal_one():
    """After k steps the residual polynomial is the shifted Chebyshev polynomial of
    degree k on [emin, emax]: check it on a diagonal matrix."""
    emin, emax, k = 0.5, 2.0, 5
    lam = np.linspace(emin, emax, 201)
    A = sp.diags(lam).tocsr()
    b = np.ones_like(lam)
    x = pa.chebyshev_jacobi(A, np.ones_like(lam), b, emin, emax, k)
    res = 1.0 - lam * x
    t = (emax + emin - 2 * lam) / (emax - emin)
    Tk = np.cos(k * np.arccos(np.clip(t, -1, 1)))
    t0 = (emax + emin) / (emax - emin)
    Tk0 = np.cosh(k * np.arccosh(t0))
    assert np.abs(res - Tk / Tk0).max() < 1e-13
    assert np.abs(res).max() <= 1.0 / Tk0 + 1e-14          # ~ 2 (1/3)^5 = 8.2e-3


def test_exact_schur_complement_gives_two_gmres_iterations():
    """doc/source/math.rst:34-36: with the exact Schur complement the block-triangular
    right preconditioner makes GMRES converge in (at most) two iterations."""
    prob, _ = problems.backward_facing_step(1, variant="BRM1")
    A00inv = pa.direct_solver(prob.A00)
    S = -(prob.A10 @ spla.spsolve(prob.A00.tocsc(), prob.A01.tocsc())).toarray()
    Sinv = np.linalg.inv(S)

    def pc(x):
        yu, yp = pa.fieldsplit_upper_apply(x[:prob.n_u], x[prob.n_u:], lambda r: Sinv @ r, prob.A01, A00inv)
        return np.concatenate([yu, yp])
    x, its, hist, _ = pa.fgmres(prob.system_matrix(), pc, prob.rhs(), rtol=1e-10)
    assert its <= 2


@pytest.mark.parametrize("variant", ["BRM1", "BRM2"])
@pytest.mark.parametrize("ls", ["direct", "iterative"])
def test_pcd_converges_on_the_reference_scenarios(variant, ls):
    """The only assertion of the reference bench (test_pcd_scaling.py:223): converged.
    Iteration counts stay mesh independent-ish (level 2 vs level 3)."""
    its = []
    for level in (2, 3):
        p0, _ = problems.backward_facing_step(level, variant=variant)
        x = pa.direct_solver(p0.system_matrix())(p0.rhs())
        prob, _ = problems.backward_facing_step(level, variant=variant, wind=x[:p0.n_u].reshape(-1, 2), stabilise=True)
        if ls == "direct":
            pc = pa.PCDPreconditioner(prob, "direct")
        else:
            P00 = prob.P00 if prob.P00 is not None else prob.A00
            pc = pa.PCDPreconditioner(prob, "iterative", amg_u=amg.build_hierarchy(P00), amg_p=amg.build_hierarchy(prob.Ap))
        A, b = prob.system_matrix(), prob.rhs()
        xs, n, hist, _ = pa.fgmres(A, pc, b, rtol=1e-6, restart=150, max_it=300)
        assert np.linalg.norm(b - A @ xs) <= 2e-6 * np.linalg.norm(b)
        its.append(n)
    assert its[1] <= 2.0 * its[0] + 5 and max(its) < 120


def test_gmres_and_fgmres_agree_for_a_fixed_preconditioner():
    prob, _ = problems.backward_facing_step(2, variant="BRM1")
    pc = pa.PCDPreconditioner(prob, "direct")
    A, b = prob.system_matrix(), prob.rhs()
    x1, i1, h1, n1 = pa.fgmres(A, pc, b, flexible=True)
    x2, i2, h2, n2 = pa.fgmres(A, pc, b, flexible=False)
    assert i1 == i2 and n2 == n1 + 1 and np.allclose(h1, h2, rtol=1e-8)
    assert np.linalg.norm(x1 - x2) <= 1e-8 * np.linalg.norm(x1)
    # restarts: same solution, residual still meets the tolerance
    x3, i3, h3, _ = pa.fgmres(A, pc, b, restart=7)
    assert np.linalg.norm(b - A @ x3) <= 1.01e-6 * np.linalg.norm(b)


def test_c_port_matches_the_numpy_restatement():
    prob, _ = problems.lid_driven_cavity(6, dim=3)
    P00 = prob.P00 if prob.P00 is not None else prob.A00
    Hu, Hp = amg.build_hierarchy(P00), amg.build_hierarchy(prob.Ap)
    S = amg.strength_graph(prob.Ap, 0.08)
    a1, n1 = amg.aggregate_greedy(S, use_c=False)
    a2, n2 = amg.aggregate_greedy(S, use_c=True)
    assert n1 == n2 and np.array_equal(a1, a2)
    pc = pa.PCDPreconditioner(prob, "iterative", amg_u=Hu, amg_p=Hp)
    cp = cref.CPCD.from_problem(prob, Hu, Hp)
    rng = np.random.default_rng(0)
    xu, xp = rng.standard_normal(prob.n_u), rng.standard_normal(prob.n_p)
    yu, yp = pc.apply_split(xu, xp)
    cu, cq = cp.apply_split(xu, xp)
    assert np.linalg.norm(yu - cu) <= 1e-13 * np.linalg.norm(yu) and np.linalg.norm(yp - cq) <= 1e-12 * np.linalg.norm(yp)
    x1, i1, h1, _ = pa.fgmres(prob.system_matrix(), pc, prob.rhs())
    x2, i2, h2, nap = cp.fgmres(prob.rhs())
    assert i1 == i2 == nap and np.linalg.norm(x1 - x2) <= 1e-10 * np.linalg.norm(x1)


def test_block_local_hierarchy_reduces_to_the_serial_one_and_converges():
    prob, _ = problems.channel(10, 4, 4, variant="BRM1")
    A = prob.P00 if prob.P00 is not None else prob.A00
    H1 = amg.build_hierarchy(A)
    H1b = amg.build_hierarchy(A, blocks=[0, A.shape[0]])
    assert [l.A.nnz for l in H1.levels] == [l.A.nnz for l in H1b.levels]
    H2 = amg.build_hierarchy(A, blocks=[0, 3 * (A.shape[0] // 6), A.shape[0]], replicate_size=100000)
    # no aggregate crosses the block boundary: P is block diagonal
    P = H2.levels[0].P.tocoo()
    b0 = H2.begins[0][1]
    c0 = P.col[P.row < b0].max() + 1          # aggregates of block 0 are numbered first
    assert np.all((P.row < b0) == (P.col < c0))
    assert H2.begins[1] == [0, H2.levels[1].A.shape[0]]     # small levels are coarsened serially again
    rng = np.random.default_rng(1)
    b = rng.standard_normal(A.shape[0])
    x = np.zeros_like(b)
    for _ in range(12):
        x += H2(b - A @ x)
    assert np.linalg.norm(b - A @ x) < 1e-3 * np.linalg.norm(b)


def test_golden_fixture_regression():
    g = np.load(GOLD)
    for variant in ("BRM1", "BRM2"):
        p0, _ = problems.backward_facing_step(2, variant=variant)
        x = pa.direct_solver(p0.system_matrix())(p0.rhs())
        prob, _ = problems.backward_facing_step(2, variant=variant, wind=x[:p0.n_u].reshape(-1, 2), stabilise=True)
        xu, xp = g[f"{variant}_xu"], g[f"{variant}_xp"]
        yu, yp = pa.PCDPreconditioner(prob, "direct").apply_split(xu, xp)
        assert np.linalg.norm(yu - g[f"{variant}_direct_yu"]) <= 1e-9 * np.linalg.norm(yu)
        assert np.linalg.norm(yp - g[f"{variant}_direct_yp"]) <= 1e-9 * np.linalg.norm(yp)
        ch = pa.chebyshev_jacobi(prob.Mp, 1.0 / prob.Mp.diagonal(), xp, 0.5, 2.0, 5)
        assert np.linalg.norm(ch - g[f"{variant}_cheb"]) <= 1e-13 * np.linalg.norm(ch)


def test_streamline_diffusion_parameter_matches_oracle_assembler():
    """fenapack_b200.StabilizationParameterSD (host helper, reference stabilization.py:64-67)
    against the oracle assembler's per-cell parameter on the BFS mesh."""
    from fenapack_b200 import StabilizationParameterSD
    from oracle import fem
    p0, space = problems.backward_facing_step(2)
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    wind = x[:p0.n_u].reshape(-1, 2)
    asm = fem.Assembler(space)
    nu = 0.002                                   # low viscosity: part of the cells has PE > 1
    delta = asm.sd_parameter(wind, nu)
    lam_mid = np.full((1, 3), 1.0 / 3.0)
    phi_mid, _ = fem.p2_basis(lam_mid, space.pairs)
    wmid = np.einsum("l,cld->cd", phi_mid[0], wind[space.cell_nodes])
    sd = StabilizationParameterSD(lambda cells: wmid, nu)
    got = sd.eval_cells(space.cell_diameter())
    assert 0 < np.count_nonzero(got) < got.size
    assert np.array_equal(got, delta)
    # density scales the Peclet number only
    assert np.count_nonzero(StabilizationParameterSD(wmid, nu, 0.5).eval_cells(space.cell_diameter())) <= np.count_nonzero(got)


def test_frozen_prolongator_refresh_is_one_linear_pass_and_keeps_the_iteration_count():
    """Prototype of the device-side numeric refresh: with frozen prolongators
    the coarse operators are linear in the fine values, A_c.data = W @ A.data (oracle.amg.galerkin_plan).
    Exact to rounding, and -- unlike keeping stale coarse levels (pc_amg_lag) -- as good as a rebuild:
    BFS level 3, hierarchy built for half the wind, refreshed for the full wind."""
    p0, _ = problems.backward_facing_step(3, variant="BRM1")
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    w = x[:p0.n_u].reshape(-1, 2)
    p1, _ = problems.backward_facing_step(3, variant="BRM1", wind=0.5 * w)
    p2, _ = problems.backward_facing_step(3, variant="BRM1", wind=w)
    H1 = amg.build_hierarchy_kron(p1.A00, bs=2)
    plans = amg.refresh_plans(H1)
    Hr = amg.refresh_hierarchy(H1, p2.A00, plans, new_rho=False)
    for k in range(1, len(Hr.levels)):
        prev = Hr.levels[k - 1]
        G = (prev.R @ (prev.A @ prev.P)).tocsr()
        assert abs(G - Hr.levels[k].A).max() <= 1e-14 * abs(G).max()
        assert Hr.levels[k].P is H1.levels[k].P

    def its(Hu):
        pc = pa.PCDPreconditioner(p2, "iterative", amg_u=Hu, amg_p=amg.build_hierarchy(p2.Ap))
        return pa.fgmres(p2.system_matrix(), pc, p2.rhs(), rtol=1e-6, restart=150)[1]
    rebuilt, refreshed = its(amg.build_hierarchy_kron(p2.A00, bs=2)), its(Hr)
    assert refreshed <= rebuilt + 5          # measured: 54 (rebuild), 57 (refresh), 67 (stale coarse levels)
