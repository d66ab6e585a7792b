"""Drop-in API tests: the reference's Python surface (PCDKSP, PCDKrylovSolver,
PCDAssembler, PCDNewtonSolver, PCDPC_BRM1/2 as python-PC contexts) driving
libfenapack_cuda.  Mirrors test/unit/test_fieldsplit.py (prefix handling) and the
convergence assertion of test/bench/test_pcd_scaling.py:223."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import fenapack_b200 as fp
from fenapack_b200 import capi
from fenapack_b200.petsc_shim import PC, Mat, Options, Vec
from fenapack_b200.field_split_backend import PCDInterface
from fenapack_b200.field_split import dofmap_dofs_is

from fem_forms import BFSModel


def make_assembler(model, stabilised=False):
    return fp.PCDAssembler(model.a, model.L, [], model.a_pc if stabilised else None, ap=model.ap, kp=model.kp,
                           mp=model.mp, bcs_pcd=model.bc_pcd, function_space=model.W)


def set_iterative_options(prefix, variant):
    o = Options(prefix)
    o.setValue("ksp_gmres_restart", 150)
    o.setValue("fieldsplit_p_pc_python_type", "fenapack.PCDPC_" + variant)
    o.setValue("fieldsplit_u_ksp_type", "richardson")
    o.setValue("fieldsplit_u_ksp_max_it", 1)
    o.setValue("fieldsplit_u_pc_type", "hypre")
    o.setValue("fieldsplit_u_pc_hypre_type", "boomeramg")
    o.setValue("fieldsplit_p_PCD_Ap_ksp_type", "richardson")
    o.setValue("fieldsplit_p_PCD_Ap_ksp_max_it", 2)
    o.setValue("fieldsplit_p_PCD_Ap_pc_type", "hypre")
    o.setValue("fieldsplit_p_PCD_Ap_pc_hypre_type", "boomeramg")
    o.setValue("fieldsplit_p_PCD_Mp_ksp_type", "chebyshev")
    o.setValue("fieldsplit_p_PCD_Mp_ksp_max_it", 5)
    o.setValue("fieldsplit_p_PCD_Mp_ksp_chebyshev_eigenvalues", "0.5, 2.0")
    o.setValue("fieldsplit_p_PCD_Mp_pc_type", "jacobi")


def test_allow_only_one_call_contract_of_reference_unit_test():
    """The checks of the reference's test/unit/test_utils.py::test_allow_only_one_call, against the
    drop-in helper: docstrings survive, arguments pass through, every decorated method has its own
    once-only flag, undecorated methods are untouched."""
    from fenapack.utils import allow_only_one_call        # the alias package, as the reference imports it

    class C(object):
        @allow_only_one_call
        def foo(self, *args, **kwargs):
            """Foo"""
            return args, kwargs

        @allow_only_one_call
        def bar(self, *args, **kwargs):
            """Bar"""
            return args, kwargs

        def baz(self, *args, **kwargs):
            """Baz"""
            return args, kwargs
    o = C()
    assert (o.foo.__doc__, o.bar.__doc__, o.baz.__doc__) == ("Foo", "Bar", "Baz")
    expect = ((1, 2, 3), {"four": 5})
    assert o.foo(1, 2, 3, four=5) == expect and o.bar(1, 2, 3, four=5) == expect and o.baz(1, 2, 3, four=5) == expect
    for method in (o.foo, o.bar):
        with pytest.raises(RuntimeError):
            method(1, 2, 3, four=5)
    assert o.baz(1, 2, 3, four=5) == expect


def test_reference_module_paths_resolve_to_the_dropin_classes():
    """``from fenapack.preconditioners import PCDPC_BRM1`` etc. -- the reference's module layout."""
    import importlib
    for mod, names in (("preconditioners", ("BasePCDPC", "PCDPC_BRM1", "PCDPC_BRM2", "BasePCDRPC", "PCDRPC_BRM1", "PCDRPC_BRM2")),
                       ("field_split", ("PCDKSP", "PCDKrylovSolver")), ("field_split_backend", ("PCDInterface",)),
                       ("assembling", ("PCDAssembler", "PCDForm")),
                       ("nonlinear_solvers", ("PCDNewtonSolver", "PCDNonlinearProblem")),
                       ("stabilization", ("StabilizationParameterSD",)), ("utils", ("allow_only_one_call",))):
        alias = importlib.import_module("fenapack." + mod)
        impl = importlib.import_module("fenapack_b200." + mod)
        for name in names:
            assert getattr(alias, name) is getattr(impl, name), (mod, name)


def test_options_database_forwarding():
    """What PCDKSP.setFromOptions picks up from the options database: the reference's names under the
    user prefix, the library's own knobs (<prefix>fnp_*), and nothing else -- options of solvers the
    library does not have (``mat_mumps_icntl_4``, ``ksp_monitor``; demo_navier-stokes-pcd.py:149-160) are
    ignored as PETSc ignores unused options."""
    Options.clear()
    o = Options("baz_")
    for k, v in (("ksp_gmres_restart", 150), ("fieldsplit_u_ksp_type", "richardson"), ("fieldsplit_u_pc_type", "hypre"),
                 ("fieldsplit_u_mat_mumps_icntl_4", 2), ("ksp_monitor", ""), ("fnp_sell_sigma", 2048),
                 ("fieldsplit_u_pc_amg_refresh", "galerkin")):
        o.setValue(k, v)
    Options("other_").setValue("fnp_cuda_graph", 0)
    ksp = fp.PCDKSP()
    ksp.setOptionsPrefix("baz_")
    ksp.setFromOptions()
    assert ksp._outer_opts == {"ksp_type": "gmres", "ksp_gmres_restart": "150", "fnp_sell_sigma": "2048"}
    assert ksp._u_opts == {"fieldsplit_u_ksp_type": "richardson", "fieldsplit_u_pc_type": "hypre",
                           "fieldsplit_u_pc_amg_refresh": "galerkin"}
    Options.clear()


def test_allow_only_one_call_and_public_names():
    from fenapack_b200.utils import allow_only_one_call

    class A:
        @allow_only_one_call
        def f(self):
            return 1
    a = A()
    assert a.f() == 1
    with pytest.raises(RuntimeError):
        a.f()
    assert A().f() == 1
    for name in ("PCDKSP", "PCDKrylovSolver", "PCDAssembler", "PCDForm", "PCDNewtonSolver", "PCDNonlinearProblem",
                 "PCDPC_BRM1", "PCDPC_BRM2", "PCDRPC_BRM1", "PCDRPC_BRM2", "StabilizationParameterSD"):
        assert hasattr(fp, name)


def test_pcd_form_flags_and_assembler_defaults():
    m = BFSModel(level=0)
    asm = make_assembler(m)
    assert asm.get_pcd_form("ap").is_constant() and asm.get_pcd_form("mp").is_constant()
    assert not asm.get_pcd_form("kp").is_constant()
    assert asm.get_pcd_form("gp").phantom and asm.get_pcd_form("gp").is_phantom()
    assert asm.get_dolfin_form("fp") is None             # reference assembling.py:117-119
    with pytest.raises(AttributeError):
        asm.fp(Mat())                                    # ... assembling a missing form is the error
    with pytest.raises(AttributeError):
        asm.get_pcd_form("nope")
    assert fp.PCDAssembler(m.a, m.L, [], function_space=m.W).pcd_bcs() == []   # default, assembling.py:184-189
    with pytest.raises(AttributeError):
        fp.PCDAssembler(m.a, m.L, [], bcs_pcd=None, function_space=m.W).pcd_bcs()
    Ap = Mat()
    asm.ap(Ap)
    d = Ap.csr.diagonal()
    assert np.all(d[m.bc_pcd.dofs()] == 1.0)
    sub = Ap.csr[m.bc_pcd.dofs(), :]
    assert sub.sum() == len(m.bc_pcd.dofs())       # identity rows (symmetric application)


def test_interface_bc_indices_follow_subfield_position():
    m = BFSModel(level=1, variant="BRM2")
    asm = make_assembler(m)
    A = Mat(m.a())
    itf = PCDInterface(asm, A, dofmap_dofs_is(m.W.sub(0).dofmap()), dofmap_dofs_is(m.W.sub(1).dofmap()))
    idx, vals = itf.pcd_bc_indices()
    assert np.array_equal(m.is_p[idx], m.bc_pcd.dofs()) and np.all(vals == 0.0)
    v = Vec(np.ones(m.is_p.size))
    itf.apply_pcd_bcs(v)
    assert np.all(v.array[idx] == 0.0) and v.array.sum() == m.is_p.size - idx.size
    Kp = itf.setup_mat_Kp()
    assert itf.setup_mat_Kp(mat=Kp) is Kp           # non-constant: refilled in place
    Mp = itf.setup_mat_Mp()
    assert itf.setup_mat_Mp(mat=Mp) is None         # constant: not re-assembled


@pytest.mark.gpu
def test_set_options_prefix_before_and_after_init():
    """Reference test/unit/test_fieldsplit.py:84-96."""
    Options.clear()
    m = BFSModel(level=1)
    set_iterative_options("foo_", "BRM1")
    Options("foo_").setValue("fieldsplit_p_PCD_Mp_ksp_max_it", 3)
    solver = fp.PCDKrylovSolver()
    solver.set_options_prefix("foo_")
    solver.set_from_options()
    A = Mat(m.a())
    solver.set_operators(A, A)
    solver.init_pcd(make_assembler(m))
    assert solver.get_options_prefix() == "foo_"
    assert solver.ksp()._pcd_pc._device_opts["fieldsplit_p_PCD_Mp_ksp_max_it"] == "3"
    with pytest.raises(RuntimeError):
        solver.set_options_prefix("bar_")
    with pytest.raises(RuntimeError):
        solver.init_pcd(make_assembler(m))          # init_pcd only once (field_split.py:60)
    with pytest.raises(RuntimeError):
        solver.ksp()._pcd_pc.init_pcd(None)         # PCDPC re-initialisation (preconditioners.py:66-67)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["BRM1", "BRM2"])
def test_newton_solver_picard_bfs(variant):
    """Nonlinear (Picard) solve through the reference-shaped API; the only thing the
    reference asserts is convergence (test_pcd_scaling.py:223)."""
    Options.clear()
    m = BFSModel(level=3, variant=variant)
    set_iterative_options("", variant)
    linear_solver = fp.PCDKrylovSolver()
    linear_solver.parameters["relative_tolerance"] = 1e-6
    linear_solver.parameters["maximum_iterations"] = 600
    linear_solver.set_from_options()
    # stabilised a_pc for the AMG 00-block, as the reference's "iterative" set-up does
    problem = fp.PCDNonlinearProblem(make_assembler(m, stabilised=True))
    solver = fp.PCDNewtonSolver(linear_solver)
    solver.parameters["relative_tolerance"] = 1e-5
    its, converged = solver.solve(problem, m.w)
    assert converged and 2 <= its <= 30
    assert solver.krylov_iterations() / its < 150
    # the converged state solves the nonlinear problem: compare with a direct Picard loop
    ref = BFSModel(level=3, variant=variant)
    for _ in range(its):
        J, b = ref._system()
        dx = spla.spsolve(J.tocsc(), b)
        ws = np.concatenate([ref.w.array[ref.is_u], ref.w.array[ref.is_p]]) - dx
        ref.w.array[ref.is_u] = ws[:ref.n_u]
        ref.w.array[ref.is_p] = ws[ref.n_u:]
    err = np.linalg.norm(m.w.array - ref.w.array) / np.linalg.norm(ref.w.array)
    assert err < 1e-4
    # the value refresh happened: the device Kp differs from the first (zero-wind) one
    assert linear_solver.ksp()._pcd_pc.mat_Kp.state >= its


@pytest.mark.gpu
def test_python_pc_protocol_schur_only_mode():
    """PCDPC_BRM1 used the way PETSc's fieldsplit would use it: created through
    setPythonContext, configured from the options database, applied to split
    pressure vectors -- checked against the oracle."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(__file__))
    from fenapack_b200 import capi
    from oracle import petsc_algos as pa
    from util import oracle_hierarchy_from_device, relerr
    Options.clear()
    m = BFSModel(level=2, variant="BRM1")
    # non-trivial wind: a Stokes solve
    J, b = m._system()
    dx = spla.spsolve(J.tocsc(), b)
    ws = -dx
    m.w.array[m.is_u] = ws[:m.n_u]
    m.w.array[m.is_p] = ws[m.n_u:]
    set_iterative_options("", "BRM1")
    asm = make_assembler(m)
    A = Mat(m.a())
    is_u, is_p = dofmap_dofs_is(m.W.sub(0).dofmap()), dofmap_dofs_is(m.W.sub(1).dofmap())
    pc = PC(prefix="fieldsplit_p_")
    pc.setFromOptions()                               # instantiates fenapack.PCDPC_BRM1 by dotted name
    ctx = pc.getPythonContext()
    assert type(ctx).__name__ == "PCDPC_BRM1"
    ctx.init_pcd(PCDInterface(asm, A, is_u, is_p, deep_submats=True))
    pc.setUp()
    x = Vec(np.random.default_rng(0).standard_normal(is_p.getSize()))
    y = x.duplicate()
    x0 = x.array.copy()
    pc.apply(x, y)
    assert np.array_equal(x.array, x0)
    Hp = oracle_hierarchy_from_device(ctx._ctx, capi.MAT_AP)
    Mp, Ap, Kp = ctx.mat_Mp.csr, ctx.mat_Ap.csr, ctx.mat_Kp.csr
    dinv = 1.0 / Mp.diagonal()
    idx, vals = ctx.interface.pcd_bc_indices()
    ref = pa.brm1_apply(x0, lambda r: pa.richardson(Ap, Hp, r, 2), Kp,
                        lambda r: pa.chebyshev_jacobi(Mp, dinv, r, 0.5, 2.0, 5), idx, vals)
    assert relerr(y.array, ref) <= 1e-8
    with pytest.raises(ValueError):
        pc.apply(x, x)


@pytest.mark.gpu
def test_pcdr_through_the_python_pc_protocol():
    """PCDRPC_BRM1 as a python-type PC context (Schur-only mode): Mu and Bt come from the
    PCDInterface exactly as in reference preconditioners.py:191-206."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from fenapack_b200 import capi
    from oracle import petsc_algos as pa
    from util import oracle_hierarchy_from_device, relerr
    Options.clear()
    m = BFSModel(level=2, variant="BRM1", idt=5.0)
    J, b = m._system()
    ws = -spla.spsolve(J.tocsc(), b)
    m.w.array[m.is_u] = ws[:m.n_u]
    m.w.array[m.is_p] = ws[m.n_u:]
    set_iterative_options("", "BRM1")
    o = Options("")
    o.setValue("fieldsplit_p_pc_python_type", "fenapack.PCDRPC_BRM1")
    for k, v in (("ksp_type", "richardson"), ("ksp_max_it", 1), ("pc_type", "hypre"), ("pc_hypre_type", "boomeramg")):
        o.setValue("fieldsplit_p_PCD_Rp_" + k, v)
    asm = fp.PCDAssembler(m.a, m.L, [], None, ap=m.ap, kp=m.kp, mp=m.mp, mu=m.mu, bcs_pcd=m.bc_pcd,
                          function_space=m.W)
    A = Mat(m.a())
    is_u, is_p = dofmap_dofs_is(m.W.sub(0).dofmap()), dofmap_dofs_is(m.W.sub(1).dofmap())
    pc = PC(prefix="fieldsplit_p_")
    pc.setFromOptions()
    ctx = pc.getPythonContext()
    assert type(ctx).__name__ == "PCDRPC_BRM1"
    ctx.init_pcd(PCDInterface(asm, A, is_u, is_p, deep_submats=True))
    pc.setUp()
    x = Vec(np.random.default_rng(0).standard_normal(is_p.getSize()))
    y = x.duplicate()
    pc.apply(x, y)
    Mp, Ap, Kp, Bt = ctx.mat_Mp.csr, ctx.mat_Ap.csr, ctx.mat_Kp.csr, ctx.mat_Bt.csr
    Rp = pa.build_rp(Bt, ctx.mat_Mu.csr.diagonal())
    Hp = oracle_hierarchy_from_device(ctx._ctx, capi.MAT_AP)
    Hr = oracle_hierarchy_from_device(ctx._ctx, capi.MAT_RP)
    dinv = 1.0 / Mp.diagonal()
    idx, vals = ctx.interface.pcd_bc_indices()
    ref = pa.pcdr_brm1_apply(x.array, lambda r: pa.richardson(Ap, Hp, r, 2), Kp,
                             lambda r: pa.chebyshev_jacobi(Mp, dinv, r, 0.5, 2.0, 5),
                             lambda r: pa.richardson(Rp, Hr, r, 1), idx, vals)
    assert relerr(y.array, ref) <= 1e-8


@pytest.mark.gpu
def test_unsteady_time_loop_with_per_step_refresh():
    """cfg3 of BASELINE.json at reduced size: backward Euler on the BFS, ONE solver object
    for the whole run (wiring once, Fp/velocity values refreshed at every Newton step of
    every time step -- reference demo_unsteady-navier-stokes-pcd.py:188-208), reaction term
    (1/dt) in Kp (:138)."""
    Options.clear()
    dt = 0.2
    m = BFSModel(level=2, variant="BRM1", idt=1.0 / dt)
    set_iterative_options("", "BRM1")
    kp_steady = m.kp

    def kp_unsteady():        # kp += (1/nu)(1/dt) p q
        from fem_forms import struct_add
        return struct_add(kp_steady(), m._embed_p(m.asm.p1_mass((1.0 / dt) / m.nu)))
    asm = fp.PCDAssembler(m.a, m.L, [], m.a_pc, ap=m.ap, kp=kp_unsteady, mp=m.mp, bcs_pcd=m.bc_pcd,
                          function_space=m.W)
    linear_solver = fp.PCDKrylovSolver()
    linear_solver.parameters["relative_tolerance"] = 1e-6
    linear_solver.parameters["maximum_iterations"] = 400
    linear_solver.set_from_options()
    problem = fp.PCDNonlinearProblem(asm)
    solver = fp.PCDNewtonSolver(linear_solver)
    solver.parameters["relative_tolerance"] = 1e-5
    solver.parameters["absolute_tolerance"] = 1e-9
    ref = BFSModel(level=2, variant="BRM1", idt=1.0 / dt)
    newton_its, t = 0, 0.0
    for step in range(4):
        t += dt
        m.set_time(t)
        ref.set_time(t)
        its, converged = solver.solve(problem, m.w)
        assert converged
        newton_its += its
        m.w0.array[:] = m.w.array
        # the same time step with direct linear solves
        for _ in range(its):
            J, b = ref._system()
            dx = spla.spsolve(J.tocsc(), b)
            ws = np.concatenate([ref.w.array[ref.is_u], ref.w.array[ref.is_p]]) - dx
            ref.w.array[ref.is_u] = ws[:ref.n_u]
            ref.w.array[ref.is_p] = ws[ref.n_u:]
        ref.w0.array[:] = ref.w.array
        assert np.linalg.norm(m.w.array - ref.w.array) <= 1e-3 * np.linalg.norm(ref.w.array)
    ksp = linear_solver.ksp()
    assert ksp._pcd_pc.mat_Kp.state >= newton_its          # one value refresh per Newton step
    assert solver.krylov_iterations() / newton_its < 120
    with pytest.raises(RuntimeError):
        linear_solver.init_pcd(asm)                           # wiring happened exactly once


@pytest.mark.gpu
def test_krylov_operator_is_the_users_A_not_P():
    """The Krylov MatMult must use the system matrix A in full -- including a non-zero 11 block
    (pressure-stabilised discretisations) -- while the triangular apply cuts its blocks from P
    (PCFIELDSPLIT, useAmat = false).  A re-assembled coupling block must reach the device."""
    import scipy.sparse as sp
    Options.clear()
    m = BFSModel(level=2, variant="BRM1")
    x0 = spla.spsolve(m._system()[0].tocsc(), m._system()[1])
    ws = -x0
    m.w.array[m.is_u], m.w.array[m.is_p] = ws[:m.n_u], ws[m.n_u:]          # a non-trivial wind
    set_iterative_options("", "BRM1")
    eps = [1e-3]

    def a_stab():            # A with a (negative definite) pressure-stabilisation block; stored pattern kept
        from fem_forms import struct_add
        return struct_add(m.a(), -eps[0] * m.mp())

    def a_pc():              # P: stabilised velocity block AND a scaled 01 block
        P = sp.lil_matrix(m.a_pc())
        return sp.csr_matrix(P)
    asm = fp.PCDAssembler(a_stab, m.L, [], a_pc, ap=m.ap, kp=m.kp, mp=m.mp, bcs_pcd=m.bc_pcd, function_space=m.W)
    A, P = Mat(), Mat()
    asm.system_matrix(A)
    asm.pc_matrix(P)
    ksp = fp.PCDKSP()
    ksp.setOperators(A, P)
    ksp.setTolerances(rtol=1e-9, max_it=500)
    ksp.init_pcd(asm)
    assert capi.MAT_A11 in ksp._uploaded and capi.MAT_P00 in ksp._uploaded and capi.MAT_P01 not in ksp._uploaded
    b = Vec(np.asarray(m.L(), dtype=float))
    x = Vec(np.zeros(m.N))
    ksp.solve(b, x)
    assert ksp.getConvergedReason() in (2, 3)
    r = b.array - A.csr @ x.array
    assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b.array)            # A's system, 11 block included
    # the 11 block changes (same pattern): the refresh must not keep the old one
    eps[0] = 5e-3
    asm.system_matrix(A)
    ksp.solve(b, x)
    r = b.array - A.csr @ x.array
    assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b.array)
    # a convergence by the absolute tolerance is reported as such, not as a failure
    ksp.setTolerances(rtol=1e-30, atol=1e-6 * np.linalg.norm(b.array))
    ksp.solve(b, x)
    assert ksp.getConvergedReason() == 3


def test_ksp_python_context_protocol():
    """PCDKSPPython (the KSPPYTHON context for a SNES that owns its KSP, reference
    demo/defcon/navier-stokes.py:252-280): PETSc's call sequence create -> setFromOptions -> setUp -> solve,
    repeated setUp/solve after every Jacobian update.  Wiring only (a recording stand-in for PCDKSP): the
    prefix is taken before init_pcd and left alone afterwards, init_pcd happens exactly once, operators are
    re-set at every setUp, tolerances / iteration count / reason travel between the two objects."""
    from fenapack_b200.field_split import PCDKSPPython
    calls = []

    class FakeInner:
        def __init__(self, comm=None, device=None):
            self._ctx = None
            calls.append(("new", comm, device))

        def setOptionsPrefix(self, p):
            calls.append(("prefix", p))

        def setFromOptions(self):
            calls.append(("fromoptions",))

        def setOperators(self, A, P):
            calls.append(("operators", A, P))

        def init_pcd(self, asm, cls):
            self._ctx = object()
            calls.append(("init_pcd", asm, cls))

        def setTolerances(self, rtol=None, atol=None, max_it=None):
            calls.append(("tol", rtol, atol, max_it))

        def solve(self, b, x):
            calls.append(("solve", b, x))
            return 7

        def getResidualNorm(self):
            return 1e-9

        def getConvergedReason(self):
            return 2

        def getIterationNumber(self):
            return 7

    class FakeKSP:
        comm = "COMM"

        def __init__(self):
            self.its = self.reason = self.rnorm = None

        def getOptionsPrefix(self):
            return "ns_"

        def getOperators(self):
            return ("A", "P")

        def getTolerances(self):
            return (1e-7, 1e-50, 1e5, 300)

        def setIterationNumber(self, n):
            self.its = n

        def setConvergedReason(self, r):
            self.reason = r

        def setResidualNorm(self, r):
            self.rnorm = r

    ksp = FakeKSP()
    ctx = PCDKSPPython("ASM", pcd_pc_class="CLS", device=3, ksp_factory=FakeInner)
    ctx.create(ksp)
    ctx.setFromOptions(ksp)
    ctx.setUp(ksp)
    ctx.solve(ksp, "b", "x")
    ctx.setFromOptions(ksp)            # SNES calls it again: the prefix must not be touched after init_pcd
    ctx.setUp(ksp)                     # after a Jacobian update
    ctx.solve(ksp, "b2", "x2")
    kinds = [c[0] for c in calls]
    assert calls[0] == ("new", "COMM", 3)
    assert kinds.count("init_pcd") == 1 and kinds.count("prefix") == 1 and kinds.count("operators") == 2
    assert calls[kinds.index("prefix")] == ("prefix", "ns_") and kinds.index("prefix") < kinds.index("init_pcd")
    assert ("tol", 1e-7, 1e-50, 300) in calls and ("init_pcd", "ASM", "CLS") in calls
    assert (ksp.its, ksp.reason, ksp.rnorm) == (7, 2, 1e-9)
    with pytest.raises(RuntimeError):
        c2 = PCDKSPPython(ksp_factory=FakeInner)
        c2.create(ksp)
        c2.setUp(ksp)
