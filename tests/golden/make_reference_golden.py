"""Golden vectors produced by the REFERENCE'S OWN code for the part of the hot path it owns.

/root/reference/fenapack/preconditioners.py and field_split_backend.py are executed here, unmodified,
on top of the numpy/scipy stand-ins of tests/golden/refstubs.py (petsc4py and DOLFIN are not
installed; see that file for what is real and what is a stand-in).  The run goes through the
python-PC protocol exactly as PETSc drives it -- create(pc), init_pcd(PCDInterface), setUp(pc),
apply(pc, x, y) -- in the reference's DEFAULT configuration (PREONLY + Cholesky inner solves), for
PCDPC_BRM1, PCDPC_BRM2, PCDRPC_BRM1, PCDRPC_BRM2, with shallow and deep sub-matrices.

Inputs: backward-facing step, level 2, nu = 0.02, Oseen wind = the Stokes solution; PCDR: dt = 0.2.
The matrices come from oracle/fem.py (the reference assembles them with DOLFIN), embedded in the
mixed space with the DOLFIN-like interleaved numbering of oracle.problems.interleaved_index_sets.

Output: tests/golden/ref_pcd_apply.npz -- x, the reference's y per class, the Rp matrix the
reference builds, and the BC index list its SubfieldBC mapping produces; ref_pcd_apply_more.npz -- the
same for PCDPC_BRM1/2 on BFS level 3 and on a 3D channel (6 x 2 x 3 bricks).  tests/test_reference_golden.py
checks the oracle against these (CPU) and the CUDA library against them (GPU).  Needs /root/reference:
    python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import refstubs as rs  # noqa: E402
from oracle import fem, petsc_algos as pa, problems  # noqa: E402

LEVEL, NU, DT = 2, 0.02, 0.2


def build_problem(variant, pcdr):
    p0, _ = problems.backward_facing_step(LEVEL, nu=NU, variant=variant)
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    wind = x[:p0.n_u].reshape(-1, 2)
    idt = 1.0 / DT if pcdr else 0.0
    prob, space = problems.backward_facing_step(LEVEL, nu=NU, variant=variant, wind=wind, idt=idt, pcdr=pcdr)
    return prob, space, idt


def reference_apply(pcmod, fsb, cls_name, prob, space, idt, x, deep):
    n = prob.n_u + prob.n_p
    is_u, is_p = prob.is_u, prob.is_p
    blocks = {"ap": prob.Ap, "mp": prob.Mp, "kp": prob.Kp}
    if idt:
        asm = fem.Assembler(space)
        blocks["mu"] = asm.velocity_block(asm.p2_scalar(mass_coeff=idt))
    bc_mixed = {int(is_p[i]): float(v) for i, v in zip(prob.bc_idx, prob.bc_val)}
    # a velocity Dirichlet dof as well: SubfieldBC must ignore dofs outside the index set
    bc_mixed[int(is_u[0])] = 7.0
    assembler = rs.PCDAssembler(n, is_u, is_p, blocks, bc_mixed)
    # monolithic system matrix (PCDR takes Bt from it: gp is a phantom form)
    M = sp.bmat([[prob.A00, prob.A01], [prob.A10, None]], format="coo")
    perm = np.concatenate([is_u, is_p])
    A = rs.Mat(sp.csr_matrix((M.data, (perm[M.row], perm[M.col])), shape=(n, n)))
    interface = fsb.PCDInterface(assembler, A, rs.IS(is_u), rs.IS(is_p), deep_submats=deep)
    pc = rs._PCHandle("fieldsplit_p_")
    ctx = getattr(pcmod, cls_name)()
    ctx.create(pc)
    ctx.setFromOptions(pc)
    ctx.init_pcd(interface)
    ctx.setUp(pc)
    xv, yv = rs.Vec(x), rs.Vec(np.full_like(x, np.nan))
    ctx.apply(pc, xv, yv)
    assert np.array_equal(xv.array, x), "apply must not modify x"
    # second PCSetUp + apply (what a Newton step does): Ap/Mp are not re-assembled, Kp is
    calls_before = list(assembler.calls)
    ctx.setUp(pc)
    again = [c for c in assembler.calls[len(calls_before):]]
    y2 = rs.Vec(np.zeros_like(x))
    ctx.apply(pc, xv, y2)
    assert np.array_equal(y2.array, yv.array)
    extra = {"refresh_assembles": np.array(again)}
    extra["prefixes"] = np.array([ctx.ksp_Ap.getOptionsPrefix(), ctx.ksp_Mp.getOptionsPrefix(), ctx.mat_Kp.getOptionsPrefix()])
    sub = interface._subbcs[0]
    extra["bc_idx"], extra["bc_val"] = sub.idx, sub.val
    if hasattr(ctx, "ksp_Rp"):
        Rp = ctx.ksp_Rp.getOperators()[0].csr.tocsr()
        Rp.sort_indices()
        extra["Rp_indptr"], extra["Rp_indices"], extra["Rp_data"] = Rp.indptr, Rp.indices, Rp.data
    return yv.array.copy(), extra


def assembler_inputs():
    """A small mixed space (26 velocity + 14 pressure dofs, interleaved) with random sparse
    'forms' as host callables; velocity BCs with non-zero values, PCD BCs on two pressure dofs."""
    rng = np.random.default_rng(11)
    n = 40
    is_p = np.arange(2, n, 3)[:14]
    is_u = np.setdiff1d(np.arange(n), is_p)

    def rand(seed, rows, cols, density=0.3):
        r = np.random.default_rng(seed)
        M = sp.random(n, n, density=density, random_state=r, format="coo")
        keep = np.isin(M.row, rows) & np.isin(M.col, cols)
        M = sp.coo_matrix((M.data[keep], (M.row[keep], M.col[keep])), shape=(n, n)).tocsr()
        return (M + sp.diags(np.isin(np.arange(n), np.intersect1d(rows, cols)).astype(float))).tocsr()
    mats = {"a": rand(1, np.arange(n), np.arange(n)), "a_pc": rand(2, np.arange(n), np.arange(n)),
            "mp": rand(3, is_p, is_p), "mu": rand(4, is_u, is_u), "ap": rand(5, is_p, is_p),
            "fp": rand(6, is_p, is_p), "kp": rand(7, is_p, is_p), "gp": rand(8, is_u, is_p)}
    Lvec = rng.standard_normal(n)
    forms = {k: (lambda M=M: M.copy()) for k, M in mats.items()}
    forms["L"] = lambda: Lvec.copy()
    bcs = [rs.HostBC(is_u[[0, 3, 7]], [1.5, -2.0, 0.25]), rs.HostBC(is_u[[10]], [4.0])]
    bcs_pcd = [rs.HostBC(is_p[[1, 5]], [0.0, 0.0])]
    x_newton = rng.standard_normal(n)
    return n, forms, bcs, bcs_pcd, x_newton


def run_assembler(cls, n, forms, bcs, bcs_pcd, x_newton):
    """Drive a PCDAssembler class (the reference's or the drop-in) through every public method."""
    asm = cls(forms["a"], forms["L"], bcs, forms["a_pc"], mp=forms["mp"], mu=forms["mu"], ap=forms["ap"],
              fp=forms["fp"], kp=forms["kp"], gp=forms["gp"], bcs_pcd=bcs_pcd)
    out = {}
    for name in ("system_matrix", "pc_matrix", "ap", "mp", "mu", "fp", "kp", "gp"):
        T = rs.HostTensor()
        getattr(asm, name)(T)
        out[name] = T.csr.toarray()
    b = rs.HostTensor(n)
    asm.rhs_vector(b)
    out["rhs"] = b.array.copy()
    xv = rs.HostTensor(n)
    xv.array[:] = x_newton
    b2 = rs.HostTensor(n)
    asm.rhs_vector(b2, xv)
    out["rhs_newton"] = b2.array.copy()
    keys = ("ap", "mp", "mu", "fp", "kp", "gp")
    out["flags_constant"] = np.array([asm.get_pcd_form(k).is_constant() for k in keys])
    out["flags_phantom"] = np.array([asm.get_pcd_form(k).is_phantom() for k in keys])

    def raises(f):
        try:
            f()
        except AttributeError:
            return True
        return False
    out["unknown_form_raises"] = np.array([raises(lambda: asm.get_pcd_form("nope"))])
    bare = cls(forms["a"], forms["L"], bcs)
    out["missing_form_is_none"] = np.array([bare.get_dolfin_form("fp") is None])
    out["default_pcd_bcs_empty"] = np.array([not raises(bare.pcd_bcs) and len(bare.pcd_bcs()) == 0])
    out["none_pcd_bcs_raises"] = np.array([raises(cls(forms["a"], forms["L"], bcs, bcs_pcd=None).pcd_bcs)])
    out["pc_matrix_without_a_pc_is_noop"] = np.array([bare.pc_matrix(rs.HostTensor()) is None])
    return out


def more_problems(variant):
    """Inputs of the second fixture file: a second BFS size and a 3D problem (row lengths 15 on the
    pressure space instead of 7, Chebyshev bounds of 3D P1 tetrahedra)."""
    p0, _ = problems.backward_facing_step(3, nu=NU, variant=variant)
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    bfs3 = problems.backward_facing_step(3, nu=NU, variant=variant, wind=x[:p0.n_u].reshape(-1, 2))
    ch3d = problems.channel(6, 2, 3, nu=NU, variant=variant)
    return {"bfs3": bfs3, "channel3d": ch3d}


def more_cases(pcmod, fsb):
    out = {}
    rng = np.random.default_rng(11)
    for cls_name, variant in (("PCDPC_BRM1", "BRM1"), ("PCDPC_BRM2", "BRM2")):
        for case, (prob, space) in more_problems(variant).items():
            x = rng.standard_normal(prob.n_p)
            y, extra = reference_apply(pcmod, fsb, cls_name, prob, space, 0.0, x, deep=True)
            out[f"{case}_{cls_name}_x"], out[f"{case}_{cls_name}_y"] = x, y
            out[f"{case}_{cls_name}_bc_idx"] = extra["bc_idx"]
            print(case, cls_name, "n_p", prob.n_p, "|y|", np.linalg.norm(y))
    return out


def main():
    asm_mod = rs.load_reference_assembling()
    ref = run_assembler(asm_mod.PCDAssembler, *assembler_inputs())
    np.savez_compressed(os.path.join(HERE, "ref_assembler.npz"), **ref)
    print("PCDAssembler protocol:", {k: v.tolist() for k, v in ref.items() if v.size <= 6})
    pcmod, fsb = rs.load_reference_modules()
    out = {"level": np.array([LEVEL]), "nu": np.array([NU]), "dt": np.array([DT])}
    rng = np.random.default_rng(7)
    for cls_name, variant, pcdr in (("PCDPC_BRM1", "BRM1", False), ("PCDPC_BRM2", "BRM2", False),
                                    ("PCDRPC_BRM1", "BRM1", True), ("PCDRPC_BRM2", "BRM2", True)):
        prob, space, idt = build_problem(variant, pcdr)
        x = rng.standard_normal(prob.n_p)
        y_shallow, extra = reference_apply(pcmod, fsb, cls_name, prob, space, idt, x, deep=False)
        y_deep, _ = reference_apply(pcmod, fsb, cls_name, prob, space, idt, x, deep=True)
        assert np.array_equal(y_shallow, y_deep)
        out[f"{cls_name}_x"], out[f"{cls_name}_y"] = x, y_shallow
        for k, v in extra.items():
            out[f"{cls_name}_{k}"] = v
        print(cls_name, "n_p", prob.n_p, "|y|", np.linalg.norm(y_shallow), "refresh assembles", list(extra["refresh_assembles"]))
    np.savez_compressed(os.path.join(HERE, "ref_pcd_apply.npz"), **out)
    np.savez_compressed(os.path.join(HERE, "ref_pcd_apply_more.npz"), **more_cases(pcmod, fsb))


if __name__ == "__main__":
    main()
