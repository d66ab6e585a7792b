"""Operator / sub-matrix / BC plumbing between PCDAssembler and the device
library -- drop-in for fenapack/field_split_backend.py (PCDInterface).

Kept from the reference: assemble-once logic for constant forms (:230-263),
per-step re-assembly of non-constant ones into the *same* sub-matrix object
(deep sub-matrix, MAT_REUSE_MATRIX => same pattern, new values; :285-291,
331-334), a cache of the sub-field BCs (:294-308) whose index computation is
SubfieldBC::compute_subfield_bc (SubfieldBC.h:92-160: position of the
constrained dof inside the owned pressure index set + rank offset).
New: ``pcd_bc_indices()`` hands that index list to the device once instead of
calling a collective VecSetValues/VecAssembly per apply (SubfieldBC.h:162-182).
"""
from __future__ import annotations

from weakref import proxy

import numpy as np

from ._backend import PETSc
from .assembling import PCDAssembler


class PCDInterface(object):
    def __init__(self, pcd_assembler, A, is_u, is_p, deep_submats=False, device=0):
        assert isinstance(pcd_assembler, PCDAssembler)
        self.assembler = pcd_assembler
        try:
            self.A = proxy(A)
        except TypeError:
            self.A = A
        self.is_u = is_u
        self.is_p = is_p
        assert isinstance(deep_submats, bool)
        self.deep_submats = deep_submats
        self.device = device
        self.scratch = {}
        self._subbcs = None

    # -- boundary conditions ---------------------------------------------------
    def pcd_bc_indices(self):
        """(idx, values): PCD Dirichlet dofs in local pressure ("p" split) numbering."""
        if self._subbcs is None:
            pos = {int(g): k for k, g in enumerate(np.asarray(self.is_p.getIndices()))}
            idx, vals = [], []
            for bc in self.assembler.pcd_bcs():
                dofs = np.asarray(bc.dofs(), dtype=np.int64)
                v = np.broadcast_to(np.asarray(bc.values(), dtype=np.float64), dofs.shape)
                for d, val in zip(dofs, v):
                    k = pos.get(int(d))
                    if k is not None:          # only owned sub-field entries (SubfieldBC.h:145-155)
                        idx.append(k)
                        vals.append(val)
            self._subbcs = (np.asarray(idx, dtype=np.int32), np.asarray(vals, dtype=np.float64))
        return self._subbcs

    def pcd_bc_indices_global(self, comm=None):
        """The same index list in the GLOBAL numbering of the "p" split vector: local position plus
        the exclusive scan of the owned sizes over the ranks (SubfieldBC.h:138-140) -- what the
        reference hands to VecSetValues."""
        from ._comm import HostComm
        idx, vals = self.pcd_bc_indices()
        hc = HostComm(comm if comm is not None else self.is_p.comm)
        return idx.astype(np.int64) + hc.exscan(self.is_p.getLocalSize()), vals

    def apply_pcd_bcs(self, vec):
        """Apply bcs to an intermediate pressure vector of PCD pc (host-side
        equivalent of SubfieldBC::apply; the device path does this itself)."""
        idx, vals = self.pcd_bc_indices()
        vec.getArray()[idx] = vals

    # -- operators ---------------------------------------------------------------
    def _work_mat(self, key):
        m = self.scratch.get(key)
        if m is None:
            m = self.scratch[key] = PETSc.Mat(comm=getattr(self.is_p, "comm", None))
        return m

    def _assemble_operator_deep(self, key, assemble_func, isrow, iscol=None, submat=None):
        full = self._work_mat(key)
        assemble_func(full)
        return full.createSubMatrix(isrow, isrow if iscol is None else iscol, submat=submat)

    def _setup_mat(self, key, isrow, mat, iscol=None):
        form = self.assembler.get_pcd_form(key)
        if mat is None or not form.is_constant():
            return self._assemble_operator_deep(key, getattr(self.assembler, key), isrow, iscol, submat=mat)
        return None

    def setup_mat_Kp(self, mat=None):
        """Assemble the pressure convection matrix; returns None when it is
        constant and already assembled (reference :79-83)."""
        return self._setup_mat("kp", self.is_p, mat)

    def setup_mat_Ap(self, mat=None):
        return self._setup_mat("ap", self.is_p, mat)

    def setup_mat_Mp(self, mat=None):
        return self._setup_mat("mp", self.is_p, mat)

    def setup_mat_Fp(self, mat=None):
        return self._setup_mat("fp", self.is_p, mat)

    def setup_mat_Mu(self, mat=None):
        return self._setup_mat("mu", self.is_u, mat)

    def setup_mat_Bt(self, mat=None):
        form = self.assembler.get_pcd_form("gp")
        if mat is None or not form.is_constant():
            if form.is_phantom():
                return self.A.createSubMatrix(self.is_u, self.is_p, submat=mat)
            return self._assemble_operator_deep("gp", self.assembler.gp, self.is_u, self.is_p, submat=mat)
        return None
