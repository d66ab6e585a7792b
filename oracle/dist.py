"""Row-partition / halo-exchange logic of the multi-rank path, restated in Python on
torch.distributed (any backend; the CPU tests use gloo) -- TEST INFRASTRUCTURE.
Mirrors fenapack_b200/csrc/dist.cu:build_halo / halo_exchange, which in turn
reproduce PETSc's MPIAIJ layout (owned contiguous ranges, ghost columns fetched per
MatMult) that the reference inherits (fenapack/SubfieldBC.h:138-140)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist


class HaloPlan:
    """Plan for one operator given this rank's rows with GLOBAL column ids."""

    def __init__(self, A_local_rows, begins, rank):
        A = sp.csr_matrix(A_local_rows)
        self.rank, self.nranks = rank, len(begins) - 1
        b0, b1 = begins[rank], begins[rank + 1]
        self.n_own = b1 - b0
        cols = A.indices.astype(np.int64)
        owned = (cols >= b0) & (cols < b1)
        self.ghosts = np.unique(cols[~owned])                      # sorted => grouped by owner
        loc = np.where(owned, cols - b0, self.n_own + np.searchsorted(self.ghosts, cols))
        self.A = sp.csr_matrix((A.data, loc.astype(np.int32), A.indptr), shape=(A.shape[0], self.n_own + self.ghosts.size))
        owner = np.searchsorted(np.asarray(begins[1:]), self.ghosts, side="right")
        self.recv_count = np.bincount(owner, minlength=self.nranks)
        self.recv_off = np.concatenate([[0], np.cumsum(self.recv_count)[:-1]])
        request = self.ghosts - np.asarray(begins)[owner]             # local index at the owner
        # exchange the request lists: counts first, then the indices
        counts = torch.tensor(self.recv_count, dtype=torch.int64)
        all_counts = [torch.zeros_like(counts) for _ in range(self.nranks)]
        dist.all_gather(all_counts, counts)
        self.send_count = np.array([int(all_counts[q][rank]) for q in range(self.nranks)])
        self.send_idx = [np.zeros(0, dtype=np.int64)] * self.nranks
        reqs = []
        bufs = {}
        for q in range(self.nranks):
            if q == rank:
                continue
            if self.recv_count[q]:
                t = torch.from_numpy(np.ascontiguousarray(request[self.recv_off[q]:self.recv_off[q] + self.recv_count[q]]))
                reqs.append(dist.isend(t, q))
            if self.send_count[q]:
                bufs[q] = torch.zeros(int(self.send_count[q]), dtype=torch.int64)
                reqs.append(dist.irecv(bufs[q], q))
        for r in reqs:
            r.wait()
        for q, b in bufs.items():
            self.send_idx[q] = b.numpy().copy()

    def exchange(self, x_own):
        """Ghost values of x in plan order."""
        ghost = np.zeros(self.ghosts.size)
        reqs, bufs = [], {}
        for q in range(self.nranks):
            if q == self.rank:
                continue
            if self.send_count[q]:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x_own[self.send_idx[q]])), q))
            if self.recv_count[q]:
                bufs[q] = torch.zeros(int(self.recv_count[q]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[q], q))
        for r in reqs:
            r.wait()
        for q, b in bufs.items():
            ghost[self.recv_off[q]:self.recv_off[q] + self.recv_count[q]] = b.numpy()
        return ghost

    def spmv(self, x_own):
        return self.A @ np.concatenate([x_own, self.exchange(x_own)])


def global_dot(a, b):
    t = torch.tensor([float(a @ b)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def split_rows(n, world, align=1):
    return [((n // align) * r // world) * align for r in range(world)] + [n]
