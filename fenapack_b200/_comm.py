"""The four host-side collectives the drop-in layer needs, on whatever communicator the caller
hands to ``PCDKSP(comm)`` (reference field_split.py:46,75-77): a ``PETSc.Comm`` (``tompi4py()``), an
``mpi4py`` communicator, or the stand-ins of ``petsc_shim`` (serial, or ``TorchDistComm`` on
torch.distributed).  Used for set-up only: ownership offsets (the exscan of SubfieldBC.h:138-140),
the broadcast of the NCCL unique id, sizes.  The data path never goes through the host."""


class HostComm(object):
    def __init__(self, comm):
        c = comm.tompi4py() if hasattr(comm, "tompi4py") else comm
        self._c = c
        self.size = c.Get_size() if hasattr(c, "Get_size") else c.size
        self.rank = c.Get_rank() if hasattr(c, "Get_rank") else c.rank

    def bcast(self, obj, root=0):
        return self._c.bcast(obj, root=root) if self.size > 1 else obj

    def allgather(self, obj):
        return self._c.allgather(obj) if self.size > 1 else [obj]

    def exscan(self, value):
        """Sum over the lower ranks (0 on rank 0, where MPI leaves the result undefined)."""
        if self.size == 1:
            return 0
        r = self._c.exscan(value)
        return 0 if (r is None or self.rank == 0) else r

    def allreduce(self, value):
        return self._c.allreduce(value) if self.size > 1 else value
