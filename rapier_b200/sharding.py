"""Island sharding across ranks: whole connected components are assigned to ranks,
every rank simulates only the bodies it owns, and one all-gather per step of the owned body
states (13 floats per body: t3 q4 linvel3 angvel3) gives every rank the full world state.

The reference is single-process (one awake set, src/dynamics/island_manager/manager.rs:20-25); the
components it already tracks as persistent islands (island_manager/persistent.rs:1-3) are the unit
of distribution here.  Works with any torch.distributed backend: NCCL over NVLink on the GPU box,
gloo in the CPU tests (where the world is the host emulation of the kernels).
"""
import ctypes as C

import numpy as np
import torch


def partition_components(component_of_body, world_size):
    """Contiguous blocks of components (ascending root id) per rank, balanced by body count.
    Returns owner[b] in [0, world_size) or -1 for non-dynamic bodies."""
    comp = np.asarray(component_of_body)
    roots, counts = np.unique(comp[comp >= 0], return_counts=True)
    cum = np.cumsum(counts) - counts
    total = max(int(counts.sum()), 1)
    owner_of_root = np.minimum((cum * world_size) // total, world_size - 1)
    lut = dict(zip(roots.tolist(), owner_of_root.tolist()))
    return np.array([lut[int(c)] if c >= 0 else -1 for c in comp], np.int32)


class _CudaBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _view(ptr, nbytes, device):
    if device.type == "cuda":
        t = torch.as_tensor(_CudaBuf(ptr, nbytes), device=device)
    else:
        t = torch.frombuffer((C.c_float * (nbytes // 4)).from_address(ptr), dtype=torch.float32)
    return t.view(-1, 13)


def state_tensor(pipe, device):
    """Zero-copy torch view [nb, 13] of the library's packed body-state buffer."""
    ptr, nbytes = pipe.state_buffer()
    if device.type == "cuda":
        t = torch.as_tensor(_CudaBuf(ptr, nbytes), device=device)
    else:
        t = torch.frombuffer((C.c_float * (nbytes // 4)).from_address(ptr), dtype=torch.float32)
    return t.view(-1, 13)


class IslandShard:
    """`overlap=True`: the all-gather of step n runs asynchronously (on NCCL's stream) while step n+1 is
    computed, and its result is imported before step n+2.  Bodies simulated by other ranks are then seen
    one step late -- they never touch this rank's components, the states only feed proximity detection
    and download -- and `finish()` drains the pipeline so every rank ends with the exact state of every
    body.  With contiguous equal shards the library double-buffers its packed state (step k writes buffer
    k & 1) and the gather runs IN PLACE on the buffer just written: per step one collective and one
    import kernel, no packing, nothing on the critical path but the import."""

    def __init__(self, pipe, dist, rank, world_size, device, overlap=False):
        self.pipe, self.dist, self.rank, self.world_size = pipe, dist, rank, world_size
        self.overlap = bool(overlap)
        self.device = device
        self.pending = None
        self.tick = 0
        comp = pipe.label_components()
        self.owner = partition_components(comp, world_size)
        pipe.set_owned_bodies((self.owner == rank).astype(np.uint8))
        self.state = state_tensor(pipe, device)
        idx = [np.nonzero(self.owner == r)[0].astype(np.int64) for r in range(world_size)]
        self.counts = [len(i) for i in idx]
        self.maxc = max(max(self.counts), 1)
        self.my_idx = torch.from_numpy(idx[rank]).to(device)
        self.send = torch.zeros(self.maxc, 13, device=device)
        self.recv = torch.zeros(world_size * self.maxc, 13, device=device)
        others = [r for r in range(world_size) if r != rank]
        self.imp_idx = torch.cat([torch.from_numpy(idx[r]) for r in others] or [torch.zeros(0, dtype=torch.int64)]).to(device).int()
        self.rows = torch.cat([torch.arange(self.counts[r]) + r * self.maxc for r in others] or [torch.zeros(0, dtype=torch.int64)]).to(device)
        self.imp_src = torch.zeros(len(self.imp_idx), 13, device=device)
        # Fast path: every rank owns one contiguous, equally long range of body indices, in rank order
        # (what weak-scaling replicas of a scene give).  The all-gather then runs IN PLACE on the library's
        # state buffer -- one collective and one import kernel per step, no packing.
        lo = [int(i[0]) if len(i) else -1 for i in idx]
        self.inplace = (len(set(self.counts)) == 1 and self.counts[0] > 0 and
                        all(np.array_equal(i, np.arange(l, l + len(i))) for i, l in zip(idx, lo)) and
                        all(lo[r + 1] == lo[r] + self.counts[r] for r in range(world_size - 1)))
        if self.inplace:
            n = self.counts[0]
            self.block = self.state[lo[0]:lo[0] + world_size * n].view(world_size, n * 13)
            self.block_flat = self.block.view(-1)
        if self.overlap and self.inplace:   # double-buffered library state: gather the buffer the step just wrote
            p0, p1, nbytes = pipe.state_buffers()
            self.tables = [_view(p0, nbytes, device), _view(p1, nbytes, device)]
            self.table_ptr = [p0, p1]
            n = self.counts[0]
            self.blocks = [t[lo[0]:lo[0] + world_size * n].view(world_size, n * 13) for t in self.tables]
        elif self.overlap:                  # packed: double-buffered snapshots / receive buffers
            self.send2 = [torch.zeros(self.maxc, 13, device=device) for _ in range(2)]
            self.recv2 = [torch.zeros(world_size * self.maxc, 13, device=device) for _ in range(2)]

    def _import_from(self, recv):
        if len(self.imp_idx):
            torch.index_select(recv, 0, self.rows, out=self.imp_src)
            self.pipe.import_states(self.imp_idx.data_ptr(), self.imp_src.data_ptr(), len(self.imp_idx))

    def finish(self):
        """Drain the asynchronous exchange: afterwards every rank holds the current state of every body."""
        if self.pending is not None:
            work, recv = self.pending
            work.wait()
            if isinstance(recv, int):
                if len(self.imp_idx):
                    self.pipe.import_states_from(self.imp_idx.data_ptr(), self.table_ptr[recv], len(self.imp_idx))
            else:
                self._import_from(recv)
            self.pending = None

    def exchange(self):
        """All-gather the owned body states and import the states simulated by the other ranks."""
        if self.inplace and hasattr(self, "tables"):
            k = self.table_ptr.index(self.pipe.state_buffer()[0])   # the buffer the step that was just enqueued writes
            if not self.overlap:   # (overlap switched off at run time: same buffers, synchronous)
                self.finish()
                self.dist.all_gather_into_tensor(self.blocks[k].view(-1), self.blocks[k][self.rank])
                if len(self.imp_idx):
                    self.pipe.import_states_from(self.imp_idx.data_ptr(), self.table_ptr[k], len(self.imp_idx))
                return
            work = self.dist.all_gather_into_tensor(self.blocks[k].view(-1), self.blocks[k][self.rank], async_op=True)
            previous, self.pending = self.pending, (work, k)
            if previous is not None:   # the previous step's gather: import it before the next step
                previous[0].wait()
                if len(self.imp_idx):
                    self.pipe.import_states_from(self.imp_idx.data_ptr(), self.table_ptr[previous[1]], len(self.imp_idx))
            return
        if self.overlap:
            send, recv = self.send2[self.tick & 1], self.recv2[self.tick & 1]
            self.tick += 1
            send[:self.counts[self.rank]] = self.state.index_select(0, self.my_idx)   # snapshot: the next step overwrites the rows
            work = self.dist.all_gather_into_tensor(recv, send, async_op=True)
            previous, self.pending = self.pending, (work, recv)
            if previous is not None:   # import what the previous step's gather brought, before the next step
                previous[0].wait()
                self._import_from(previous[1])
            return
        if self.inplace:
            self.dist.all_gather_into_tensor(self.block_flat, self.block[self.rank])
            if len(self.imp_idx):
                self.pipe.import_states(self.imp_idx.data_ptr(), 0, len(self.imp_idx))
            return
        self.send[:self.counts[self.rank]] = self.state.index_select(0, self.my_idx)
        self.dist.all_gather_into_tensor(self.recv, self.send)
        if len(self.imp_idx):
            torch.index_select(self.recv, 0, self.rows, out=self.imp_src)
            self.pipe.import_states(self.imp_idx.data_ptr(), self.imp_src.data_ptr(), len(self.imp_idx))
