"""Island sharding across ranks: whole connected components are assigned to ranks and every rank
simulates only the bodies it owns.  What crosses NVLink is the north star's "all-gather of boundary body states":

  * HALO bodies -- bodies of other ranks inside (or near) the bounding box of this rank's bodies -- are tracked here
    for proximity detection only; their states (13 floats: t3 q4 linvel3 angvel3) are all-gathered EVERY step and
    imported before the next one.  A rank with no halo takes part in no per-step collective at all.
  * every `refresh_every` (128) steps the owned states of ALL bodies are all-gathered, every rank re-derives its halo set
    from them (bounding box of its own bodies, inflated by a margin that covers `refresh_every` steps at the fastest
    body's current speed) and tells the others which of their bodies it now tracks.
  * a CONTACT between an owned and a halo body means two shards' islands merged: the device raises RB_ERR_SHARD
    (reported by the next synchronising call); nothing is silently simulated wrong.
  * `finish()` gathers everything once more, so all ranks end with the exact state of every body.

The reference is single-process (one awake set, src/dynamics/island_manager/manager.rs:20-25); the components it
tracks as persistent islands (island_manager/persistent.rs:1-3) are the unit of distribution.  Works with any
torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests (host emulation of the kernels).
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


def state_tensor(pipe, device):
    """Zero-copy torch view [nb, 13] of the library's packed body-state buffer."""
    ptr, nbytes = pipe.state_buffer()
    if device.type == "cuda":
        t = torch.as_tensor(_CudaBuf(ptr, nbytes), device=device)
    else:
        t = torch.frombuffer((C.c_float * (nbytes // 4)).from_address(ptr), dtype=torch.float32)
    return t.view(-1, 13)


def body_radii(pipe):
    """Bounding radius of each body about its origin (from the collider descriptors the pipeline uploaded)."""
    r = np.zeros(pipe.nb, np.float32)
    for c in pipe._c:
        if c.parent < 0:
            continue
        he = np.array(c.half_extents[:], np.float64)
        ext = float(he[0]) if c.shape == 0 else float(np.linalg.norm(he))
        off = float(np.linalg.norm(np.array(c.pos_wrt_parent_t[:], np.float64)))
        r[c.parent] = max(r[c.parent], off + ext)
    return r


class IslandShard:
    def __init__(self, pipe, dist, rank, world_size, device, refresh_every=128, min_margin=0.5, overlap=None):
        self.pipe, self.dist, self.rank, self.world_size, self.device = pipe, dist, rank, world_size, device
        self.refresh_every = int(refresh_every)
        self.min_margin = float(min_margin)
        self.dt = float(pipe.params.dt)
        self.tick = 0
        comp = pipe.label_components()
        self.owner = partition_components(comp, world_size)
        pipe.set_owned_bodies((self.owner == rank).astype(np.uint8))
        nb = pipe.nb
        self.state = state_tensor(pipe, device)
        owner_t = torch.from_numpy(self.owner.astype(np.int64)).to(device)
        self.mine = owner_t == rank
        self.foreign = (owner_t >= 0) & ~self.mine
        self.radius = torch.from_numpy(body_radii(pipe)).to(device)
        idx = [np.nonzero(self.owner == r)[0].astype(np.int64) for r in range(world_size)]
        self.counts = [len(i) for i in idx]
        self.maxc = max(max(self.counts), 1)
        self.my_idx = torch.from_numpy(idx[rank]).to(device)
        # full exchange: packed owned rows of every rank -> rows of the state table
        self.full_send = torch.zeros(self.maxc, 13, device=device)
        self.full_recv = torch.zeros(world_size * self.maxc, 13, device=device)
        others = [r for r in range(world_size) if r != rank]
        self.other_idx = torch.cat([torch.from_numpy(idx[r]) for r in others] or [torch.zeros(0, dtype=torch.int64)]).to(device)
        self.other_rows = torch.cat([torch.arange(self.counts[r]) + r * self.maxc for r in others] or [torch.zeros(0, dtype=torch.int64)]).to(device)
        self.halo = torch.zeros(nb, dtype=torch.uint8, device=device)   # foreign bodies tracked here
        self.owner_t = owner_t
        self.export_max = 0          # per-step halo exchange: rows per rank (0 = no per-step collective)
        self.halo_steps = 0          # steps that needed the per-step halo exchange (diagnostic)
        self.refreshes = 0
        self.refresh(classify=True, force=True)

    # ---- full exchange (every refresh_every steps, at start-up and in finish) ----
    def _gather_all(self):
        n = self.counts[self.rank]
        self.full_send[:n] = self.state.index_select(0, self.my_idx)
        self.dist.all_gather_into_tensor(self.full_recv, self.full_send)
        if len(self.other_idx):
            self.state.index_copy_(0, self.other_idx, self.full_recv.index_select(0, self.other_rows))

    def refresh(self, classify=True, force=False):
        """All ranks get all states; each re-derives the bodies of other ranks it has to track (its halo) and, when any
        rank's halo changed, learns which of its own bodies the others track (its exports).  One host synchronisation."""
        self._gather_all()
        if not classify:
            return
        self.refreshes += 1
        pos, lin = self.state[:, 0:3], self.state[:, 7:10]
        speed = lin.norm(dim=1).max() if lin.numel() else torch.zeros((), device=self.device)
        margin = self.min_margin + 2.0 * self.refresh_every * self.dt * speed
        r = self.radius.unsqueeze(1)
        lo_all, hi_all = pos - r, pos + r
        if bool(self.counts[self.rank]):
            lo = lo_all[self.my_idx].min(dim=0).values - margin
            hi = hi_all[self.my_idx].max(dim=0).values + margin
            halo = ((hi_all >= lo).all(dim=1) & (lo_all <= hi).all(dim=1) & self.foreign).to(torch.uint8)
        else:
            halo = torch.zeros_like(self.halo)
        changed = (halo != self.halo).any().to(torch.int32).reshape(1)
        self.dist.all_reduce(changed, op=self.dist.ReduceOp.MAX)
        if not force and int(changed.item()) == 0:     # (the host synchronisation of a refresh)
            return
        self.halo = halo
        self.pipe.set_halo_bodies(self.halo.data_ptr())
        self.pipe.import_halo()
        # who tracks whom: all halo masks -> my export list and everybody's (for unpacking)
        masks = torch.zeros(self.world_size, self.halo.numel(), dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(masks.view(-1), self.halo)
        tracked = masks.any(dim=0)                      # bodies tracked by a rank other than their owner
        self.export_idx = []
        counts = []
        for rk in range(self.world_size):
            e = torch.nonzero(tracked & (self.owner_t == rk)).flatten()
            self.export_idx.append(e)
            counts.append(int(e.numel()))
        self.export_max = max(counts) if counts else 0
        if self.export_max > 0:
            self.halo_send = torch.zeros(self.export_max, 13, device=self.device)
            self.halo_recv = torch.zeros(self.world_size * self.export_max, 13, device=self.device)
            self.halo_dst = torch.cat([e for rk, e in enumerate(self.export_idx) if rk != self.rank] or [torch.zeros(0, dtype=torch.int64, device=self.device)])
            self.halo_rows = torch.cat([torch.arange(int(e.numel()), device=self.device) + rk * self.export_max
                                        for rk, e in enumerate(self.export_idx) if rk != self.rank] or [torch.zeros(0, dtype=torch.int64, device=self.device)])

    # ---- per step ----
    def exchange(self):
        """Call after every step: boundary (halo) states every step, everything every `refresh_every` steps."""
        self.tick += 1
        if self.tick % self.refresh_every == 0:
            self.refresh(classify=True)
            return
        if self.export_max == 0:
            return
        self.halo_steps += 1
        mine = self.export_idx[self.rank]
        if mine.numel():
            self.halo_send[:mine.numel()] = self.state.index_select(0, mine)
        self.dist.all_gather_into_tensor(self.halo_recv, self.halo_send)
        if self.halo_dst.numel():
            self.state.index_copy_(0, self.halo_dst, self.halo_recv.index_select(0, self.halo_rows))
        self.pipe.import_halo()

    def finish(self):
        """Afterwards every rank holds the current state of every body."""
        self.refresh(classify=False)
