"""One-sided halo transports for the row-slab engine (distributed.py).

The default ghost exchange is a grouped NCCL send/recv per refresh.  A transport replaces it by WRITES into the
neighbours' arrays over NVLink peer memory: every slab array lives in symmetric memory (the same allocation on every
rank, each rank's copy mapped into every other rank's address space), and a refresh is

    1. READY handshake   every kernel that writes my copy of the arrays has been issued (stream order) -> tell both
                         neighbours, wait for theirs: a neighbour's pass may still be storing its (stale) ghost rows,
                         and must not overwrite what is pushed next;
    2. PUSH              my first / last G owned rows -> the upper / lower ghost rows of the neighbours' copies
                         (plain device stores through the peer mapping; a contiguous block per array and neighbour);
    3. DATA handshake    tell both neighbours their ghost rows are complete, wait for mine.

All three are stream-ordered device work (no host synchronisation, no NCCL), so whole cycles still capture into CUDA
graphs.  The row bookkeeping (which of my rows land where in the neighbour's local array) is shared by all transports
and is exercised on the CPU by `FileShmTransport` (file-backed shared mappings + gloo barriers) in
tests/test_distributed_cpu.py; `SymmMemTransport` is the GPU implementation on torch's symmetric memory
(`torch.distributed._symmetric_memory`: `get_buffer` peer views, `put_signal` / `wait_signal` signal pads).

Status: opt-in (`DistributedCycleEngine(transport=...)`, `bench.py --halo p2p`); the NCCL path stays the default
until the peer path has been measured on 2-8 GPUs."""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


class PeerPushTransport:
    """Row bookkeeping + protocol; subclasses provide the symmetric allocation, the peer views and the signals."""

    READY, DATA = 0, 1

    def __init__(self, rank: int, world: int, ghost: int):
        self.rank, self.world, self.ghost = rank, world, ghost
        self._flat: List[torch.Tensor] = []                 # allocation k: my flat buffer
        self._meta: Dict[int, Tuple[int, int, int]] = {}    # data_ptr of a local view -> (k, rows_max, pitch)
        self.pushes = 0

    # -- to be provided ---------------------------------------------------------------------------------------
    def pitch(self, ny: int, dtype) -> int:
        return ny

    def _alloc_flat(self, k: int, n: int, dtype) -> torch.Tensor:
        raise NotImplementedError

    def _peer_flat(self, k: int, peer: int) -> torch.Tensor:
        raise NotImplementedError

    def _signal(self, peers: Sequence[int], channel: int) -> None:
        raise NotImplementedError

    def _wait(self, peers: Sequence[int], channel: int) -> None:
        raise NotImplementedError

    # -- allocation (collective: every rank calls it in the same order with the same sizes) -----------------------
    def alloc(self, loc_nx: int, rows_max: int, ny: int, dtype) -> torch.Tensor:
        """A zeroed (loc_nx, ny) field inside a symmetric allocation of rows_max rows (the tallest slab of any rank)."""
        if loc_nx > rows_max:
            raise ValueError("loc_nx exceeds the symmetric allocation")
        ldp = self.pitch(ny, dtype)
        k = len(self._flat)
        flat = self._alloc_flat(k, rows_max * ldp, dtype)
        flat.zero_()
        self._flat.append(flat)
        view = flat.view(rows_max, ldp)[:loc_nx, :ny]
        self._meta[view.data_ptr()] = (k, rows_max, ldp)
        return view

    def owns(self, t: torch.Tensor) -> bool:
        return t.data_ptr() in self._meta

    # -- the refresh ------------------------------------------------------------------------------------------
    def neighbours(self) -> List[int]:
        return [r for r in (self.rank - 1, self.rank + 1) if 0 <= r < self.world]

    def plan(self, slab, per: int) -> List[Tuple[int, Tuple[int, int], Tuple[int, int]]]:
        """[(peer, my rows [a, b), the peer's rows [c, d))] for one level: my first G owned rows go to the UPPER ghost
        rows of the lower neighbour, my last G owned rows to the LOWER ghost rows of the upper neighbour.  `per` =
        owned rows per rank on this level (the last rank owns one more, the boundary row, which nobody needs)."""
        G = self.ghost
        lo, hi = slab.own_local
        out = []
        if self.rank > 0:  # lower neighbour: its local array = [ghost below (none on rank 0)] + per owned + G ghost
            g_lo_n = 0 if self.rank - 1 == 0 else G
            out.append((self.rank - 1, (lo, lo + G), (g_lo_n + per, g_lo_n + per + G)))
        if self.rank < self.world - 1:  # upper neighbour: its local rows [0, G) are its lower ghost rows
            # my last G owned rows; on every rank but the last, hi - lo == per
            out.append((self.rank + 1, (hi - G, hi), (0, G)))
        return out

    def exchange(self, part, items) -> None:
        """Refresh the ghost rows of the (array, level) pairs `items` on every rank (collective, stream-ordered)."""
        if self.world == 1 or not items:
            return
        nb = self.neighbours()
        self._signal(nb, self.READY)
        self._wait(nb, self.READY)
        for t, l in items:
            k, rows_max, ldp = self._meta[t.data_ptr()]
            s = part.slab(l)
            per = (s.nx_glob - 1) // self.world
            mine = self._flat[k].view(rows_max, ldp)
            for peer, (a, b), (c, d) in self.plan(s, per):
                theirs = self._peer_flat(k, peer).view(rows_max, ldp)
                theirs[c:d].copy_(mine[a:b])  # one contiguous block, pitch padding included
                self.pushes += 1
        self._signal(nb, self.DATA)
        self._wait(nb, self.DATA)


class SymmMemTransport(PeerPushTransport):
    """GPU implementation on torch symmetric memory (CUDA IPC / NVLink peer mappings, signal pads)."""

    def __init__(self, device, ghost: int, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm, self.device = symm, torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        super().__init__(dist.get_rank(self.group), dist.get_world_size(self.group), ghost)
        try:  # older releases need the group enabled explicitly; newer ones deprecate the call
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self._handles: List = []
        self._peers: Dict[Tuple[int, int], torch.Tensor] = {}
        self._dtypes: List[torch.dtype] = []
        # a tiny allocation whose signal pads carry the handshakes of every exchange
        ctrl = symm.empty(64, dtype=torch.float32, device=self.device)
        self._ctrl = (ctrl, symm.rendezvous(ctrl, self.group))

    def pitch(self, ny: int, dtype) -> int:
        from .device import pitch_for
        return pitch_for(ny)

    def _alloc_flat(self, k, n, dtype):
        t = self.symm.empty(n, dtype=dtype, device=self.device)
        self._handles.append(self.symm.rendezvous(t, self.group))
        self._dtypes.append(dtype)
        return t

    def _peer_flat(self, k, peer):
        v = self._peers.get((k, peer))
        if v is None:
            n = self._flat[k].numel()
            v = self._peers[(k, peer)] = self._handles[k].get_buffer(peer, (n,), self._dtypes[k], 0)
        return v

    def _signal(self, peers, channel):
        for p in peers:
            self._ctrl[1].put_signal(p, channel)

    def _wait(self, peers, channel):
        for p in peers:
            self._ctrl[1].wait_signal(p, channel)


class FileShmTransport(PeerPushTransport):
    """CPU stand-in used by the gloo tests: every allocation is a file-backed shared mapping in `directory` (one file
    per allocation and rank), the handshakes are process-group barriers.  Same bookkeeping, same protocol order."""

    def __init__(self, directory: str, ghost: int, group=None):
        self.dir, self.group = directory, group
        super().__init__(dist.get_rank(group), dist.get_world_size(group), ghost)
        self._peers: Dict[Tuple[int, int], torch.Tensor] = {}
        self._sizes: List[Tuple[int, torch.dtype]] = []

    def _path(self, k, r):
        return os.path.join(self.dir, f"slab{k}_rank{r}.bin")

    def _alloc_flat(self, k, n, dtype):
        t = torch.from_file(self._path(k, self.rank), shared=True, size=n, dtype=dtype)
        self._sizes.append((n, dtype))
        dist.barrier(group=self.group)  # every rank's file exists and has its full size
        return t

    def _peer_flat(self, k, peer):
        v = self._peers.get((k, peer))
        if v is None:
            n, dtype = self._sizes[k]
            v = self._peers[(k, peer)] = torch.from_file(self._path(k, peer), shared=True, size=n, dtype=dtype)
        return v

    def _signal(self, peers, channel):
        pass

    def _wait(self, peers, channel):
        dist.barrier(group=self.group)  # the READY / DATA handshake of every pair at once
