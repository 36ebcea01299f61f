"""Multi-GPU plumbing for the two ways the CoreSLAM hot path shards.

* sessions (BASELINE cfg5): independent CoreSLAMProcessor instances, session i -> rank i mod world; no
  data-path communication at all (`session_shard`).
* candidates (BASELINE cfg4): the map is replicated, rank g evaluates the flat candidate indices
  `candidate_slice(T*I + 1, world, g)`, and ONE 8-byte exchange — a MIN all-reduce of the packed
  (distance << 32 | flat index) key — replaces the serial cross-thread arg-min of
  ParallelMonteCarloSearch (CoreSLAM/CoreSLAMProcessor.cs:694-705).  `SplitSearch` drives
  cs_update_begin -> all_reduce(MIN) -> cs_update_finish.

One process per GPU, `torch.distributed` for the exchange (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Nothing here computes distances or draws rays.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

INT32_MAX = 2147483647
KEY_EMPTY = (1 << 64) - 1  # armed key: no candidate evaluated yet (decodes to distance int.MaxValue)


def session_shard(n_sessions: int, world: int, rank: int) -> List[int]:
    """Sessions owned by `rank`: i mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_sessions, world))


def candidate_slice(n_flat: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, first+count) of the n_flat = T*I + 1 flat candidate indices (0 = searchPose) evaluated by
    `rank`: contiguous, [rank*n/world, (rank+1)*n/world).  Slices of all ranks partition [0, n_flat)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    lo = (rank * n_flat) // world
    hi = ((rank + 1) * n_flat) // world
    return lo, hi - lo


def pack_key(distance: int, flat_index: int) -> int:
    """(uint32 distance << 32) | uint32 flat index: unsigned order == the reference's tie-break order
    (strict <, lowest thread / earliest iteration wins; distance in [0, 2^31))."""
    return ((int(distance) & 0xFFFFFFFF) << 32) | (int(flat_index) & 0xFFFFFFFF)


def unpack_key(key: int) -> Tuple[int, int]:
    """-> (distance, flat index); the armed key decodes to (int.MaxValue, 0): 'searchPose wins'."""
    key = int(key) & KEY_EMPTY
    if key == KEY_EMPTY:
        return INT32_MAX, 0
    return (key >> 32) & 0xFFFFFFFF, key & 0xFFFFFFFF


def key_to_i64(key: int) -> int:
    """The exchange runs on int64 (NCCL/gloo MIN on a signed view).  Real keys are < 2^63 (distance <
    2^31); the armed key 2^64-1 would read as -1 and win every MIN, so it travels as int64 max."""
    key = int(key) & KEY_EMPTY
    return (1 << 63) - 1 if key >= (1 << 63) else key


def key_from_i64(v: int) -> int:
    v = int(v)
    return KEY_EMPTY if v == (1 << 63) - 1 else v


def pin_rank_to_cores(local_rank: int, local_world: int) -> List[int]:
    """One process per GPU on one node: give rank `local_rank` its own contiguous share of the cores this process may run on
    (os.sched_setaffinity), so that the ranks' pose-polling threads and staging copies do not migrate onto each other's
    cores.  Returns the cores kept (everything, untouched, when there are fewer cores than ranks or no affinity API)."""
    import os
    if not hasattr(os, "sched_getaffinity") or local_world <= 1:
        return sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
    cores = sorted(os.sched_getaffinity(0))
    per = len(cores) // local_world
    if per < 1:
        return cores
    mine = cores[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cores
    return mine


class _DevicePtr:
    """Zero-copy view of `count` int64 at a raw CUDA address for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, count: int = 1):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def device_key_tensor(ptr: int, device: int):
    """torch int64[1] tensor aliasing the 8-byte packed key in the session (no copy)."""
    import torch
    return torch.as_tensor(_DevicePtr(ptr), device=torch.device("cuda", device))


def allreduce_min_key(t, group=None):
    """The one exchange step of the candidate split: in-place MIN all-reduce of the int64 key tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def attach_group(proc, rank: int, world: int, group=None):
    """Candidate-split group over the ranks of a torch.distributed job (one process per GPU): every rank exports the handle
    of its exchange table, the 64-byte handles are all-gathered (this is the only use of the process group: the exchange
    itself runs inside the search kernel over peer-mapped memory, see include/coreslam_b200.h), every rank attaches.  After
    this, `proc.update(...)` / `proc.replay(...)` evaluate this rank's candidate slice and end on the group's winner."""
    import torch
    import torch.distributed as dist
    mine = proc.group_export()
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        proc.group_attach(0, 1, [mine])
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    proc.group_attach(rank, world, [bytes(x.cpu().tolist()) for x in out])
    dist.barrier(group=group)  # nobody starts exchanging before every rank has mapped every table


class SplitSearch:
    """Candidate-split Update over the ranks of a torch.distributed group: every rank holds a replica of
    the map (a `Processor` on its GPU, created on `torch_stream`), evaluates its slice of the candidates
    and takes part in the 8-byte exchange; all ranks end each Update with the same pose and bit-identical
    maps.  A rank's key is always < 2^63 after its search (every evaluated candidate contributes at least
    (int.MaxValue << 32 | index)), so the MIN runs directly on the int64 view of the in-session key."""

    def __init__(self, proc, rank: int, world: int, device: int, torch_stream=None, group=None):
        if proc.n_cand + 1 < world:
            raise ValueError("fewer candidates than ranks")
        self.proc, self.rank, self.world, self.device, self.group = proc, rank, world, device, group
        self.stream = torch_stream
        self.first, self.count = candidate_slice(proc.n_cand + 1, world, rank)
        self._views = {}

    def update(self, points, odometry_pose, cand_offsets=None):
        import contextlib
        import torch
        ptr = self.proc.update_begin(points, odometry_pose, cand_offsets, self.first, self.count)
        if ptr:  # 0: this scan is integrated without a search (scanCount < PositionSearchBeginning)
            key = self._views.get(ptr)
            if key is None:
                key = self._views[ptr] = device_key_tensor(ptr, self.device)
            ctx = torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
            with ctx:
                allreduce_min_key(key, self.group)
        return self.proc.update_finish()
