"""Multi-rank host logic on CPU: world_size-2 `gloo` process groups.

The N>1 paths are (1) sessions sharded over ranks with no data-path collective and (2) the candidate
split with ONE 8-byte MIN all-reduce of the packed (distance, flat index) key.  Here the per-rank compute
is played by the CPU oracle (the checker), and the exchange / sharding code under test is exactly the one
the GPU path uses (slam.net_b200.parallel).  No GPU needed.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam.net_b200 import parallel as par
from slam.net_b200 import synth
from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, fn, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, q, args)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, val = q.get(timeout=300)
        out[r] = val
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return [out[r] for r in range(world)]


def _entry(rank, world, port, fn, q, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def test_candidate_slices_partition():
    for n in (1, 2, 7, 4097, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, cnt = par.candidate_slice(n, world, r)
                seen += list(range(lo, lo + cnt))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        par.candidate_slice(10, 2, 2)


def test_session_shards_partition():
    for n in (1, 5, 1024):
        for world in (1, 2, 4, 8):
            allv = sorted(sum((par.session_shard(n, world, r) for r in range(world)), []))
            assert allv == list(range(n))
            sizes = [len(par.session_shard(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_key_packing_order_is_reference_tie_break():
    # strict '<' with the lowest flat index winning ties (CoreSLAMProcessor.cs:644, :695-705)
    assert par.pack_key(5, 9) < par.pack_key(6, 0)
    assert par.pack_key(5, 3) < par.pack_key(5, 4)
    assert par.unpack_key(par.pack_key(33536000, 4096)) == (33536000, 4096)
    assert par.unpack_key(par.KEY_EMPTY) == (par.INT32_MAX, 0)
    assert par.pack_key(par.INT32_MAX, 7) < (1 << 63)  # every real key fits the signed exchange type
    assert par.key_from_i64(par.key_to_i64(par.KEY_EMPTY)) == par.KEY_EMPTY
    assert par.key_to_i64(par.pack_key(1, 2)) == par.pack_key(1, 2)


def _split_search_rank(rank, world, seed):
    """One scan of the candidate split: local arg-min over this rank's slice (oracle distances), the 8-byte
    exchange, decode."""
    size, phys, iters, threads, P = 128, 10.0, 64, 4, 90
    rng = np.random.default_rng(seed)  # same inputs on every rank
    m = orc.HoleMap(size, phys)
    m.pixels[:] = synth.random_map(size, seed)
    pts = rng.normal(0, 2.0, (P, 2)).astype(np.float32)
    sp = np.array([5.0, 5.0, 0.3], dtype=np.float32)
    off = rng.normal(0, [0.2, 0.2, 0.3], (iters * threads, 3)).astype(np.float32)
    best, bd, d, bi = orc.parallel_search(m, pts, sp, off, iters, threads)
    lo, cnt = par.candidate_slice(iters * threads + 1, world, rank)
    local = min(par.pack_key(int(d[i]), i) for i in range(lo, lo + cnt))
    t = torch.tensor([par.key_to_i64(local)], dtype=torch.int64)
    par.allreduce_min_key(t)
    dist_w, idx_w = par.unpack_key(par.key_from_i64(int(t[0])))
    return (dist_w, idx_w, int(bd), int(bi), lo, cnt)


@pytest.mark.parametrize("world", [2, 3])
def test_candidate_split_exchange_gloo(world):
    res = _run(world, _split_search_rank, 17)
    for dist_w, idx_w, bd, bi, lo, cnt in res:
        assert (dist_w, idx_w) == (bd, bi)  # every rank decodes the single-process winner
    assert sum(r[5] for r in res) == 64 * 4 + 1


def _all_out_of_bounds_rank(rank, world):
    # nothing in bounds anywhere: every candidate has distance int.MaxValue, flat index 0 (searchPose) wins
    n = 33
    lo, cnt = par.candidate_slice(n, world, rank)
    local = min(par.pack_key(par.INT32_MAX, i) for i in range(lo, lo + cnt))
    t = torch.tensor([par.key_to_i64(local)], dtype=torch.int64)
    par.allreduce_min_key(t)
    return par.unpack_key(par.key_from_i64(int(t[0])))


def test_candidate_split_all_misses_gloo():
    for r in _run(2, _all_out_of_bounds_rank):
        assert r == (par.INT32_MAX, 0)


def _session_rank(rank, world, n_sessions, n_scans):
    """Sessions sharded i mod world, no collective on the data path; the checksums are gathered only to be
    compared with the single-process run."""
    rp = synth.make_replay(n_scans, 60, 20.0, seed=3)
    mine = par.session_shard(n_sessions, world, rank)
    sums = torch.zeros(n_sessions, dtype=torch.int64)
    for s in mine:
        o = orc.Processor(20.0, 96, rp.odometry[0], 0.05 + 0.01 * s, 0.1, 16, 2)
        o.hole_width = 0.4 + 0.1 * s
        for k in range(n_scans):
            o.update(rp.points[k], rp.odometry[k], synth.candidate_offsets(100 + s, k, 32, 0.05 + 0.01 * s, 0.1))
        sums[s] = o.map.crc32()
    dist.all_reduce(sums)  # test-only gather (each slot is written by exactly one rank)
    return sums.tolist()


def test_session_sharding_gloo():
    n_sessions, n_scans = 5, 7
    res = _run(2, _session_rank, n_sessions, n_scans)
    assert res[0] == res[1]
    single = _session_rank_single(n_sessions, n_scans)
    assert res[0] == single


def _session_rank_single(n_sessions, n_scans):
    rp = synth.make_replay(n_scans, 60, 20.0, seed=3)
    out = []
    for s in range(n_sessions):
        o = orc.Processor(20.0, 96, rp.odometry[0], 0.05 + 0.01 * s, 0.1, 16, 2)
        o.hole_width = 0.4 + 0.1 * s
        for k in range(n_scans):
            o.update(rp.points[k], rp.odometry[k], synth.candidate_offsets(100 + s, k, 32, 0.05 + 0.01 * s, 0.1))
        out.append(o.map.crc32())
    return out


def test_pin_rank_to_cores_partitions_the_allowed_cores():
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    """parallel.pin_rank_to_cores: the ranks of a node get disjoint contiguous shares of the cores the process may use (run in
    child processes: the affinity of the test runner itself must not change)."""
    import subprocess
    import sys
    allowed = sorted(os.sched_getaffinity(0))
    world = 2 if len(allowed) >= 2 else 1
    got = []
    for r in range(world):
        code = ("import sys, os; sys.path.insert(0, %r); from slam.net_b200 import parallel as p; "
                "m = p.pin_rank_to_cores(%d, %d); assert sorted(os.sched_getaffinity(0)) == sorted(m); print(','.join(map(str, m)))" % (ROOT, r, world))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        got.append([int(x) for x in out.stdout.strip().split(",")])
    flat = [c for g in got for c in g]
    assert len(set(flat)) == len(flat) and set(flat) <= set(allowed)
    if world == 2:
        assert len(got[0]) == len(got[1]) == len(allowed) // 2
