"""torchrun worker for tests/test_gpu_multi.py::test_candidate_split_two_processes_nccl and for manual
`gpurun --gpus 2` checks: candidate-split replay on WORLD_SIZE GPUs vs the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam.net_b200 as sn  # noqa: E402
from slam.net_b200 import parallel as par, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_scans, P, size, phys, iters, threads = 14, 400, 512, 40.0, 128, 4
    rp = synth.make_replay(n_scans, P, phys, seed=5)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p = sn.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads, device=local, max_points=P, seed=99,
                         stream=stream.cuda_stream)
        use_group = len(sys.argv) > 1 and sys.argv[1] == "group"  # exchange inside the search kernel (cs_group_*) instead of NCCL
        if use_group:
            par.attach_group(p, rank, world)
        ss = p if use_group else par.SplitSearch(p, rank, world, local, torch_stream=stream)
        o = orc.Processor(phys, size, rp.odometry[0], 0.1, 0.17, iters, threads)
        for k in range(n_scans):
            philox = k % 2 == 1
            off = sn.philox_offsets(99, k, iters * threads, 0.1, 0.17) if philox else synth.candidate_offsets(8, k, iters * threads, 0.1, 0.17)
            r = ss.update(rp.points[k], rp.odometry[k], None if philox else off)
            o.update(rp.points[k], rp.odometry[k], off)
            assert np.array_equal(r.pose, o.pose), (rank, k, r.pose, o.pose)
            if k >= 5:
                assert (r.distance, r.index) == (o.last_distance, o.last_index)
        assert np.array_equal(p.map_download(), np.array(o.map.pixels))
        p.close()
    dist.barrier()
    if rank == 0:
        print("SPLIT_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
