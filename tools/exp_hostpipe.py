"""Where does a HostPipeline step spend its HOST time?  torchrun --nproc-per-node N tools/exp_hostpipe.py [Q]
Times submit() and result() on the host clock next to the device-paced step (C3 gallery, Q queries over N ranks)."""
import collections
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hcir_b200  # noqa: E402
from hcir_b200 import synth  # noqa: E402
from hcir_b200.sharded import QueryShardedGallery  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 512 * world
    n, d, k = 1_000_000, 768, 100
    bank, _ = synth.make_clustered(n, d, 100, 1234, device=dev)
    qs, _ = synth.make_clustered(q, d, 100, 4321, device=dev)
    gal = QueryShardedGallery(bank, None, device=dev)
    del bank
    qh = torch.empty((q, d), dtype=torch.float32, pin_memory=True)
    qh.copy_(qs)
    pipe = hcir_b200.HostPipeline.for_gallery(gal, q, k, want="topk")
    t_sub, t_res = [], []

    def steps(count, prof=None):
        pend = collections.deque()
        for _ in range(count):
            t0 = time.perf_counter()
            if prof:
                prof.enable()
            pend.append(pipe.submit(qh))
            if prof:
                prof.disable()
            t1 = time.perf_counter()
            if len(pend) >= 2:
                pend.popleft().result()
            t2 = time.perf_counter()
            t_sub.append(t1 - t0)
            t_res.append(t2 - t1)
        while pend:
            pend.popleft().result()

    steps(5)
    dist.barrier()
    torch.cuda.synchronize()
    t_sub.clear(), t_res.clear()
    t0 = time.perf_counter()
    steps(50)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50
    sub_avg, res_avg = 1e3 * sum(t_sub) / len(t_sub), 1e3 * sum(t_res) / len(t_res)
    # device-only pipelined pace for comparison
    qc = qs
    def dsteps(count):
        pend = collections.deque()
        for _ in range(count):
            pend.append(gal.submit_topk(qc, k))
            if len(pend) >= 2:
                pend.popleft().result()
        while pend:
            pend.popleft().result()
    dsteps(5)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dsteps(50); torch.cuda.synchronize()
    dwall = (time.perf_counter() - t0) / 50
    # the same loop reading back only this rank's slice of the answer (is the full-result D2H on every rank the cost?)
    from hcir_b200.sharded import ShardPlan
    sp = ShardPlan(q, world)
    a, b = sp.start(rank), sp.stop(rank)
    pipe_full = pipe
    pipe = hcir_b200.HostPipeline(lambda x: gal.submit_topk(x, k), q, d, device=dev, rows=(a, b),
                                  pick=lambda r: (r[0][a:b], r[1][a:b]))
    steps(5)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); steps(50); torch.cuda.synchronize()
    swall = (time.perf_counter() - t0) / 50
    pipe = pipe_full
    if rank == 0:
        print(f"own-slice read-back: {swall*1e3:.3f} ms/step")
        print(f"world {world} Q {q}: host-pipeline {wall*1e3:.3f} ms/step (submit {sub_avg:.3f} ms host, "
              f"result wait {res_avg:.3f} ms) | device-resident pipelined {dwall*1e3:.3f} ms/step")
        prof = cProfile.Profile()
        steps(30, prof)
        pstats.Stats(prof).sort_stats("cumulative").print_stats(28)
    else:
        steps(30)
    gal.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
