"""Multi-GPU parity worker (one rank per GPU; launched by tests/test_multigpu.py or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/mgpu_worker.py
Every sharded answer must be BIT-identical to the single-GPU answer of the same inputs, for both
partitions (gallery rows / query replicas) and both exchanges (the library's push over NVLink peer
memory / one NCCL all-gather), including the steps that take the uncertified-completion branch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import hcir_b200  # noqa: E402
from hcir_b200 import synth  # noqa: E402
from hcir_b200.sharded import QueryShardedGallery, ShardPlan, ShardedGallery  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    checks = 0
    reps = 3 if world <= 2 else 2

    def case(bank, bl, qs, k, tag, expect_uncertified=False):
        nonlocal checks
        n = bank.shape[0]
        ref = hcir_b200.GalleryBank(bank, bl, device=dev)
        p_ref, s_ref, i_ref = ref.predict(qs, k, return_neighbors=True)
        if expect_uncertified:
            assert ref.last_stats["uncertified"] > 0, tag
        sp = ShardPlan(n, world)
        lo, hi = sp.start(rank), sp.stop(rank)
        for exchange in ("peer", "nccl"):
            gal = ShardedGallery(bank[lo:hi], bl[lo:hi], n_total=n, device=dev, classes=ref.classes_,
                                 exchange=exchange)
            for rep in range(reps):   # replays alternate the two parities of the peer region
                s, i = gal.topk(qs, k)
                assert torch.equal(i, i_ref) and torch.equal(s, s_ref), (tag, exchange, "gallery topk", rep)
                p = gal.predict(qs, k)
                assert torch.equal(p, p_ref), (tag, exchange, "gallery predict", rep)
            if expect_uncertified:
                assert gal.bank.last_stats["uncertified"] > 0, (tag, exchange)
            qg = QueryShardedGallery(bank, bl, device=dev, classes=ref.classes_, exchange=exchange)
            for rep in range(reps):
                p = qg.predict(qs, k)
                assert torch.equal(p, p_ref), (tag, exchange, "query predict", rep)
                p = qg.predict(qs.cuda(), k)
                assert torch.equal(p.cpu(), p_ref), (tag, exchange, "query predict (device queries)", rep)
                s, i = qg.topk(qs, k)          # retrieval through the replicas: packed rows travel over the peers
                assert torch.equal(i, i_ref) and torch.equal(s, s_ref), (tag, exchange, "query topk", rep)
            if exchange == "peer":   # pipelined submission: two steps in flight, checks read one step late
                qc = qs.cuda()
                for obj in (gal, qg):
                    pend = [obj.submit_predict(qc, k) for _ in range(3)]
                    for h in pend:
                        assert torch.equal(h.result().cpu(), p_ref), (tag, "submit_predict", type(obj).__name__)
                        assert h.redone == bool(expect_uncertified), (tag, "redone", type(obj).__name__)
                for obj in (gal, qg):
                    pend = [obj.submit_topk(qc, k) for _ in range(3)]
                    for h in pend:
                        s, i = h.result()
                        assert torch.equal(i.cpu(), i_ref) and torch.equal(s.cpu(), s_ref), (tag, "submit_topk", type(obj).__name__)
                # host batches through the three-stream serving loop: host answers, on every rank
                qh = qs.pin_memory()
                for obj, want in ((qg, "topk"), (qg, "pred"), (gal, "pred"), (gal, "topk")):
                    pipe = hcir_b200.HostPipeline.for_gallery(obj, qs.shape[0], k, want=want)
                    pend = [pipe.submit(qh) for _ in range(3)]
                    for h in pend:
                        r = h.result()
                        if want == "topk":
                            assert torch.equal(r[1], i_ref) and torch.equal(r[0], s_ref), (tag, "host pipeline topk")
                        else:
                            assert torch.equal(r, p_ref), (tag, "host pipeline pred", type(obj).__name__)
                        assert h.redone == bool(expect_uncertified), (tag, "host pipeline redone")
                pt_ref = ref.predict(qs, k, T=0.07)   # the temperature vote through both partitions
                assert torch.equal(gal.predict(qs, k, T=0.07), pt_ref) and torch.equal(qg.predict(qs, k, T=0.07), pt_ref)
            checks += 1
            del gal, qg
        del ref
        torch.cuda.empty_cache()

    bank, bl = synth.make_clustered(40000, 256, 13, 91)
    qs, _ = synth.make_clustered(1001, 256, 13, 92)          # ragged: 1001 queries over `world` ranks
    case(bank, bl, qs, 20, "clustered")
    # one query tile: pipelined submissions over gallery shards alternate between two lanes (streams), each
    # with its own captured session and peer channel
    case(bank, bl, qs[:64], 20, "streaming regime (two lanes)")
    if world <= 4:
        case(bank[:30011], bl[:30011], qs[:257], 7, "ragged gallery")
    # near-duplicate gallery: every query is uncertified -> the collective completion branch
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 256, generator=g)
    nd = max(16384, 4096 * world)   # every shard stays on the tensor path
    dup = base + 1e-4 * torch.randn(nd, 256, generator=g)
    qd = base + 1e-4 * torch.randn(256, 256, generator=g)
    case(dup, torch.arange(nd) % 5, qd, 10, "near-duplicates", expect_uncertified=True)

    # device-driven completion (on from k = 128): dense clusters make the first pass uncertified on every
    # rank; the in-graph second pass resolves them BEFORE the peers are signalled, so no rank takes the host
    # completion branch, nothing is redone, and the answers are still bit-identical
    g = torch.Generator().manual_seed(17)
    bases = torch.randn(8, 256, generator=g)
    dense = (bases[:, None, :] + 1e-3 * torch.randn(8, 400, 256, generator=g)).reshape(-1, 256)
    nb = 40000 + 3200
    cbank = torch.cat([torch.randn(40000, 256, generator=g), dense])[torch.randperm(nb, generator=g)]
    cl = torch.arange(nb) % 11
    cq = bases.repeat_interleave(8, 0) + 1e-3 * torch.randn(64, 256, generator=g)
    kd = 130
    ref = hcir_b200.GalleryBank(cbank, cl, device=dev)
    p_ref, s_ref, i_ref = ref.predict(cq, kd, return_neighbors=True)
    assert ref.last_stats["uncertified"] > 0
    sp = ShardPlan(nb, world)
    for exchange in ("peer", "nccl"):
        gal = ShardedGallery(cbank[sp.start(rank):sp.stop(rank)], cl[sp.start(rank):sp.stop(rank)], n_total=nb,
                             device=dev, classes=ref.classes_, exchange=exchange)
        qg = QueryShardedGallery(cbank, cl, device=dev, classes=ref.classes_, exchange=exchange)
        for rep in range(reps):
            s, i = gal.topk(cq, kd)
            assert torch.equal(i, i_ref) and torch.equal(s, s_ref), ("device completion", exchange, "gallery topk", rep)
            assert gal.bank.last_stats["uncertified"] == 0, gal.bank.last_stats
            assert torch.equal(gal.predict(cq, kd), p_ref) and torch.equal(qg.predict(cq, kd), p_ref)
            assert qg.bank.last_stats["uncertified"] == 0, qg.bank.last_stats
            s, i = qg.topk(cq, kd)
            assert torch.equal(i, i_ref) and torch.equal(s, s_ref), ("device completion", exchange, "query topk", rep)
        if exchange == "peer":
            for obj in (gal, qg):
                h = obj.submit_predict(cq.cuda(), kd)
                assert torch.equal(h.result().cpu(), p_ref) and not h.redone
        # (a gallery shard holds only 1/world of every dense cluster and may certify at once; a replica cannot)
        first = qg.last_session.counters.tolist()[1]
        tot = torch.tensor([first], device=dev)
        dist.all_reduce(tot)
        assert int(tot.item()) > 0, "the first pass should have left uncertified queries on the replicas"
        gal.close()
        qg.close()
    del ref
    checks += 1

    # more shapes than the session cache holds: evicted sessions close their peer regions collectively
    ref = hcir_b200.GalleryBank(bank, bl, device=dev)
    sp = ShardPlan(bank.shape[0], world)
    gal = ShardedGallery(bank[sp.start(rank):sp.stop(rank)], bl[sp.start(rank):sp.stop(rank)], n_total=bank.shape[0],
                         device=dev, classes=ref.classes_, exchange="peer")
    for nq in (129, 130, 131, 132, 133, 134):
        assert torch.equal(gal.predict(qs[:nq], 20), ref.predict(qs[:nq], 20)), ("eviction", nq)
    checks += 1

    # batch sizes whose per-rank split differs (1024 -> 128 rows everywhere, 1023 -> one
    # rank gets 127) on the SAME object must not desynchronise the ranks' session caches
    qg = QueryShardedGallery(bank, bl, device=dev, classes=ref.classes_, exchange="peer")
    big, _ = synth.make_clustered(1024, 256, 13, 93)
    for nq in (1024, 1023, 1024, 1022, 1023):
        assert torch.equal(qg.predict(big[:nq], 20), ref.predict(big[:nq], 20)), ("ragged replay", nq)
    qg.close()
    gal.close()
    checks += 1

    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} cases={checks}")
    sys.stdout.flush()
    os._exit(0)   # captured collectives keep the communicator busy at teardown (see bench.py)


if __name__ == "__main__":
    main()
