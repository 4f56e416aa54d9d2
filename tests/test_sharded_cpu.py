"""CPU, world_size 2 and 3 over gloo: the host side of the multi-GPU path -- contiguous row
sharding (ShardPlan), global index offsets, padding of short shards and the single all-gather
(exchange_candidates).  The per-shard search and the K5 merge are CUDA kernels and are covered
by the -m gpu tests; here each rank's local exact top-k comes from the oracle and the gathered
[G, Q, k] buffer is merged with a numpy restatement of K5's rule (descending similarity, ties
-> ascending global index), then compared with the oracle's global top-k."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hcir_b200.sharded import ShardPlan, exchange_candidates
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_numpy(g_s, g_i, k):
    """K5's rule on the gathered buffer: canonical order over the union, -1 = empty slot."""
    G, Q, _ = g_s.shape
    out_s = np.empty((Q, k), np.float32)
    out_i = np.empty((Q, k), np.int64)
    for q in range(Q):
        s = g_s[:, q, :].reshape(-1)
        i = g_i[:, q, :].reshape(-1)
        ok = i >= 0
        s, i = s[ok], i[ok]
        order = np.lexsort((i, -s))[:k]
        out_s[q], out_i[q] = s[order], i[order]
    return out_s, out_i


def _worker(rank, world, port, n, d, nq, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        bank = O.normalize(torch.randn(n, d, generator=g))
        bank[n // 2] = bank[1]  # an exact tie that straddles two shards
        qs = O.normalize(torch.randn(nq, d, generator=g))
        sp = ShardPlan(n, world)
        lo, hi = sp.start(rank), sp.stop(rank)
        kl = min(k, hi - lo)
        cs, ci = O.canonical_topk(qs, bank[lo:hi], kl)
        sims = torch.full((nq, k), float("-inf"))
        idx = torch.full((nq, k), -1, dtype=torch.int64)
        sims[:, :kl] = torch.from_numpy(cs)
        idx[:, :kl] = torch.from_numpy(ci) + lo  # global indices
        lab = (idx % 7).to(torch.int32)
        g_s, g_i, g_l = exchange_candidates(sims, idx, lab)
        assert g_s.shape == (world, nq, k) and g_i.shape == (world, nq, k) and g_l.shape == (world, nq, k)
        # rank r's slice of the gathered buffer is what rank r contributed
        assert torch.equal(g_i[rank], idx) and torch.equal(g_s[rank], sims)
        m_s, m_i = _merge_numpy(g_s.numpy(), g_i.numpy(), k)
        ref_s, ref_i = O.canonical_topk(qs, bank, k)
        np.testing.assert_array_equal(m_i, ref_i)
        np.testing.assert_array_equal(m_s, ref_s)
        ret[rank] = int(m_i.sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k", [(2, 1001, 10), (3, 100, 40)])
def test_sharded_exchange_and_merge_over_gloo(world, n, k):
    # (3, 100, 40): shards of 34/33/33 rows < k -> padded lists with (-inf, -1)
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, 32, 9, k, ret), nprocs=world, join=True)
    assert len(ret) == world and len(set(ret.values())) == 1  # every rank ends with the same result


def test_shard_plan_offsets_cover_global_index_space():
    sp = ShardPlan(10_000_000, 8)
    assert [sp.size(r) for r in range(8)] == [1_250_000] * 8
    sp = ShardPlan(11, 4)
    assert [(sp.start(r), sp.stop(r)) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 11)]
    assert [sp.owner(i) for i in range(11)] == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3]


def _qworker(rank, world, port, nq, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hcir_b200.sharded import gather_rows
        sp = ShardPlan(nq, world)
        full = torch.arange(nq * 3, dtype=torch.int64).view(nq, 3)
        mine = full[sp.start(rank):sp.stop(rank)]          # this rank's query slice "answers"
        out = gather_rows(mine, sp)
        assert torch.equal(out, full)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 10), (3, 10), (3, 2)])
def test_query_sharding_result_gather_over_gloo(world, nq):
    """Query-replica mode: unequal (and empty) per-rank query slices come back in query order."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_qworker, args=(world, port, nq, ret), nprocs=world, join=True)
    assert len(ret) == world


def test_choose_sharding_policy():
    from hcir_b200 import choose_sharding
    assert choose_sharding(200_000, 10_000, 8) == "query"        # C2: small gallery, many queries
    assert choose_sharding(10_000_000, 64, 8) == "gallery"       # C4: streaming regime
    assert choose_sharding(10_000_000, 16_384, 8, d=2048) == "gallery"  # C5: replica would not fit the budget
    assert choose_sharding(200_000, 1000, 8) == "gallery"        # too few query tiles per rank


def test_lane_rotation_rule(monkeypatch):
    """Pipelined submissions rotate over three lanes only where a rank's step is short (sharded._Lanes):
    the same decision and the same lane order on every rank, lane 0 = the caller's stream."""
    import torch
    from hcir_b200.sharded import QueryShardedGallery, ShardedGallery, _Lanes

    monkeypatch.delenv("HCIR_LANES", raising=False)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: object())

    def lanes_of(cls, nq, n_local, count=6):
        obj = cls.__new__(cls)
        obj.device = "cpu"
        obj._init_lanes()
        out = [obj._lane(nq, n_local) for _ in range(count)]
        assert all((st is None) == (lane == 0) for lane, st in out)
        return [lane for lane, _ in out]

    assert lanes_of(ShardedGallery, 64, 1_250_000) == [0, 1, 2, 0, 1, 2]         # C4 shard at 8 GPUs
    assert lanes_of(ShardedGallery, 4096, 125_000) == [0, 1, 2, 0, 1, 2]         # C3 in gallery shards at 8 GPUs
    assert lanes_of(ShardedGallery, 4096, 500_000) == [0] * 6                    # ... at 2 GPUs: long step
    assert lanes_of(QueryShardedGallery, 1250, 200_000) == [0, 1, 2, 0, 1, 2]    # C2 replicas at 8 GPUs
    assert lanes_of(QueryShardedGallery, 512, 1_000_000) == [0] * 6              # C3 replicas at 8 GPUs (measured: worse)
    monkeypatch.setenv("HCIR_LANES", "2")
    assert lanes_of(QueryShardedGallery, 512, 1_000_000, 4) == [0, 1, 0, 1]
    assert issubclass(ShardedGallery, _Lanes) and issubclass(QueryShardedGallery, _Lanes)
