"""Round-2 GPU parity tests (``-m gpu``): the fused tail of K3 (labels -> vote -> peer stores ->
arrival), the fused wait + merge + vote kernel, the peer exchange exercised on ONE GPU with fake
ranks, the certification bound with exactly kc staged keys, oracle checks of the BASELINE configs at
full size (a query subsample against ``torch.mm`` + ``topk`` on the host copy of the bank), and the
reference-shaped wrappers that had no test.  All calls go through the C ABI (ctypes)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import hcir_b200
from hcir_b200 import GalleryBank, KNeighborsClassifierB200, _lib, synth
from hcir_b200.engine import SearchSession, knn_predict, knn_topk, l2_normalize
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a B200; there is no CPU fallback"
    assert _lib.load().hcir_device_supported() == 1
    torch.cuda.set_device(0)


def _clean(bad):
    return not any(bad.values())


def _st():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------ K3 certification bound
@pytest.mark.parametrize("width", [1, 2, 3])   # 128 / 256 / 1024-thread variants of the kernel
def test_k3_never_certifies_when_exactly_kc_keys_are_staged(width):
    """when exactly kc keys exceed the staging hint but more candidates sit in the lists,
    the rows left out are bounded only by the kc-th staged score, not by the list thresholds.  A
    near-duplicate gallery (every score within eps of the k-th) with hand-made lists -- ALL rows listed,
    list thresholds = -inf, hint = the (kc+1)-th best bf16 score -- must come back UNCERTIFIED."""
    lib = _lib.load()
    n, d, k = 4096, 64, 5
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, d, generator=g)
    bank = (base + 1e-3 * torch.randn(n, d, generator=g)).cuda()
    q = (base + 1e-3 * torch.randn(1, d, generator=g)).cuda()
    g32, gbf, gdl = l2_normalize(bank)
    q32, qbf, qdl = l2_normalize(q)
    ld = g32.shape[1]
    score = (qbf.float() @ gbf.float().t())[0]                      # what the tensor-core contraction produces
    srt, order = torch.sort(score, descending=True)
    # a position >= 2k+64 where the sorted scores strictly drop: kc = that many keys exceed the hint
    kc = next(p for p in range(2 * k + 64, 600) if srt[p - 1] > srt[p])
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(1, n, ld, kc, 148, plan))
    plan.cap = n                                                     # one list holds the whole gallery
    nl = plan.nlists
    ws = torch.zeros(int(plan.keys_off) + nl * n * 8, dtype=torch.uint8, device="cuda")
    counts = ws[int(plan.counts_off): int(plan.counts_off) + nl * 4].view(torch.int32)
    counts[0] = n
    ws[int(plan.thr_out_off): int(plan.thr_out_off) + nl * 4].view(torch.float32).fill_(float("-inf"))
    ws[int(plan.thr_hi_off): int(plan.thr_hi_off) + 4].view(torch.float32).fill_(float(srt[kc]))
    raw = (score.view(torch.int32).to(torch.int64) << 32) | (0xFFFFFFFF - torch.arange(n, device="cuda"))
    ws[int(plan.keys_off): int(plan.keys_off) + n * 8].view(torch.int64).copy_(raw)
    plan.flags = width << 8
    o_s = torch.empty((1, k), device="cuda")
    o_i = torch.empty((1, k), dtype=torch.int64, device="cuda")
    ul = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    state = torch.zeros(4, dtype=torch.int32, device="cuda")
    _lib.check(lib.hcir_select_rescore(q32.data_ptr(), g32.data_ptr(), ld, 1, n, k, 0, plan, ws.data_ptr(),
                                       qdl.data_ptr(), float(gdl.max()), ld * 2.0 ** -22, o_s.data_ptr(),
                                       o_i.data_ptr(), ul.data_ptr(), state.data_ptr(), None, _st()), "K3")
    assert state.tolist() == [0, 1, 0, 0] and ul.tolist() == [0], state.tolist()   # uncertified, counters reset
    # sanity: the same lists with a far-away k-th score certify (well-separated gallery)
    sep = torch.randn(n, d, generator=g).cuda()
    sep[17] = q[0] * 3.0
    for j in range(1, k):
        sep[100 + j] = q[0] + 0.15 * j * torch.randn(d, generator=g).cuda()
    g32b, gbfb, gdlb = l2_normalize(sep)
    scoreb = (qbf.float() @ gbfb.float().t())[0]
    srtb, _ = torch.sort(scoreb, descending=True)
    ws[int(plan.thr_hi_off): int(plan.thr_hi_off) + 4].view(torch.float32).fill_(float(srtb[kc]))
    rawb = (scoreb.view(torch.int32).to(torch.int64) << 32) | (0xFFFFFFFF - torch.arange(n, device="cuda"))
    ws[int(plan.keys_off): int(plan.keys_off) + n * 8].view(torch.int64).copy_(rawb)
    _lib.check(lib.hcir_select_rescore(q32.data_ptr(), g32b.data_ptr(), ld, 1, n, k, 0, plan, ws.data_ptr(),
                                       qdl.data_ptr(), float(gdlb.max()), ld * 2.0 ** -22, o_s.data_ptr(),
                                       o_i.data_ptr(), ul.data_ptr(), state.data_ptr(), None, _st()), "K3")
    assert state.tolist() == [0, 0, 0, 0] and int(o_i[0, 0]) == 17


def test_k3_width_variants_are_bit_identical():
    bank, bl = synth.make_clustered(30000, 768, 27, 91)
    qs, _ = synth.make_clustered(300, 768, 27, 92)
    ref = None
    for width in (0, 1, 2, 3):
        gb = GalleryBank(bank, bl)
        gb.k3_width = width
        s, i = gb.topk(qs, 100, mode="tensor")
        if ref is None:
            ref = (s, i)
        assert torch.equal(i, ref[1]) and torch.equal(s, ref[0]), width


# ------------------------------------------------------------------------------------ fused tail (single GPU)
@pytest.mark.parametrize("T", [None, 0.07])
@pytest.mark.parametrize("nq", [64, 300, 2500])   # 1024-, 256- and 128-thread variants (k = 20)
def test_session_fused_vote_equals_oracle_vote(T, nq):
    """K3's tail votes inside the query's CTA: predictions == the oracle vote on the SAME neighbour
    lists (uniform: sklearn `_mode`; temperature: the extension), and == the separate K4 kernel."""
    bank, bl = synth.make_clustered(30000, 256, 27, 101)
    bl = bl * 2 + 2                                               # non-contiguous label values
    qs, _ = synth.make_clustered(nq, 256, 27, 102)
    gb = GalleryBank(bank, bl)
    sess = gb.session(nq, 20, T=T)
    assert sess.kernels_per_run == 5                               # K1, sample, thresholds, main, K3+tail: no fills
    for rep in range(2):                                           # the self-resetting counters survive a replay
        pred, sims, idx = sess.run(qs.cuda())
        nl = bl.numpy()[idx.cpu().numpy()]
        assert O.labels_agree_except_vote_ties(pred.cpu().numpy(), nl, gb.classes_, sims.cpu().numpy(), T) == 0
        assert torch.equal(pred, gb.vote_from_idx(sims, idx, T=T))
        assert sess.unc_state.tolist() == [0, 0, 0, 0]
    p_ref = gb.predict(qs, 20, T=T)
    assert torch.equal(pred.cpu(), p_ref)


# ------------------------------------------------------------------------------------ peer exchange on ONE GPU
class FakePeers:
    """G fake ranks on one device: G regions from hcir_peer_alloc (used in-process, no IPC open) and
    one completed-step counter per rank -- exactly what PeerExchange sets up across processes."""

    def __init__(self, G, slot_bytes):
        self.lib = _lib.load()
        self.G, self.slot_bytes = G, slot_bytes
        self.total = int(self.lib.hcir_peer_region_bytes(G, slot_bytes))
        self.ptrs = []
        for _ in range(G):
            p, h = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(self.lib.hcir_peer_alloc(self.total, C.byref(p), h), "peer_alloc")
            self.ptrs.append(int(p.value))
        self.steps = [torch.zeros((), dtype=torch.int64, device="cuda") for _ in range(G)]
        self.regions = (C.c_void_p * G)(*self.ptrs)

    def fill_tail(self, tail, rank, payload):
        tail.world, tail.rank, tail.payload, tail.slot_bytes = self.G, rank, payload, self.slot_bytes
        for r, p in enumerate(self.ptrs):
            tail.regions[r] = p
        tail.step = self.steps[rank].data_ptr()

    def region(self, r):
        from hcir_b200.peer import _DeviceBytes
        return torch.as_tensor(_DeviceBytes(self.ptrs[r], self.total), device="cuda")

    def header(self, r):
        return self.region(r)[:512].view(torch.int64).cpu().numpy()

    def slot(self, r, parity, rank):
        off = int(self.lib.hcir_peer_slot_offset(self.G, self.slot_bytes, parity, rank))
        return self.region(r)[off: off + self.slot_bytes]

    def close(self):
        torch.cuda.synchronize()
        for p in self.ptrs:
            self.lib.hcir_peer_free(p)


def test_peer_exchange_with_fake_ranks_on_one_gpu():
    """Verdict r1 item 3b: the peer kernels run on the driver's single-GPU box.  G = 3 shard banks on one
    device; every fake rank's K3 stores its packed rows into ALL regions and signals arrival; the fused
    wait + merge + vote kernel of every rank then gives the single-bank answer, bit for bit, and agrees
    with hcir_merge_topk_packed on the concatenated local blocks.  Three steps: both parities + reuse.
    In the last step the consumer of rank 0 is launched FIRST on a side stream, so it really spins on the
    arrival counters until the producers have run."""
    from hcir_b200.sharded import ShardPlan
    lib = _lib.load()
    G, n, d, k, nq, C_ = 3, 36000, 256, 20, 130, 9
    bank, bl = synth.make_clustered(n, d, C_, 111)
    full = GalleryBank(bank, bl, classes=np.arange(C_))
    sp = ShardPlan(n, G)
    shards = [GalleryBank(bank[sp.start(r):sp.stop(r)], bl[sp.start(r):sp.stop(r)], idx_offset=sp.start(r),
                          classes=np.arange(C_)) for r in range(G)]
    block = int(lib.hcir_packed_block_bytes(nq, k, 1))
    peers = FakePeers(G, block)
    sessions = [SearchSession(sh, nq, k, vote=False, pack=True,
                              tail_hook=lambda s_, t_, r=r: peers.fill_tail(t_, r, _lib.PAYLOAD_BLOCK))
                for r, sh in enumerate(shards)]
    # the sessions' eager warm-up pass already produced step 1 on every fake rank: consume it
    o = [[torch.empty((nq, k), dtype=dt, device="cuda") for dt in (torch.float32, torch.int64, torch.int32)]
         + [torch.empty((nq,), dtype=torch.int64, device="cuda")] for _ in range(G)]
    cls = full._classes_device()

    def consume(r, stream=None, T=0.0):
        _lib.check(lib.hcir_peer_merge_vote(peers.ptrs[r], G, nq, k, 1, block, peers.steps[r].data_ptr(),
                                            int(20e9), o[r][0].data_ptr(), o[r][1].data_ptr(), o[r][2].data_ptr(),
                                            C_, T, cls.data_ptr(), o[r][3].data_ptr(),
                                            stream if stream is not None else _st()), "peer_merge_vote")

    for r in range(G):
        consume(r)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for step in (2, 3, 4):
        qs, _ = synth.make_clustered(nq, d, C_, 120 + step)
        qd = qs.cuda()
        T = 0.07 if step == 3 else 0.0
        if step == 4:           # consumer first: it must wait for the arrivals
            torch.cuda.synchronize()
            consume(0, side.cuda_stream, T)
        for r in range(G):      # producers: graph replays (K1 .. K3 + tail)
            sessions[r].run(qd, check=False)
        for r in range(G):
            if not (step == 4 and r == 0):
                consume(r, None, T)
        torch.cuda.synchronize()
        s_ref, i_ref = full.topk(qd, k, return_device=True)
        p_ref = full.vote_from_idx(s_ref, i_ref, T=T if T > 0 else None)
        gathered = torch.cat([s.pack for s in sessions])
        m_s = torch.empty((nq, k), device="cuda")
        m_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        m_l = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        _lib.check(lib.hcir_merge_topk_packed(gathered.data_ptr(), G, nq, k, 1, 0, m_s.data_ptr(), m_i.data_ptr(),
                                              m_l.data_ptr(), _st()))
        for r in range(G):
            h = peers.header(r)
            assert h[:G].tolist() == [step] * G and h[48] == step and h[49] == 0 and h[50] == 0, (step, r, h[:52])
            assert int(peers.steps[r]) == step
            par = step & 1
            assert h[16 + par * 16: 16 + par * 16 + G].tolist() == [0] * G          # uncertified counts
            for src in range(G):                                                     # slots == the local blocks
                assert torch.equal(peers.slot(r, par, src), sessions[src].pack[:block]), (step, r, src)
            assert torch.equal(o[r][1], i_ref) and torch.equal(o[r][0], s_ref), (step, r)
            assert torch.equal(o[r][1], m_i) and torch.equal(o[r][0], m_s) and torch.equal(o[r][2], m_l)
            assert torch.equal(o[r][2], full.neighbour_labels(i_ref))
            assert torch.equal(o[r][3], p_ref), (step, r)
    peers.close()


def test_peer_prediction_payload_and_wait_kernel_with_fake_ranks():
    """Query-replica exchange on one GPU: every fake rank answers its own query slice, K3's tail stores
    the int64 predictions into all regions (payload 2), hcir_peer_wait completes the step; also the
    standalone hcir_peer_push producer."""
    lib = _lib.load()
    G, nq, k = 2, 200, 20
    bank, bl = synth.make_clustered(30000, 256, 27, 131)
    bl = bl * 3 + 1
    gb = GalleryBank(bank, bl)
    peers = FakePeers(G, nq * 8)
    sessions = [SearchSession(gb, nq, k, T=0.07, tail_hook=lambda s_, t_, r=r: peers.fill_tail(t_, r, _lib.PAYLOAD_PRED))
                for r in range(G)]

    def wait(r):
        _lib.check(lib.hcir_peer_wait(peers.ptrs[r], G, peers.steps[r].data_ptr(), int(20e9), _st()), "peer_wait")

    for r in range(G):
        wait(r)
    for step in (2, 3):
        qs = [synth.make_clustered(nq, 256, 27, 140 + 10 * step + r)[0].cuda() for r in range(G)]
        for r in range(G):
            sessions[r].run(qs[r], check=False)
        for r in range(G):
            wait(r)
        torch.cuda.synchronize()
        for r in range(G):
            for src in range(G):
                got = peers.slot(r, step & 1, src)[: nq * 8].view(torch.int64)
                assert torch.equal(got, gb.predict(qs[src], k, T=0.07).cuda()), (step, r, src)
            assert int(peers.steps[r]) == step and peers.header(r)[49] == 0
    # standalone push of a block that already sits in memory (PeerExchange.exchange)
    blocks = [torch.arange(nq, dtype=torch.int64, device="cuda") * (r + 1) for r in range(G)]
    metas = [torch.tensor([7 + r], dtype=torch.int32, device="cuda") for r in range(G)]
    for r in range(G):
        _lib.check(lib.hcir_peer_push(blocks[r].data_ptr(), nq * 8, peers.regions, G, r, nq * 8,
                                      peers.steps[r].data_ptr(), metas[r].data_ptr(), _st()), "peer_push")
    for r in range(G):
        wait(r)
    torch.cuda.synchronize()
    for r in range(G):
        h = peers.header(r)
        assert h[16 + 0 * 16: 16 + G].tolist() == [7, 8] and int(peers.steps[r]) == 4 and h[51] == 0
        for src in range(G):
            assert torch.equal(peers.slot(r, 0, src)[: nq * 8].view(torch.int64), blocks[src])
    peers.close()


# ------------------------------------------------------------------------------------ device-driven completion
def _dense_cluster_case(seed=11, nbase=8, per=300, n_other=20000, d=256, nq=32):
    g = torch.Generator().manual_seed(seed)
    bases = torch.randn(nbase, d, generator=g)
    dense = (bases[:, None, :] + 1e-3 * torch.randn(nbase, per, d, generator=g)).reshape(-1, d)
    n = n_other + nbase * per
    bank = torch.cat([torch.randn(n_other, d, generator=g), dense])[torch.randperm(n, generator=g)]
    qs = bases.repeat_interleave(nq // nbase, 0) + 1e-3 * torch.randn(nq, d, generator=g)
    return bank, qs


@pytest.mark.parametrize("vote", [False, True])
def test_device_driven_completion_runs_the_second_pass_inside_the_graph(vote):
    """300 near-identical neighbours per query: the first pass cannot certify the fp32 top-10; with
    device_completion the second tensor pass (gather -> gated main pass -> K3 on the compact batch, rows
    mapped back) is part of the graph, resolves every query, and the host finds NOTHING left to do."""
    bank, qs = _dense_cluster_case()
    labels = torch.arange(bank.shape[0]) % 7 * 3 + 1
    gb = GalleryBank(bank, labels)
    sess = SearchSession(gb, 32, 10, vote=vote, T=0.07 if vote else None, device_completion=True)
    assert sess.kernels_per_run == 8                      # 5 + setup, gated main pass, K3 on the batch
    for rep in range(2):
        pred, sims, idx = sess.run(qs.cuda())
        st = gb.last_stats
        assert st["completion"] == "device" and st["uncertified_first_pass"] > 0 and st["uncertified"] == 0, st
        s2, i2 = gb.topk(qs, 10, mode="exact")
        assert torch.equal(idx.cpu(), i2) and torch.equal(sims.cpu(), s2)
        if vote:
            assert torch.equal(pred.cpu(), gb.predict(qs, 10, T=0.07, mode="exact"))
        assert sess.counters.tolist()[4:] == [0, 0, 0, 0]
    # another batch through the same graph (few or no uncertified queries: the gated kernels mostly idle)
    easy, _ = synth.make_clustered(32, 256, 5, 7)
    _, sims, idx = sess.run(easy.cuda())
    assert gb.last_stats["uncertified"] == 0 and gb.last_stats["uncertified_first_pass"] < 8, gb.last_stats
    s2, i2 = gb.topk(easy, 10, mode="exact")
    assert torch.equal(idx.cpu(), i2) and torch.equal(sims.cpu(), s2)


def test_device_driven_completion_leaves_near_duplicates_and_overflow_to_the_exact_kernel():
    """A near-duplicate gallery defeats the second pass too, and 256 uncertified queries exceed the
    128-row completion batch: both kinds end on the final list and are finished by the exact kernel."""
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 256, generator=g)
    gb = GalleryBank(base + 1e-4 * torch.randn(8192, 256, generator=g), torch.arange(8192) % 5)
    qd = base + 1e-4 * torch.randn(256, 256, generator=g)
    sess = SearchSession(gb, 256, 10, device_completion=True)
    pred, sims, idx = sess.run(qd.cuda())
    st = gb.last_stats
    assert st["uncertified_first_pass"] == 256 and st["uncertified"] == 256 and gb.retry_stats["exact"] == 256, st
    s2, i2 = gb.topk(qd, 10, mode="exact")
    assert torch.equal(idx.cpu(), i2) and torch.equal(sims.cpu(), s2)
    assert torch.equal(pred.cpu(), gb.predict(qd, 10, mode="exact"))


def test_device_driven_completion_corrects_the_peer_rows_before_the_signal():
    """Fake ranks on one GPU, device completion on: the first K3 stores its rows but does not signal; the
    completion K3 overwrites the rows it certifies and signals -- the consumer merges CORRECT rows without
    any host-side repair or second exchange."""
    from hcir_b200.sharded import ShardPlan
    lib = _lib.load()
    G, k, nq = 2, 10, 32
    bank, qs = _dense_cluster_case(seed=21)
    n = bank.shape[0]
    bl = torch.arange(n) % 9
    full = GalleryBank(bank, bl, classes=np.arange(9))
    sp = ShardPlan(n, G)
    shards = [GalleryBank(bank[sp.start(r):sp.stop(r)], bl[sp.start(r):sp.stop(r)], idx_offset=sp.start(r),
                          classes=np.arange(9)) for r in range(G)]
    block = int(lib.hcir_packed_block_bytes(nq, k, 1))
    peers = FakePeers(G, block)
    sessions = [SearchSession(sh, nq, k, vote=False, pack=True, device_completion=True,
                              tail_hook=lambda s_, t_, r=r: peers.fill_tail(t_, r, _lib.PAYLOAD_BLOCK))
                for r, sh in enumerate(shards)]
    o_s = torch.empty((nq, k), device="cuda")
    o_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    o_l = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    pred = torch.empty((nq,), dtype=torch.int64, device="cuda")
    cls = full._classes_device()

    def consume(r):
        _lib.check(lib.hcir_peer_merge_vote(peers.ptrs[r], G, nq, k, 1, block, peers.steps[r].data_ptr(), int(20e9),
                                            o_s.data_ptr(), o_i.data_ptr(), o_l.data_ptr(), 9, 0.0, cls.data_ptr(),
                                            pred.data_ptr(), _st()), "peer_merge_vote")

    for r in range(G):
        consume(r)            # the sessions' warm-up pass produced step 1
    first = 0
    for r in range(G):
        sessions[r].run(qs.cuda(), check=False)
        cnt = sessions[r].counters.tolist()
        first += cnt[1]
        assert cnt[5] == 0, cnt          # everything resolved by the in-graph completion
    assert first > 0                     # ... and there WAS something to resolve
    s_ref, i_ref = full.topk(qs, k, mode="exact", return_device=True)
    for r in range(G):
        consume(r)
        torch.cuda.synchronize()
        assert torch.equal(o_i, i_ref) and torch.equal(o_s, s_ref), r
        assert torch.equal(pred, full.vote_from_idx(s_ref, i_ref))
        assert peers.header(r)[:G].tolist() == [2] * G and peers.header(r)[16:16 + G].tolist() == [0] * G
    peers.close()


def test_unrepresentative_sample_cannot_break_exactness():
    """The optimistic threshold assumes the strided sample is representative.  Here it is not: exactly
    thr_rank of the SAMPLE rows (one per chunk) are near-copies of the query and nothing else is, so the
    threshold lands on a planted row, fewer than k rows pass the filter, and the first pass cannot even fill
    the top-k.  Exactness must not depend on the sample: the query is uncertified, completed, and equal to
    the exact fp32 path -- through the eager path, the captured session and the in-graph completion."""
    lib = _lib.load()
    n, d, k, nq = 40000, 128, 20, 8
    g = torch.Generator().manual_seed(29)
    u = torch.nn.functional.normalize(torch.randn(1, d, generator=g), dim=1)
    bank = torch.randn(n, d, generator=g)
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(nq, n, lib.hcir_padded_dim(d), 2 * k + 64, 148, plan))
    assert plan.sample_rows > 0 and plan.thr_rank < k          # the regime the test is about
    for c in range(plan.thr_rank):                              # one planted row in each of thr_rank chunks
        row = (c * plan.chunk_w + 3) * plan.sample_stride       # sample row i lives at gallery row i * stride
        bank[row] = u[0] * 5.0 + 0.05 * torch.randn(d, generator=g)
    qs = u + 1e-3 * torch.randn(nq, d, generator=g)
    gb = GalleryBank(bank)
    s_ex, i_ex = gb.topk(qs, k, mode="exact")
    s1, i1 = gb.topk(qs, k, mode="tensor")
    assert gb.last_stats["uncertified"] == nq, gb.last_stats   # nobody could be certified from the short lists
    assert torch.equal(i1, i_ex) and torch.equal(s1, s_ex)
    for dc in (False, True):
        sess = SearchSession(gb, nq, k, vote=False, device_completion=dc)
        _, s2, i2 = sess.run(qs.cuda())
        assert gb.last_stats["uncertified_first_pass"] == nq
        assert torch.equal(i2.cpu(), i_ex) and torch.equal(s2.cpu(), s_ex), dc


def test_fuzz_tensor_path_equals_exact_path_bit_for_bit():
    """Seeded random shapes and data shapes (clustered, heavy-tailed cluster sizes, duplicated rows, a few
    zero rows): the tensor path (eager and captured, every K3 width the shape selects, host- and
    device-driven completion) must return exactly what the fp32 CUDA-core path returns."""
    rng = np.random.default_rng(2024)
    for case in range(18):
        d = int(rng.choice([64, 96, 256, 512, 768, 1024, 2048]))
        n = int(rng.integers(9000, 120000 if d <= 768 else 40000))
        nq = int(rng.choice([1, 7, 64, 129, 300, 700, 2500]))
        k = int(rng.choice([1, 5, 20, 50, 100, 160, 300]))
        if n < 8 * (2 * k + 64) or n < 4096:
            continue
        ncls = int(rng.integers(2, 40))
        g = torch.Generator().manual_seed(1000 + case)
        bank, _ = synth.make_clustered(n, d, ncls, 3000 + case)
        if case % 3 == 0:      # a block of exact duplicates and two zero rows
            bank[100:100 + min(k + 7, 200)] = bank[99]
            bank[7] = 0.0
            bank[n - 1] = 0.0
        if case % 4 == 1:      # a tight cluster: many near-ties around the k-th neighbour
            bank[2000:2000 + 3 * k + 50] = bank[1999] + 2e-3 * torch.randn(3 * k + 50, d, generator=g)
        qs, _ = synth.make_clustered(nq, d, ncls, 4000 + case)
        if case % 4 == 1:
            qs[0] = bank[1999] * 0.7
        gb = GalleryBank(bank)
        s_ex, i_ex = gb.topk(qs, k, mode="exact")
        s1, i1 = gb.topk(qs, k, mode="tensor")
        assert torch.equal(i1, i_ex) and torch.equal(s1, s_ex), (case, n, d, nq, k, gb.last_stats)
        sess = SearchSession(gb, nq, k, vote=False, device_completion=bool(case % 2))
        _, s2, i2 = sess.run(qs.cuda())
        assert torch.equal(i2.cpu(), i_ex) and torch.equal(s2.cpu(), s_ex), (case, n, d, nq, k, gb.last_stats)
        del sess, gb


# ------------------------------------------------------------------------------------ full-size oracle checks
def _oracle_subsample_check(gb, qs_dev, sims, idx, k, rows, chunk=64):
    """fp32 ``torch.mm`` + ``topk`` (qualitative_test.py:79-84) on the HOST copy of the unit bank for a
    query subsample; tie policy of BASELINE.md section 4."""
    bank_cpu = gb.g32[:, : gb.d].cpu()
    qn = O.normalize(qs_dev[rows].cpu())
    ov, oi = O.mm_topk_chunked(qn, bank_cpu, min(k + 8, gb.n), chunk=chunk)
    bad = O.check_topk_against_topk(idx[rows].cpu().numpy(), sims[rows].cpu().numpy(), oi.numpy(), ov.numpy(), atol=5e-7)
    assert _clean(bad), bad
    agree = (idx[rows].cpu() == oi[:, :k]).float().mean().item()
    assert agree > 0.995, agree          # identical except near-tie swaps


def test_c2_full_size_against_host_oracle():
    """BASELINE configs[1] at full size: 200k x 768, 10k queries, k=20; 256 queries against the oracle,
    all 10k predictions against the oracle vote on the returned neighbour lists (uniform and T=0.07)."""
    n, d, q, k = 200_000, 768, 10_000, 20
    bank, bl = synth.make_clustered(n, d, 61, 1236, device="cuda")
    qs, _ = synth.make_clustered(q, d, 61, 4323, device="cuda")
    gb = GalleryBank(bank, bl, classes=np.arange(61))
    del bank
    for T in (None, 0.07):
        sess = gb.session(q, k, T=T)
        pred, sims, idx = sess.run(qs)
        assert gb.last_stats["uncertified"] == 0
        nl = bl[idx].cpu().numpy()
        assert O.labels_agree_except_vote_ties(pred.cpu().numpy(), nl, np.arange(61), sims.cpu().numpy(), T) == 0
    _oracle_subsample_check(gb, qs, sims, idx, k, torch.arange(0, q, 39, device="cuda")[:256])


def test_c3_full_size_against_host_oracle():
    """BASELINE configs[2] at full size (the bench headline): 1M x 768, 4096 queries, top-100."""
    n, d, q, k = 1_000_000, 768, 4096, 100
    bank, _ = synth.make_clustered(n, d, 61, 1237, device="cuda")
    qs, _ = synth.make_clustered(q, d, 61, 4324, device="cuda")
    gb = GalleryBank(bank)
    del bank
    sess = gb.session(q, k, vote=False)
    _, sims, idx = sess.run(qs)
    assert gb.last_stats["uncertified"] == 0
    _oracle_subsample_check(gb, qs, sims, idx, k, torch.arange(0, q, 16, device="cuda")[:256])


def test_c4_streaming_against_host_oracle():
    """BASELINE configs[3] regime at >= 4M rows: 64 queries, k=20, every query against the oracle."""
    n, d, q, k = 4_000_000, 768, 64, 20
    bank, _ = synth.make_clustered(n, d, 61, 1238, device="cuda")
    gb = GalleryBank(bank)
    del bank
    qs, _ = synth.make_clustered(q, d, 61, 4390, device="cuda")
    sess = gb.session(q, k, vote=False)
    _, sims, idx = sess.run(qs)
    _oracle_subsample_check(gb, qs, sims, idx, k, torch.arange(q, device="cuda"), chunk=16)


def test_c5_shard_takes_the_second_pass_and_matches_the_oracle():
    """BASELINE configs[4], one rank's shard of the 8-GPU layout: 1.25M x 2048, k=200.  With 16384 queries
    a few per step cannot be certified from kc = 464 candidates and take the second tensor pass; the
    test runs 2048 queries (the uncertified ones included in the oracle subsample) against the oracle."""
    n, d, q, k = 1_250_000, 2048, 2048, 200
    bank, _ = synth.make_clustered(n, d, 61, 1239, device="cuda")
    gb = GalleryBank(bank)
    del bank
    torch.cuda.empty_cache()
    qs, _ = synth.make_clustered(q, d, 61, 4329, device="cuda")
    gb.kernel_events = []
    sims, idx = gb.topk(qs, k, mode="tensor", return_device=True)
    names = [e[0] for e in gb.kernel_events]
    gb.kernel_events = None
    n_unc = gb.last_stats["uncertified"]
    rows = torch.arange(0, q, 16, device="cuda")[:128]
    if n_unc > 0:                       # completed by the second pass, not the brute force
        assert "simtopk_retry" in names and gb.retry_stats["second_pass"] > 0, (names, gb.retry_stats)
    _oracle_subsample_check(gb, qs, sims, idx, k, rows, chunk=32)
    # the exact fp32 path agrees bit for bit on a subsample
    s2, i2 = gb.topk(qs[rows[:32]], k, mode="exact", return_device=True)
    assert torch.equal(i2, idx[rows[:32]]) and torch.equal(s2, sims[rows[:32]])


# ------------------------------------------------------------------------------------ wrappers without a test (r1)
def test_functional_wrappers_against_the_reference_formulations():
    """knn_topk == torch.mm + topk (qualitative_test.py:79-84), compute_similarity_topk ==
    compute_similarity(...).topk (dual_view_model.py:317-335), knn_predict == sklearn's
    KNeighborsClassifier(metric="cosine").fit/predict (classification_engine.py:80-82)."""
    bank, bl = synth.make_clustered(20000, 512, 27, 151)
    qs, _ = synth.make_clustered(400, 512, 27, 152)
    bn, qn = O.normalize(bank), O.normalize(qs)
    ov, oi = O.mm_topk(qn, bn, 28)
    for got in (knn_topk(bank, qs, 20), knn_topk(bn, qn, 20, normalized=True),
                hcir_b200.compute_similarity_topk(qs.numpy(), bank.numpy(), 20)):
        s, i = got
        assert _clean(O.check_topk_against_topk(np.asarray(i), np.asarray(s), oi.numpy(), ov.numpy(), atol=5e-7))
    sk_pred, _, _ = O.sklearn_knn(bn.numpy(), bl.numpy(), qn.numpy(), 20)
    ours = knn_predict(qs, bank, bl, 20)
    assert (np.asarray(ours) == sk_pred).mean() >= 0.995            # sklearn's own near-tie order only
    s, i = knn_topk(bank, qs, 20)
    nl = bl.numpy()[np.asarray(i)]
    assert O.labels_agree_except_vote_ties(np.asarray(ours), nl, np.arange(27)) == 0
    pt = knn_predict(qs, bank, bl, 20, T=0.07)
    assert O.labels_agree_except_vote_ties(np.asarray(pt), nl, np.arange(27), np.asarray(s), 0.07) == 0


def test_classifier_temperature_weights_and_leave_one_out_neighbours():
    bank, bl = synth.make_clustered(12000, 256, 13, 161)
    bl = bl * 5 + 3
    qs, _ = synth.make_clustered(333, 256, 13, 162)
    clf = KNeighborsClassifierB200(n_neighbors=20, weights="temperature", T=0.07).fit(bank, bl.numpy())
    pred = clf.predict(qs.numpy())
    dist, ind = clf.kneighbors(qs.numpy())
    nl = bl.numpy()[ind]
    assert O.labels_agree_except_vote_ties(pred, nl, clf.classes_, 1.0 - dist, 0.07) == 0
    ref, _ = O.vote_temperature(1.0 - dist, nl, clf.classes_, 0.07)
    assert (pred == ref).mean() > 0.99
    # kneighbors(X=None): neighbours of every training row, the row itself excluded == sklearn
    small, sl = synth.make_clustered(3000, 64, 5, 163)
    small[100:104] = small[99]                                      # duplicates of row 99
    ours = KNeighborsClassifierB200(n_neighbors=5).fit(small, sl.numpy())
    d_o, i_o = ours.kneighbors()
    from sklearn.neighbors import KNeighborsClassifier
    sk = KNeighborsClassifier(n_neighbors=5, metric="cosine").fit(O.normalize(small).numpy(), sl.numpy())
    d_s, i_s = sk.kneighbors()
    assert i_o.shape == i_s.shape == (3000, 5) and (i_o != np.arange(3000)[:, None]).all()
    assert (i_o == i_s).mean() > 0.995 and np.abs(d_o - d_s).max() < 2e-6
    for r in (99, 100, 103):                                        # the other duplicates come first, at distance 0
        assert set(i_o[r][:4]) == set(range(99, 104)) - {r} and d_o[r][:4].max() < 1e-6


def test_flat_index_matches_faiss_convention_of_the_oracle():
    """FlatIndex.search == faiss.normalize_L2 + IndexFlatL2.search (inference.py:75,90-108): D = squared
    L2 of unit rows = 2 - 2 cos ascending, I = the oracle's mm+topk indices."""
    bank, _ = synth.make_clustered(9000, 384, 9, 171)
    qs, _ = synth.make_clustered(50, 384, 9, 172)
    bn, qn = O.normalize(bank), O.normalize(qs)
    index = hcir_b200.FlatIndex(384)
    index.add(bank.numpy()[:4000])          # un-normalised in, like faiss.normalize_L2 would fix up
    index.add(bank.numpy()[4000:])
    D, I = index.search(qs.numpy(), 10)
    ov, oi = O.mm_topk(qn, bn, 18)
    assert _clean(O.check_topk_against_topk(I, 1.0 - D / 2.0, oi.numpy(), ov.numpy(), atol=1e-6))
    np.testing.assert_allclose(D, 2.0 - 2.0 * np.take_along_axis(O.similarity_matrix(qn, bn).numpy(), I, 1), atol=2e-6)
    assert (np.diff(D, axis=1) >= -1e-7).all() and I.dtype == np.int64 and D.dtype == np.float32


def test_cached_bank_is_rebuilt_after_an_in_place_edit():
    """retrieve_similar_images caches the normalised bank per CONTENT: an in-place edit anywhere in the
    embeddings (not only at sampled positions) must change the answer like the reference's does."""
    hcir_b200.retrieval.clear_bank_cache()
    emb = synth.make_clustered(5000, 128, 7, 181)[0].numpy()
    paths = [f"p{i}" for i in range(5000)]
    q = emb[1234] * 0.5
    first = hcir_b200.retrieve_similar_images(q, emb, paths, top_k=3)
    assert first[0]["path"] == "p1234"
    again = hcir_b200.retrieve_similar_images(q, emb, paths, top_k=3)
    assert [x["path"] for x in again] == [x["path"] for x in first] and len(hcir_b200.retrieval._BANK_CACHE) == 1
    emb[1234] = -emb[1234]                  # one row, off every sampling grid
    emb[4321] = q * 4.0
    after = hcir_b200.retrieve_similar_images(q, emb, paths, top_k=3)
    ref = O.retrieve_similar_images(q, emb, paths, top_k=3)
    assert after[0]["path"] == "p4321" == ref[0]["path"] and "p1234" not in [x["path"] for x in after]


def test_kth_neighbour_euclidean_branch_and_metrics_on_device():
    """neg_sampling.py:38-41 (metric='euclidean': descending sort of -cdist on the RAW rows) and
    quantitative_eval.py:195-209 (Recall@K / AP@K) on device-resident indices."""
    from hcir_b200 import metrics
    g = torch.Generator().manual_seed(18)
    emb = torch.randn(256, 512, generator=g) * (1.0 + torch.rand(256, 1, generator=g))   # unequal norms
    dist = torch.cdist(emb, emb)                                   # the reference's call (fp32, mm-based:
    _, order = torch.sort(-dist, dim=1, descending=True)           #   abs error up to ~0.05 at these norms)
    d64 = torch.cdist(emb.double(), emb.double(), compute_mode="donot_use_mm_for_euclid_dist")
    _, o64 = torch.sort(d64, dim=1)
    for k in (1, 7, 64):
        ours = metrics.kth_neighbour(emb, k, metric="euclidean")
        # exact ranking (fp64 distances) except true near-ties
        for r in torch.nonzero(ours != o64[:, k - 1]).flatten().tolist():
            assert abs(float(d64[r, ours[r]] - d64[r, o64[r, k - 1]])) < 1e-5 * float(d64[r, ours[r]])
        assert (ours == o64[:, k - 1]).float().mean() > 0.99
        # the reference's fp32 cdist ranking agrees wherever its own rounding error cannot flip the order
        ref = order[:, k - 1]
        for r in torch.nonzero(ours != ref).flatten().tolist():
            assert abs(float(d64[r, ours[r]] - d64[r, ref[r]])) < 0.06
        assert (ours == ref).float().mean() > 0.97
    assert torch.equal(metrics.kth_neighbour(emb.cuda(), 1, metric="euclidean").cpu(), torch.arange(256))
    with pytest.raises(ValueError):
        metrics.kth_neighbour(emb, 3, metric="manhattan")
    # Recall@K / AP@K with the neighbour lists left on the device == the reference's python loop
    bank, _ = synth.make_clustered(8000, 128, 11, 191)
    qs, _ = synth.make_clustered(40, 128, 11, 192)
    _, idx = GalleryBank(bank).topk(qs, 50, return_device=True)
    gt = [list(map(int, idx[r, [3, 17, 45]].tolist())) if r % 3 else [7999 - r] for r in range(40)]
    gt[5] = []
    got = metrics.recall_ap_at_k(idx, gt, ks=(10, 20, 50))
    paths = [str(i) for i in range(8000)]
    rec = {k: 0 for k in (10, 20, 50)}
    aps = {k: [] for k in (10, 20, 50)}
    for r in range(40):                                             # quantitative_eval.py:195-209
        retrieved = [paths[i] for i in idx[r].tolist()]
        gt_list = [paths[i] for i in gt[r]]
        for k in (10, 20, 50):
            top = retrieved[:k]
            if any(g_ in top for g_ in gt_list):
                rec[k] += 1
            hits, sp_ = 0, 0.0
            for i, p in enumerate(top):
                if p in gt_list:
                    hits += 1
                    sp_ += hits / (i + 1)
            aps[k].append(sp_ / min(len(gt_list), k) if gt_list else 0.0)
    for k in (10, 20, 50):
        assert abs(got["Recall"][k] - rec[k] / 40) < 1e-12 and abs(got["mAP"][k] - float(np.mean(aps[k]))) < 1e-12


def test_host_pipeline_overlaps_copies_and_returns_the_synchronous_answers():
    """HostPipeline: host batches in, host answers out, copies on their own streams.  Every batch must equal
    the synchronous call on the same batch -- including a batch that the first pass cannot certify (the
    unrepresentative-sample construction), which result() redoes -- for top-k and for both votes."""
    from hcir_b200 import HostPipeline
    lib = _lib.load()
    n, d, k, nq, ncls = 40000, 128, 20, 8, 7
    g = torch.Generator().manual_seed(31)
    u = torch.nn.functional.normalize(torch.randn(1, d, generator=g), dim=1)
    bank = torch.randn(n, d, generator=g)
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(nq, n, lib.hcir_padded_dim(d), 2 * k + 64, 148, plan))
    for c in range(plan.thr_rank):
        bank[(c * plan.chunk_w + 3) * plan.sample_stride] = u[0] * 5.0 + 0.05 * torch.randn(d, generator=g)
    labels = torch.randint(0, ncls, (n,), generator=g)
    gb = GalleryBank(bank, labels.numpy(), classes=np.arange(ncls))
    hard = (u + 1e-3 * torch.randn(nq, d, generator=g)).pin_memory()
    batches = [torch.randn(nq, d, generator=g).pin_memory() for _ in range(5)]
    batches.insert(2, hard)
    want_topk = [gb.topk(b, k) for b in batches]
    for depth in (1, 2, 3):
        pipe = HostPipeline.for_bank(gb, nq, k, want="topk", depth=depth)
        pend = [pipe.submit(b) for b in batches]          # more than `depth`: the oldest completes on its own
        got = [p.result() for p in pend]
        assert pend[2].redone and not pend[0].redone
        for (s, i), (ws, wi) in zip(got, want_topk):
            assert isinstance(s, torch.Tensor) and not s.is_cuda and s.is_pinned()
            assert torch.equal(i, wi) and torch.equal(s, ws)
        assert pend[1].result() is got[1]                 # idempotent
    for T in (None, 0.07):
        want_pred = [gb.predict(b, k, T=T) for b in batches]
        pipe = HostPipeline.for_bank(gb, nq, k, want="pred", T=T, depth=2)
        pend = []
        for b in batches:
            pend.append(pipe.submit(b.numpy()))           # numpy in -> numpy out
        pipe.drain()
        for p, w in zip(pend, want_pred):
            r = p.result()
            assert isinstance(r, np.ndarray) and r.dtype == np.int64
            assert np.array_equal(r, np.asarray(w))
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(nq + 1, d))
    with pytest.raises(ValueError):
        HostPipeline.for_bank(GalleryBank(torch.randn(300, 16)), 4, 5)   # exact-path shape: nothing to pipeline
