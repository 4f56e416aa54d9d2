"""Multi-GPU: one process per GPU of one box (``torch.distributed`` for rendezvous; no
reference analogue -- the reference's kNN is single-process CPU, classification_engine.py:51,63).

* ``ShardedGallery``: the gallery is row-sharded.  Each rank computes its exact local top-k; the tail of
  K3 stores every query's k candidates straight into every rank's peer region over NVLink, and ONE more
  kernel per rank waits for all blocks, merges them (K5) and votes -- or, with ``exchange="nccl"``, one
  all-gather of the packed blocks + merge + vote.
* ``QueryShardedGallery``: every rank holds the whole gallery and answers a slice of the query batch; the
  predictions / top-k rows are exchanged the same way.

``ShardPlan``, ``exchange_candidates`` and ``gather_rows`` are backend-agnostic (tested with gloo on CPU);
the local search, the peer exchange and the merge are CUDA-only."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

import os

from . import _lib
from .engine import GalleryBank, PendingStep, _as_2d_f32, _stream_ptr, _to_host, _to_host_many
import warnings

from . import peer as _peer
from .peer import PeerExchange

# how the per-rank result blocks travel: "peer" = the library's own stores over NVLink peer memory
# (K3's tail + csrc/peer.cu, captured into the step's CUDA graph), "nccl" = one ncclAllGather
DEFAULT_EXCHANGE = os.environ.get("HCIR_EXCHANGE", "peer")


def _resolve_exchange(exchange, group, device) -> str:
    """"peer" needs CUDA IPC + P2P between all GPUs of the group: a collective self-test decides,
    identically on every rank, whether the channel works; NCCL carries the blocks otherwise."""
    exchange = exchange or DEFAULT_EXCHANGE
    if exchange not in ("peer", "nccl"):
        raise ValueError(f"unknown exchange {exchange!r}")
    if exchange == "peer" and dist.get_world_size(group) > 1 and not _peer.probe(group, device):
        warnings.warn("hcir_b200: peer-memory exchange unavailable on this box (CUDA IPC / P2P); using NCCL")
        return "nccl"
    return exchange


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row partition: rank r owns rows [start(r), stop(r)); the first ``n % world``
    ranks hold one extra row."""
    n: int
    world: int

    def start(self, r: int) -> int:
        base, extra = divmod(self.n, self.world)
        return r * base + min(r, extra)

    def stop(self, r: int) -> int:
        return self.start(r + 1) if r + 1 < self.world else self.n

    def size(self, r: int) -> int:
        return self.stop(r) - self.start(r)

    def owner(self, row: int) -> int:
        base, extra = divmod(self.n, self.world)
        pivot = extra * (base + 1)
        return row // (base + 1) if row < pivot else extra + (row - pivot) // max(base, 1)


def exchange_candidates(sims: torch.Tensor, idx: torch.Tensor, labels: torch.Tensor | None = None,
                        group=None):
    """The path's single collective: ONE all-gather of every rank's [Q, k] exact local top-k
    (fp32 sims, int64 global indices, optional int32 labels, packed into one byte buffer) ->
    [G, Q, k] arrays on every rank."""
    world = dist.get_world_size(group)
    q, k = sims.shape
    parts = [sims.contiguous().view(torch.uint8).view(q, k * 4), idx.contiguous().view(torch.uint8).view(q, k * 8)]
    if labels is not None:
        parts.append(labels.contiguous().view(torch.uint8).view(q, k * 4))
    packed = torch.cat(parts, dim=1).contiguous()                      # [Q, k*(12|16)] bytes
    # concatenated-along-dim-0 output: the one layout both NCCL and gloo accept
    out = torch.empty((world * q, packed.shape[1]), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, q, packed.shape[1])
    g_sims = out[:, :, : k * 4].contiguous().view(torch.float32).view(world, q, k)
    g_idx = out[:, :, k * 4: k * 12].contiguous().view(torch.int64).view(world, q, k)
    g_lab = out[:, :, k * 12:].contiguous().view(torch.int32).view(world, q, k) if labels is not None else None
    return g_sims, g_idx, g_lab


def merge_topk(g_sims: torch.Tensor, g_idx: torch.Tensor, g_lab: torch.Tensor | None, k: int):
    """K5 on device: canonical top-k of the union of G canonical lists."""
    lib = _lib.load()
    G, nq, kk = g_sims.shape
    dev = g_sims.device
    out_sim = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_lab = torch.empty((nq, k), dtype=torch.int32, device=dev) if g_lab is not None else None
    if kk != k:
        raise ValueError("merge_topk expects per-shard lists of width k")
    with torch.cuda.device(dev):
        _lib.check(lib.hcir_merge_topk(g_sims.data_ptr(), g_idx.data_ptr(),
                                       g_lab.data_ptr() if g_lab is not None else None, G, nq, k,
                                       out_sim.data_ptr(), out_idx.data_ptr(),
                                       out_lab.data_ptr() if out_lab is not None else None, _stream_ptr()),
                   "merge_topk")
    return out_sim, out_idx, out_lab


def _hand_over(res, lane_stream, caller_stream):
    """Results allocated on a lane stream are consumed on the caller's stream: tell the caching allocator,
    so that a block the caller drops is not handed to the lane's next step while the caller's kernels
    still read it."""
    if lane_stream is None:
        return
    for t in ((res,) if isinstance(res, torch.Tensor) else res):
        t.record_stream(caller_stream)


class _Lanes:
    """Pipelined submissions (``submit_*``) rotate over ``lanes`` CUDA streams, each with its own captured
    session and peer channel, so the latency-bound tail of step i (K3, wait + merge + vote) runs beside
    the sample pass / main pass of step i+1 instead of in front of it.  ``lanes = None``: 3 lanes when a
    rank's step is short -- its query rows x its gallery rows <= ``LANE_MAX_SCORES`` -- and 1 otherwise.
    Measured, two steps in flight (profiles/README.md): 8 GPUs, 10M x 768 gallery in shards, 64 queries:
    0.361 -> 0.322 ms (2 GPUs, same shard shape: 0.361 -> 0.326 with 2 lanes -> 0.314 with 3 -> 0.303 with 4
    lanes and 4 in flight); 8 GPUs, 1M x 768 in shards, 4096 queries: 1.217 -> 1.138 ms; 8 GPUs, 200k x 768
    replicas, 1250 queries per rank: 0.412 -> 0.377 ms.  NOT where the tensor pass dominates the step: 1M x
    768 replicas at 512 queries per rank gained on 2 GPUs (0.819 -> 0.756 ms) but lost on 8 (0.753 -> 0.787:
    a rank whose next main pass starts first delays its own K3, and seven peers wait for it), and one GPU
    at full batch shows nothing (the step is paced by the power-capped tensor pass).  HCIR_LANES overrides."""

    LANE_MAX_SCORES = 3.0e8

    def _init_lanes(self):
        self.lanes = int(os.environ["HCIR_LANES"]) if os.environ.get("HCIR_LANES") else None
        self._lane_streams = {}
        self._submitted = 0

    def _lane(self, nq_per_rank: int, n_local: int):
        """-> (lane index, stream | None) for the next pipelined submission (the same on every rank)."""
        lanes = self.lanes if self.lanes else (3 if nq_per_rank * n_local <= self.LANE_MAX_SCORES else 1)
        lane = self._submitted % max(1, lanes)
        self._submitted += 1
        if lane == 0:
            return 0, None          # lane 0 is the caller's stream
        st = self._lane_streams.get(lane)
        if st is None:
            st = self._lane_streams[lane] = torch.cuda.Stream(device=self.device)
        return lane, st


class ShardedGallery(_Lanes):
    """Rank-local shard of a global gallery + the exchange/merge step.

    ``features_local`` are THIS rank's rows [plan.start(rank), plan.stop(rank)); queries are
    replicated on every rank.  Results are identical on every rank and bit-identical to the
    single-GPU result (local lists are exact fp32 canonical lists, the merge is a pure
    sort-merge on (sim desc, idx asc))."""

    LANE_MAX_SCORES = 6.0e8   # every rank re-scores ALL queries: the tail stays long next to the shard's main pass

    def __init__(self, features_local, labels_local=None, *, n_total: int, group=None, device=None,
                 classes=None, exchange: str | None = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.plan = ShardPlan(int(n_total), self.world)
        if n_total >= (1 << 32) - 1:
            raise ValueError("global gallery must hold fewer than 2^32 - 1 rows")
        feats, _ = _as_2d_f32(features_local, "features_local")
        if feats.shape[0] != self.plan.size(self.rank):
            raise ValueError(f"rank {self.rank}: got {feats.shape[0]} rows, plan says {self.plan.size(self.rank)}")
        self.bank = GalleryBank(feats, labels_local, device=device, idx_offset=self.plan.start(self.rank),
                                classes=classes)
        self.device = self.bank.device
        self.exchange = _resolve_exchange(exchange, group, self.device)
        self.profile = False       # bench.py: capture per-kernel events inside the graph
        self.last_session = None
        self._init_lanes()

    def _local(self, q: torch.Tensor, k: int, mode: str):
        """Exact local top-min(k, n_local), padded to width k with (-inf, -1)."""
        kl = min(k, self.bank.n)
        sess = self.bank.session(q.shape[0], kl, vote=False, profile=self.profile) if mode == "auto" else None
        if sess is not None:   # the whole local step (K1..K3) is one CUDA-graph launch
            _, sims, idx = sess.run(q)
            self.last_session = sess
        else:
            sims, idx = self.bank._topk_device(q, kl, mode)
        if kl < k:
            pad_s = torch.full((q.shape[0], k - kl), float("-inf"), dtype=torch.float32, device=self.device)
            pad_i = torch.full((q.shape[0], k - kl), -1, dtype=torch.int64, device=self.device)
            sims, idx = torch.cat([sims, pad_s], 1), torch.cat([idx, pad_i], 1)
        return sims, idx

    def _tail_peer(self, sess, tail):
        """K3's tail as the producer of the exchange: every query's CTA stores its packed rows
        (idx | sims | labels) into every rank's peer region over NVLink; the launch's last CTA sends
        the uncertified count and the arrival signal.  The channel is allocated collectively on the
        session's eager warm-up pass."""
        xc = getattr(sess, "xchg", None)
        if xc is None:
            xc = sess.xchg = PeerExchange(sess.block_bytes, group=self.group, device=self.device)
        xc.fill_tail(tail, _lib.PAYLOAD_BLOCK, sess.block_bytes)

    def _post_peer(self, sess, want_vote: bool, T):
        """What follows K3 in the step's graph: ONE kernel that waits for all ranks' blocks of the
        step, merges them in place (K5) and votes.  No collective-library call on the data path."""
        lib, dev = self.bank.lib, self.device
        nq, k = sess.nq, sess.k
        has_lab = sess.out_lab is not None
        xc = sess.xchg
        o_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        o_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        o_l = torch.empty((nq, k), dtype=torch.int32, device=dev) if has_lab else None
        pred = torch.empty((nq,), dtype=torch.int64, device=dev) if (want_vote and has_lab) else None
        self.bank.launches += 1
        _lib.check(lib.hcir_peer_merge_vote(
            xc.local_ptr, self.world, nq, k, int(has_lab), xc.slot_bytes, xc.step.data_ptr(), xc.timeout_ns,
            o_s.data_ptr(), o_i.data_ptr(), o_l.data_ptr() if has_lab else None,
            len(self.bank.classes_) if pred is not None else 0, float(T) if T is not None else 0.0,
            self.bank._classes_device().data_ptr() if pred is not None else None,
            pred.data_ptr() if pred is not None else None, _stream_ptr()), "peer_merge_vote")
        if not torch.cuda.is_current_stream_capturing():
            xc.host_step += 1   # the eager warm-up pass completed a step
        return {"gathered": None, "xchg": xc, "sims": o_s, "idx": o_i, "lab": o_l, "pred": pred}

    def _post_nccl(self, sess, want_vote: bool, T):
        """Tail of the step, captured into the same CUDA graph as the local search: ONE all-gather
        of every rank's packed block (+ trailer = its uncertified count), K5 reading the gathered
        blocks in place, and (predict) the vote."""
        lib, dev = self.bank.lib, self.device
        nq, k = sess.nq, sess.k
        has_lab = sess.out_lab is not None
        stride = sess.pack.numel()
        gathered = torch.empty((self.world * stride,), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, sess.pack, group=self.group)
        o_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        o_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        o_l = torch.empty((nq, k), dtype=torch.int32, device=dev) if has_lab else None
        self.bank.launches += 1
        _lib.check(lib.hcir_merge_topk_packed(gathered.data_ptr(), self.world, nq, k, int(has_lab), stride,
                                              o_s.data_ptr(), o_i.data_ptr(), o_l.data_ptr() if has_lab else None,
                                              _stream_ptr()), "merge_topk_packed")
        pred = None
        if want_vote:
            pred = self.bank.vote_from_labels(o_s, o_l, T=T)
        return {"gathered": gathered, "sims": o_s, "idx": o_i, "lab": o_l, "pred": pred}

    def _packed_session(self, nq: int, k: int, want_vote: bool, T, lane: int = 0):
        key = ("gallery-sharded", self.world, self.exchange, want_vote, None if T is None else float(T), lane)
        if self.exchange == "peer":
            return self.bank.session(nq, k, vote=False, profile=self.profile, pack=True, tail_hook=self._tail_peer,
                                     post=lambda s_: self._post_peer(s_, want_vote, T), post_key=key)
        return self.bank.session(nq, k, vote=False, profile=self.profile, pack=True, trailer=True,
                                 post=lambda s_: self._post_nccl(s_, want_vote, T), post_key=key)

    def close(self):
        """COLLECTIVE: drop the cached sessions and close their peer channels (IPC mappings, regions)."""
        self.bank.drop_sessions()

    def _check_k(self, k: int):
        if not (1 <= int(k) <= self.plan.n):
            raise ValueError(f"k={k} must be in [1, N={self.plan.n}]")

    def _step_packed(self, q: torch.Tensor, k: int, want_vote: bool, T):
        """Whole multi-GPU step as ONE graph launch per rank; None if the tensor path does not apply."""
        sess = self._packed_session(q.shape[0], k, want_vote, T)
        if sess is None or (want_vote and sess.out_lab is None):
            return None
        sess.run(q, check=False)
        self.last_session = sess
        out = sess.post_out
        # every rank sees every rank's uncertified count (gathered trailers / meta words): the (rare)
        # completion and the repeated exchange are taken by ALL ranks or by none
        if out.get("xchg") is not None:
            out["xchg"].note_replay()
            counts = out["xchg"].metas()
        else:
            stride = sess.pack.numel()
            counts = out["gathered"].view(self.world, stride)[:, sess.block_bytes: sess.block_bytes + 4].contiguous()
            counts = counts.view(torch.int32).view(-1).tolist()
        if any(c > 0 for c in counts):
            if counts[self.rank] > 0:
                sess.finish_uncertified(counts[self.rank])
            out = self._post_nccl(sess, want_vote, T)
        self.bank.last_stats = {"path": "tensor+graph", "uncertified": int(sum(counts)),
                                "nsplit": int(sess.plan.nsplit), "kc": int(sess.plan.kc), "cap": int(sess.plan.cap),
                                "workspace_bytes": int(sess.plan.bytes), "sample_rows": int(sess.plan.sample_rows),
                                "chunk_w": int(sess.plan.chunk_w)}
        return out

    def _submit(self, queries, k: int, T, want: str) -> PendingStep:
        q, kind = _as_2d_f32(queries, "queries")
        self._check_k(k)
        sync = (lambda: self.predict(q, k, T=T)) if want == "pred" else (lambda: self.topk(q, k))
        ok = (self.exchange == "peer" and q.is_cuda and q.shape[0] > 0 and all(
            GalleryBank.tensor_path_for(self.plan.size(r), int(k)) for r in range(self.world)))
        if ok and (want == "topk" or self.bank.labels is not None):
            with torch.cuda.device(self.device):
                lane, st = self._lane(q.shape[0], self.plan.size(0))
                cur = torch.cuda.current_stream()
                if st is not None:
                    st.wait_stream(cur)      # the queries were produced on the caller's stream
                with torch.cuda.stream(st if st is not None else cur):
                    sess = self._packed_session(q.shape[0], int(k), want == "pred", T, lane)
                    if sess is not None and (want == "topk" or sess.out_lab is not None):
                        sess.run(q, check=False)
                        self.last_session = sess
                        out = sess.post_out
                        xc = out["xchg"]
                        xc.note_replay()
                        slot, step = xc.header_async()
                        res = out["pred"].clone() if want == "pred" else (out["sims"].clone(), out["idx"].clone())
                        _hand_over(res, st, cur)
                        ev = torch.cuda.Event()
                        ev.record()
                        return PendingStep(ev, slot, lambda f: any(c > 0 for c in xc.metas_of(f, step)), res, sync)
        return PendingStep(None, None, None, sync(), None)

    def submit_predict(self, queries, k: int, *, T=None) -> PendingStep:
        """Pipelined ``predict`` for DEVICE queries (peer exchange only; anything else completes
        synchronously inside this call): the step is enqueued, the header of the peer region that
        carries every rank's uncertified count is read back asynchronously, and ``result()`` -> pred
        [Q] int64 on the device decides -- identically on every rank -- whether the batch has to be
        redone through the synchronous path.  Up to two steps may be in flight (the peer regions are
        double-buffered)."""
        return self._submit(queries, k, T, "pred")

    def submit_topk(self, queries, k: int) -> PendingStep:
        """Pipelined ``topk`` for DEVICE queries: ``result()`` -> (sims, idx) [Q, k] on the device."""
        return self._submit(queries, k, None, "topk")

    def topk(self, queries, k: int, *, mode: str = "auto", with_labels: bool = False):
        q, kind = _as_2d_f32(queries, "queries")
        self._check_k(k)
        with torch.cuda.device(self.device):
            if not q.is_cuda:
                q = q.contiguous().to(self.device, non_blocking=True)
            # every rank must take the same branch: the plan (not the data) decides
            packed_ok = mode == "auto" and q.shape[0] > 0 and all(
                GalleryBank.tensor_path_for(self.plan.size(r), int(k)) for r in range(self.world))
            fast = self._step_packed(q, int(k), False, None) if packed_ok else None
            if fast is not None and (fast["lab"] is not None or not with_labels):
                o_s, o_i, o_l = fast["sims"], fast["idx"], fast["lab"]
                if with_labels:
                    return o_s, o_i, o_l
                if kind == "torch_cuda":
                    return o_s, o_i
                return tuple(_to_host_many([o_s, o_i], kind))
            sims, idx = self._local(q, int(k), mode)
            lab = self.bank.neighbour_labels(idx) if with_labels else None
            g_s, g_i, g_l = exchange_candidates(sims, idx, lab, self.group)
            o_s, o_i, o_l = merge_topk(g_s, g_i, g_l, int(k))
        if with_labels:
            return o_s, o_i, o_l
        if kind == "torch_cuda":
            return o_s, o_i
        return tuple(_to_host_many([o_s, o_i], kind))

    def predict(self, queries, k: int, *, T=None, mode: str = "auto"):
        q, kind = _as_2d_f32(queries, "queries")
        self._check_k(k)
        packed_ok = mode == "auto" and q.shape[0] > 0 and self.bank.labels is not None and all(
            GalleryBank.tensor_path_for(self.plan.size(r), int(k)) for r in range(self.world))
        if packed_ok:
            with torch.cuda.device(self.device):
                if not q.is_cuda:
                    q = q.contiguous().to(self.device, non_blocking=True)
                out = self._step_packed(q, int(k), True, T)
            if out is not None:
                return _to_host(out["pred"], kind)
        o_s, o_i, o_l = self.topk(queries, k, mode=mode, with_labels=True)
        with torch.cuda.device(self.device):
            pred_idx = self.bank.vote(o_s, o_l, T=T)
            cls = torch.from_numpy(np.asarray(self.bank.classes_).astype(np.int64)).to(self.device)
            pred = cls[pred_idx.long()]
        return _to_host(pred, kind)


def choose_sharding(n_gallery: int, n_queries: int, world: int, *, d: int = 768,
                    replica_budget_bytes: float = 60e9) -> str:
    """"query" or "gallery" (load balance).  A gallery that fits every GPU
    next to its workspaces and a query batch with at least four 128-row query tiles per rank is
    served faster by REPLICAS that split the queries (no per-rank fixed cost is repeated, the
    only exchange is the final [Q] / [Q, k] result gather); a big gallery or a small query batch
    (the streaming regime) is split by gallery rows."""
    fits = n_gallery * d * 6.0 <= replica_budget_bytes      # fp32 + bf16 bank per replica
    enough = n_queries >= world * 4 * 128
    return "query" if (fits and enough) else "gallery"


def gather_rows(local: torch.Tensor, sp: ShardPlan, group=None) -> torch.Tensor:
    """ONE all-gather of per-rank row blocks of (possibly) unequal height sp.size(r) -> the
    concatenation [sp.n, ...] on every rank (blocks are padded to the tallest for the collective)."""
    world = dist.get_world_size(group)
    hmax = max(sp.size(r) for r in range(world))
    pad = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(sp.size(r) == hmax for r in range(world)):
        return out
    return torch.cat([out[r * hmax: r * hmax + sp.size(r)] for r in range(world)], 0)


class QueryShardedGallery(_Lanes):
    """Every rank holds the whole gallery; rank r answers the contiguous query slice
    [ShardPlan(Q, world).start(r), stop(r)) and every rank receives everybody's answers.  Results are
    bit-identical to the single-GPU result by construction (each query is answered by exactly the
    single-GPU code path).

    Exchange "peer": the step is ONE graph launch per rank -- K3's tail stores each query's
    prediction (``predict``) or packed top-k rows (``topk``) straight into every rank's peer region
    and a one-warp kernel waits for the other ranks' blocks; "nccl": one all-gather per result array."""

    def __init__(self, features, labels=None, *, group=None, device=None, classes=None, exchange: str | None = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.bank = GalleryBank(features, labels, device=device, classes=classes)
        self.device = self.bank.device
        self.exchange = _resolve_exchange(exchange, group, self.device)
        self.profile = False       # bench.py: capture per-kernel events inside the graph
        self.last_session = None
        self._init_lanes()

    def _slice(self, q: torch.Tensor):
        sp = ShardPlan(int(q.shape[0]), self.world)
        return sp, q[sp.start(self.rank):sp.stop(self.rank)]

    def _gather_rows(self, local: torch.Tensor, sp: ShardPlan) -> torch.Tensor:
        return gather_rows(local, sp, self.group)

    def _tail_peer(self, sess, tail, payload: int, nbytes: int):
        """K3's tail as the producer: every query's CTA stores its prediction (or its packed top-k
        rows) into every rank's peer region over NVLink; the last CTA signals arrival."""
        xc = getattr(sess, "xchg", None)
        if xc is None:   # eager warm-up pass of the session body: collective allocation
            xc = sess.xchg = PeerExchange(-(-nbytes // 16) * 16, group=self.group, device=self.device)
        xc.fill_tail(tail, payload, nbytes)

    def _post_peer(self, sess):
        """Captured behind K3: a one-warp kernel waits for everybody's block and completes the step."""
        self.bank.launches += 1
        sess.xchg.enqueue_wait()
        return {"xchg": sess.xchg}

    def close(self):
        """COLLECTIVE: drop the cached sessions and close their peer channels (IPC mappings, regions)."""
        self.bank.drop_sessions()

    def _issue_peer(self, mine: torch.Tensor, sp: ShardPlan, k: int, T, want: str = "pred", lane: int = 0):
        """Enqueue the whole step (one graph launch per rank incl. the result exchange); None if the
        tensor path does not apply.  -> (session, exchange, sizes, hmax)

        Every rank captures the SAME shape: its slice padded to ``hmax`` rows (slices differ by at
        most one row; the pad repeats the rank's last query and its answer is dropped).  The cached
        session -- and the collectively constructed PeerExchange it owns -- is therefore keyed on
        rank-independent values only, so all ranks hit or miss the cache together (a per-rank slice
        height in the key let Q=1024 then Q=1023 on 8 ranks desynchronise: seven ranks replayed the old
        graph while the eighth entered a collective allocation alone)."""
        sizes = [sp.size(r) for r in range(self.world)]
        hmax = max(sizes)
        if min(sizes) < 1 or not self.bank.use_tensor_path(hmax, k):   # same decision on every rank
            return None
        if want == "pred" and self.bank.labels is None:
            raise ValueError("this gallery was built without labels")
        if mine.shape[0] < hmax:
            mine = torch.cat([mine, mine[-1:].expand(hmax - mine.shape[0], -1)], 0)
        if want == "pred":
            sess = self.bank.session(hmax, k, T=T, profile=self.profile,
                                     tail_hook=lambda s_, t_: self._tail_peer(s_, t_, _lib.PAYLOAD_PRED, hmax * 8),
                                     post=self._post_peer, post_key=("query-sharded", self.world, hmax, lane))
        else:
            sess = self.bank.session(hmax, k, vote=False, pack=True, profile=self.profile,
                                     tail_hook=lambda s_, t_: self._tail_peer(s_, t_, _lib.PAYLOAD_BLOCK, s_.block_bytes),
                                     post=self._post_peer, post_key=("query-sharded-topk", self.world, hmax, lane))
        if sess is None:
            return None
        sess.run(mine, check=False)
        self.last_session = sess
        xc = sess.post_out["xchg"]
        xc.note_replay()
        return sess, xc, sizes, hmax

    def _gathered_preds(self, xc, sizes, hmax) -> torch.Tensor:
        g = xc.gathered()[:, : hmax * 8]
        if all(sz == hmax for sz in sizes) and xc.stride == hmax * 8:
            return g.reshape(-1).view(torch.int64).clone()   # the region is reused two steps later
        return torch.cat([g[r, : sizes[r] * 8].contiguous().view(torch.int64) for r in range(self.world)], 0)

    def _gathered_topk(self, xc, sizes, hmax, k):
        """(sims [Q, k], idx [Q, k]) from the packed blocks (idx | sims [| labels]) of all ranks."""
        g = xc.gathered()
        e = hmax * k
        if self.world > 1 and all(sz == hmax for sz in sizes):   # equal slices: one strided copy per array instead of a cat of `world` views
            return (g[:, e * 8: e * 12].view(torch.float32).reshape(-1, k),
                    g[:, : e * 8].view(torch.int64).reshape(-1, k))
        idx = torch.cat([g[r, : e * 8].view(torch.int64).view(hmax, k)[: sizes[r]] for r in range(self.world)], 0)
        sims = torch.cat([g[r, e * 8: e * 12].view(torch.float32).view(hmax, k)[: sizes[r]] for r in range(self.world)], 0)
        return sims, idx   # torch.cat copies: the region is reused two steps later

    def _stats(self, sess, counts):
        self.bank.last_stats = {"path": "tensor+graph", "uncertified": int(sum(counts)),
                                "nsplit": int(sess.plan.nsplit), "kc": int(sess.plan.kc), "cap": int(sess.plan.cap),
                                "workspace_bytes": int(sess.plan.bytes), "sample_rows": int(sess.plan.sample_rows),
                                "chunk_w": int(sess.plan.chunk_w), "exchange": "peer"}

    def _submit(self, queries, k: int, T, want: str) -> PendingStep:
        q, kind = _as_2d_f32(queries, "queries")
        sync = (lambda: self.predict(q, k, T=T)) if want == "pred" else (lambda: self.topk(q, k))
        if self.exchange == "peer" and q.is_cuda:
            sp, mine = self._slice(q)
            with torch.cuda.device(self.device):
                lane, st = self._lane(sp.size(0), self.bank.n)
                cur = torch.cuda.current_stream()
                if st is not None:
                    st.wait_stream(cur)      # the queries were produced on the caller's stream
                with torch.cuda.stream(st if st is not None else cur):
                    issued = self._issue_peer(mine, sp, int(k), T, want, lane)
                    if issued is not None:
                        sess, xc, sizes, hmax = issued
                        slot, step = xc.header_async()
                        res = self._gathered_preds(xc, sizes, hmax) if want == "pred" else \
                            self._gathered_topk(xc, sizes, hmax, int(k))
                        _hand_over(res, st, cur)
                        ev = torch.cuda.Event()
                        ev.record()
                        return PendingStep(ev, slot, lambda f: any(c > 0 for c in xc.metas_of(f, step)), res, sync)
        return PendingStep(None, None, None, sync(), None)

    def submit_predict(self, queries, k: int, *, T=None) -> PendingStep:
        """Pipelined ``predict`` for DEVICE queries (see ShardedGallery.submit_predict)."""
        return self._submit(queries, k, T, "pred")

    def submit_topk(self, queries, k: int) -> PendingStep:
        """Pipelined ``topk`` for DEVICE queries: ``result()`` -> (sims, idx) on the device."""
        return self._submit(queries, k, None, "topk")

    def _predict_peer(self, mine: torch.Tensor, sp: ShardPlan, k: int, T):
        """Whole step = one graph launch per rank incl. the result exchange; None if not applicable."""
        issued = self._issue_peer(mine, sp, k, T, "pred")
        if issued is None:
            return None
        sess, xc, sizes, hmax = issued
        counts = xc.metas()
        self._stats(sess, counts)
        if any(c > 0 for c in counts):   # rare: every rank takes the completion + collective gather
            pred = sess.pred
            if counts[self.rank] > 0:
                sess.finish_uncertified(counts[self.rank])
                pred = sess._tail()
            return self._gather_rows(pred[: sizes[self.rank]], sp)   # the session batch is padded to hmax
        return self._gathered_preds(xc, sizes, hmax)

    def _topk_peer(self, mine: torch.Tensor, sp: ShardPlan, k: int):
        issued = self._issue_peer(mine, sp, k, None, "topk")
        if issued is None:
            return None
        sess, xc, sizes, hmax = issued
        counts = xc.metas()
        self._stats(sess, counts)
        if any(c > 0 for c in counts):   # rare: every rank takes the completion + collective gathers
            if counts[self.rank] > 0:
                sess.finish_uncertified(counts[self.rank])
            n = sizes[self.rank]
            return self._gather_rows(sess.out_sim[:n].contiguous(), sp), self._gather_rows(sess.out_idx[:n].contiguous(), sp)
        return self._gathered_topk(xc, sizes, hmax, k)

    def predict(self, queries, k: int, *, T=None, mode: str = "auto"):
        q, kind = _as_2d_f32(queries, "queries")
        sp, mine = self._slice(q)
        with torch.cuda.device(self.device):
            if not mine.is_cuda:
                mine = mine.contiguous().to(self.device, non_blocking=True)
            if self.exchange == "peer" and mode == "auto":
                out = self._predict_peer(mine, sp, int(k), T)
                if out is not None:
                    return _to_host(out, kind)
            sess = self.bank.session(mine.shape[0], int(k), T=T, profile=self.profile) if mode == "auto" else None
            if sess is not None:
                pred, _, _ = sess.run(mine)   # the whole local step is one CUDA-graph launch
                self.last_session = sess
            elif mine.shape[0] > 0:
                pred = self.bank.predict(mine, int(k), T=T, mode=mode)
            else:
                pred = torch.empty((0,), dtype=torch.int64, device=self.device)
            out = self._gather_rows(pred, sp)
        return _to_host(out, kind)

    def topk(self, queries, k: int, *, mode: str = "auto"):
        q, kind = _as_2d_f32(queries, "queries")
        sp, mine = self._slice(q)
        with torch.cuda.device(self.device):
            if not mine.is_cuda:
                mine = mine.contiguous().to(self.device, non_blocking=True)
            out = self._topk_peer(mine, sp, int(k)) if (self.exchange == "peer" and mode == "auto") else None
            if out is not None:
                o_s, o_i = out
            else:
                sims, idx = self.bank._topk_device(mine, int(k), mode)
                o_s, o_i = self._gather_rows(sims, sp), self._gather_rows(idx, sp)
        if kind == "torch_cuda":
            return o_s, o_i
        return tuple(_to_host_many([o_s, o_i], kind))
