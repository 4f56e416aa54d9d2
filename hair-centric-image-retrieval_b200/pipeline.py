"""Host-buffer serving loop: query batches arrive in HOST memory and the answers are wanted in HOST
memory (the reference's call sites hand numpy arrays to sklearn / torch CPU tensors to ``topk``:
classification_engine.py:80-82, qualitative_test.py:79-84).  A synchronous call pays, per batch,
H2D copy -> search -> D2H copy one after the other; ``HostPipeline`` keeps the three on three CUDA
streams so that the copy of batch i+1 and the read-back of batch i-1 run beside the search of batch i:

    copy-in stream :  H2D(i+1) ............
    search stream  :  ........ search(i) ........ search(i+1)
    copy-out stream:  D2H(i-1) ..................... D2H(i)

The search itself is the ordinary pipelined submission (``SearchSession.submit`` /
``QueryShardedGallery.submit_*`` / ``ShardedGallery.submit_predict``): same kernels, same exactness
protocol -- the (rare) uncertified batch is redone in ``result()``."""
from __future__ import annotations

import numpy as np
import torch

from .engine import GalleryBank, PendingStep, _to_host_many


def _flatten(res):
    return [res] if isinstance(res, torch.Tensor) else [t for t in res]


class HostPending:
    """One submitted host batch.  ``result()`` -> numpy arrays (or CPU tensors when the batch was a
    torch tensor) in pinned host memory owned by the caller: ``pred [Q]`` or ``(sims [Q,k], idx [Q,k])``."""

    def __init__(self, pend: PendingStep, pick, host, ev_out, kind: str, single: bool):
        self._pend, self._pick, self._host, self._ev_out, self._kind, self._single = pend, pick, host, ev_out, kind, single
        self._done = None
        self.redone = False      # set by result(): the batch was uncertified and went through the redo path

    def result(self):
        if self._done is None:
            dev = self._pend.result()          # search-stream event + the exactness check (may redo)
            self.redone = self._pend.redone
            if self.redone:
                host = _to_host_many(_flatten(self._pick(dev)), "torch_cpu")
            else:
                self._ev_out.synchronize()
                host = self._host
            out = [h.numpy() if self._kind == "numpy" else h for h in host]
            self._done = out[0] if self._single else tuple(out)
            self._pend = self._host = self._ev_out = None
        return self._done


class HostPipeline:
    """``submit(host batch) -> HostPending`` for batches of one fixed shape ``[nq, d]``.

    ``submit_device(q_dev) -> PendingStep`` is the device-side pipelined submission; ``pick`` selects
    what travels back to the host from its device results; ``rows`` = the row range of the batch this
    rank needs on its device (a query-sharded rank copies only its own slice).  ``depth`` batches may
    be pending; submitting one more first completes the oldest."""

    MAX_DEPTH = 4   # the pinned rings behind PendingStep hold 8 entries

    def __init__(self, submit_device, nq: int, d: int, *, device, pick=lambda r: r, rows=None, depth: int = 2):
        if not 1 <= depth <= self.MAX_DEPTH:
            raise ValueError(f"depth must be in [1, {self.MAX_DEPTH}]")
        self.device = torch.device(device)
        self.nq, self.d, self.depth = int(nq), int(d), int(depth)
        self.rows = (0, self.nq) if rows is None else (int(rows[0]), int(rows[1]))
        self._submit_device, self._pick = submit_device, pick
        with torch.cuda.device(self.device):
            self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
            # rows outside ``rows`` are never read by this rank's search; zero so that they are defined
            self._stage = [torch.zeros((self.nq, self.d), dtype=torch.float32, device=self.device) for _ in range(depth)]
            self._s_in.wait_stream(torch.cuda.current_stream())   # the zero fill precedes the first copy-in
        self._slots = [None] * depth
        self._n = 0

    # ---- the usual targets -------------------------------------------------------------------
    @classmethod
    def for_bank(cls, bank: GalleryBank, nq: int, k: int, *, want: str = "topk", T=None, depth: int = 2):
        """One GPU.  want="topk" -> (sims, idx); "pred" -> predictions (uniform vote, or T-weighted)."""
        if want not in ("topk", "pred"):
            raise ValueError("want must be 'topk' or 'pred'")
        sess = bank.session(int(nq), int(k), T=T, vote=(want == "pred"))
        if sess is None:
            raise ValueError("this shape takes the exact CUDA-core path (small gallery): call bank.topk / bank.predict")
        pick = (lambda r: (r[1], r[2])) if want == "topk" else (lambda r: r[0])
        p = cls(sess.submit, nq, bank.d, device=bank.device, pick=pick, depth=depth)
        p.session = sess
        return p

    @classmethod
    def for_gallery(cls, gal, nq: int, k: int, *, want: str = "topk", T=None, depth: int = 2):
        """Multi-GPU galleries (sharded.py), peer exchange.  COLLECTIVE: every rank builds and drives
        the pipeline alike.  Every rank receives the whole answer, as from ``gal.topk`` / ``gal.predict``."""
        from .sharded import QueryShardedGallery, ShardPlan
        if want == "pred":
            sub = lambda q: gal.submit_predict(q, int(k), T=T)   # noqa: E731
        elif hasattr(gal, "submit_topk"):
            sub = lambda q: gal.submit_topk(q, int(k))           # noqa: E731
        else:
            raise ValueError(f"{type(gal).__name__} has no pipelined top-k submission")
        rows = None
        if isinstance(gal, QueryShardedGallery):
            sp = ShardPlan(int(nq), gal.world)
            rows = (sp.start(gal.rank), sp.stop(gal.rank))
        return cls(sub, nq, gal.bank.d, device=gal.device, rows=rows, depth=min(depth, 2))

    # ---- one batch ---------------------------------------------------------------------------
    def submit(self, queries) -> HostPending:
        if isinstance(queries, torch.Tensor):
            kind, qh = "torch_cpu", queries
        else:
            kind, qh = "numpy", torch.from_numpy(np.ascontiguousarray(np.asarray(queries, dtype=np.float32)))
        if qh.is_cuda:
            raise ValueError("HostPipeline takes host batches; device batches go to the submit_* calls directly")
        if tuple(qh.shape) != (self.nq, self.d) or qh.dtype != torch.float32:
            raise ValueError(f"pipeline was built for fp32 batches of shape {(self.nq, self.d)}, got "
                             f"{tuple(qh.shape)} {qh.dtype}")
        j = self._n % self.depth
        self._n += 1
        if self._slots[j] is not None:
            self._slots[j].result()          # its staging buffer (kept for the redo path) is about to be reused
        a, b = self.rows
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            stage = self._stage[j]
            with torch.cuda.stream(self._s_in):
                stage[a:b].copy_(qh[a:b], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record()
            main.wait_event(ev_in)
            pend = self._submit_device(stage)
            dev = _flatten(self._pick(pend.device_results))
            ev_main = pend.ready_event      # recorded on the stream that produced the results (a lane stream, maybe)
            if ev_main is None:
                ev_main = torch.cuda.Event()
                ev_main.record()
            host = []
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(ev_main)
                for t in dev:
                    t.record_stream(self._s_out)
                    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    h.copy_(t, non_blocking=True)
                    host.append(h)
                ev_out = torch.cuda.Event()
                ev_out.record()
        hp = HostPending(pend, self._pick, host, ev_out, kind, single=isinstance(self._pick(pend.device_results), torch.Tensor))
        self._slots[j] = hp
        return hp

    def drain(self):
        for s in self._slots:
            if s is not None:
                s.result()
        self._slots = [None] * self.depth
