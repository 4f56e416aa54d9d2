"""Host side of the hot path: a device-resident, L2-normalised gallery bank and the exact
cosine top-k / kNN-vote over it.  PyTorch is plumbing only (device memory, streams); every
arithmetic step is one of the hand-written sm_100a kernels behind the C ABI
(include/hcir_b200.h).  There is no CPU fallback.

Reference call sites this replaces (all CPU library calls in the reference):
  * ``F.normalize`` + ``torch.cat`` feature bank   HairPretraining/src/classification_engine.py:50,62,66-69
  * ``KNeighborsClassifier(k, metric="cosine")``  HairPretraining/src/classification_engine.py:80-82
  * ``torch.mm(q, G.t())`` + ``torch.topk``        experiments/DualViewHair/scripts/qualitative_test.py:79-84
  * ``cosine_similarity`` + ``np.argsort``          src/models/hair_encoder.py:193-194
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from ._lib import Plan, Tail

_SM_COUNT_CACHE: dict[int, int] = {}


def _sm_count(device: torch.device) -> int:
    i = device.index if device.index is not None else torch.cuda.current_device()
    if i not in _SM_COUNT_CACHE:
        _SM_COUNT_CACHE[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SM_COUNT_CACHE[i]


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("hcir_b200 needs a CUDA device (B200, sm_100); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"hcir_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as_2d_f32(x, name: str):
    """Accept numpy / torch (cpu or cuda); return (torch tensor fp32 2-D, kind)."""
    if isinstance(x, torch.Tensor):
        kind = "torch_cuda" if x.is_cuda else "torch_cpu"
        t = x
    else:
        kind = "numpy"
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [rows, dim], got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        t = t.float()
    return t, kind


def _to_host(t: torch.Tensor, kind: str):
    """Device result -> the caller's flavour (numpy / cpu tensor / cuda tensor)."""
    if kind == "torch_cuda":
        return t
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy() if kind == "numpy" else h


def _to_host_many(ts, kind: str):
    """Several device results -> the caller's flavour with ONE stream synchronisation."""
    if kind == "torch_cuda":
        return list(ts)
    hs = []
    for t in ts:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        hs.append(h)
    torch.cuda.current_stream().synchronize()
    return [h.numpy() if kind == "numpy" else h for h in hs]


def l2_normalize(x: torch.Tensor, *, want_f32=True, want_bf16=True, want_delta=True, pad_rows_to: int = 1):
    """K1 on a device tensor [n, d] fp32 -> (unit fp32 [n, ld] | None, unit bf16 [n, ld] | None,
    delta [n] | None) with ld = d rounded up to 64 (zero padded).  ``F.normalize(x, dim=1)``
    semantics (classification_engine.py:50).  ``pad_rows_to``: the bf16 output is a view of a
    zero-padded buffer whose row count is a multiple of it (query side of the contraction)."""
    lib = _lib.load()
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
        raise ValueError("l2_normalize expects a 2-D fp32 CUDA tensor")
    if x.stride(1) != 1:
        x = x.contiguous()
    n, d = x.shape
    ld = lib.hcir_padded_dim(d)
    dev = x.device
    o32 = torch.empty((n, ld), dtype=torch.float32, device=dev) if want_f32 else None
    obf = None
    if want_bf16:
        n_alloc = -(-n // pad_rows_to) * pad_rows_to
        buf = torch.empty((n_alloc, ld), dtype=torch.bfloat16, device=dev)
        if n_alloc > n:
            buf[n:].zero_()
        obf = buf[:n]
    dl = torch.empty((n,), dtype=torch.float32, device=dev) if want_delta else None
    with torch.cuda.device(dev):
        _lib.check(lib.hcir_l2norm_cast(x.data_ptr(), n, d, x.stride(0),
                                        o32.data_ptr() if want_f32 else None,
                                        obf.data_ptr() if want_bf16 else None, ld,
                                        dl.data_ptr() if want_delta else None, _stream_ptr()), "l2norm_cast")
    return o32, obf, dl


class GalleryBank:
    """Device-resident normalised feature bank (fp32 unit rows for the exact re-score, bf16
    unit rows for the tensor-core contraction) + optional labels.

    ``features`` [N, D] may be un-normalised (hair_encoder.py embeddings) or already unit
    (classification_engine.py bank): normalisation is idempotent up to fp32 rounding.
    ``labels`` are arbitrary integers; they are mapped to class indices through
    ``classes_ = unique(labels)`` exactly like sklearn's ``fit``."""

    def __init__(self, features, labels=None, *, device=None, idx_offset: int = 0, classes=None,
                 chunk_rows: int = 1 << 18):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        lazy_numpy = isinstance(features, np.ndarray) and features.ndim == 2
        if lazy_numpy:
            feats = features  # (memory-mapped) numpy: converted chunk by chunk below, never copied whole
        else:
            feats, _ = _as_2d_f32(features, "features")
        n, d = feats.shape
        self.n, self.d = int(n), int(d)
        self.ld = self.lib.hcir_padded_dim(d)
        self.idx_offset = int(idx_offset)
        self.sm_count = _sm_count(self.device)
        with torch.cuda.device(self.device):
            self.g32 = torch.empty((n, self.ld), dtype=torch.float32, device=self.device)
            self.gbf = torch.empty((n, self.ld), dtype=torch.bfloat16, device=self.device)
            dmax = torch.zeros((), dtype=torch.float32, device=self.device)
            for a in range(0, n, chunk_rows):
                b = min(n, a + chunk_rows)
                x = feats[a:b]
                if lazy_numpy:
                    x = torch.from_numpy(np.array(x, dtype=np.float32, order="C"))  # copy: memmaps are read-only
                if not x.is_cuda:
                    x = x.contiguous().to(self.device, non_blocking=True)
                elif x.device != self.device:
                    x = x.to(self.device)
                if x.stride(1) != 1:
                    x = x.contiguous()
                dl = torch.empty((b - a,), dtype=torch.float32, device=self.device)
                _lib.check(self.lib.hcir_l2norm_cast(x.data_ptr(), b - a, d, x.stride(0),
                                                     self.g32[a:b].data_ptr(), self.gbf[a:b].data_ptr(),
                                                     self.ld, dl.data_ptr(), _stream_ptr()), "l2norm_cast")
                dmax = torch.maximum(dmax, dl.max())
            self.g_delta_max = float(dmax.item())
        # conservative bound on the tensor-core fp32 accumulation error + canonical-dot rounding
        self.eps_acc = float(self.ld) * 2.0 ** -22
        self.labels = None
        self.classes_ = None
        if labels is not None:
            self.set_labels(labels, classes)
        self.last_stats = {}
        self.retry_stats = {}   # how the last batch of uncertified queries was completed
        # bench.py sets this to a list to get (name, start_event, end_event) per kernel launch
        self.kernel_events = None
        self.launches = 0  # number of hcir kernels launched by this bank since construction
        self.k3_width = 0  # measurement aid: force K3's CTA width (1 = 128, 2 = 256, 3 = 1024 threads; 0 = auto)

    def _timed(self, name, fn, n_kernels=1):
        """Run one C-ABI launch; optionally bracket it with CUDA events on the current stream."""
        self.launches += n_kernels
        if self.kernel_events is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn()
        b.record()
        self.kernel_events.append((name, a, b))
        return rc

    # ------------------------------------------------------------------ labels
    def set_labels(self, labels, classes=None):
        y = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
        y = y.reshape(-1)
        if y.shape[0] != self.n:
            raise ValueError(f"labels has {y.shape[0]} entries, gallery has {self.n} rows")
        self.classes_ = np.unique(y) if classes is None else np.asarray(classes)
        cls_idx = np.searchsorted(self.classes_, y)
        if (cls_idx >= len(self.classes_)).any() or (self.classes_[np.minimum(cls_idx, len(self.classes_) - 1)] != y).any():
            raise ValueError("labels contain values missing from `classes`")
        # captured graphs hold raw pointers to the old label / class tensors: drop them (collective
        # if multi-GPU sessions exist -- sharded callers change labels on every rank together)
        self.drop_sessions()
        self.labels = torch.from_numpy(cls_idx.astype(np.int32)).to(self.device)
        self._cls_dev = None

    # ------------------------------------------------------------------ planning
    def choose_kc(self, k: int) -> int:
        return 2 * k + 64

    @staticmethod
    def tensor_path_for(n: int, k: int) -> bool:
        # tiny galleries are not worth a tcgen05 launch: the exact CUDA-core kernel is the path
        return n >= max(4096, 8 * (2 * k + 64))

    def use_tensor_path(self, nq: int, k: int) -> bool:
        return self.tensor_path_for(self.n, k)

    # ------------------------------------------------------------------ search
    def topk(self, queries, k: int, *, mode: str = "auto", return_device: bool = False, use_graph: bool = False):
        """Exact cosine top-k: (sims [Q,k] fp32 descending, idx [Q,k] int64), canonical order
        (ties -> ascending index).  ``mode``: "auto" | "tensor" | "exact".  ``use_graph``: serve the
        batch from a cached fixed-shape :class:`SearchSession` (one CUDA-graph launch per call; worth it
        for a caller that repeats the batch shape -- the session owns a workspace)."""
        q, kind = _as_2d_f32(queries, "queries")
        if q.shape[1] != self.d:
            raise ValueError(f"query dim {q.shape[1]} != gallery dim {self.d}")
        k = int(k)
        if not (1 <= k <= self.n):
            raise ValueError(f"k={k} must be in [1, N={self.n}]")
        with torch.cuda.device(self.device):
            sess = self.session(q.shape[0], k, vote=False) if (use_graph and mode == "auto") else None
            if sess is not None:
                # host queries go straight into the step's static input (one H2D copy, no staging tensor)
                _, sims, idx = sess.run(q if q.is_cuda or q.is_pinned() else q.contiguous())
            else:
                if not q.is_cuda:
                    q = q.contiguous().to(self.device, non_blocking=True)
                elif q.device != self.device:
                    q = q.to(self.device)
                sims, idx = self._topk_device(q, k, mode)
            if return_device or kind == "torch_cuda":
                return sims, idx
            sims_h, idx_h = _to_host_many([sims, idx], kind)
            return sims_h, idx_h

    def _topk_device(self, q: torch.Tensor, k: int, mode: str = "auto"):
        sims, idx, _ = self._search(q, k, mode, None)
        return sims, idx

    def _search(self, q: torch.Tensor, k: int, mode: str = "auto", tail=None):
        """Exact top-k of device queries.  ``tail(sims, idx)`` enqueues whatever consumes the
        result (label gather, vote); it runs BEFORE the 4-byte read-back that decides whether the
        exact fallback is needed, so the device never idles on that host round trip, and is run
        again on the (rare) fallback.  Returns (sims, idx, tail result)."""
        lib = self.lib
        nq = q.shape[0]
        dev = self.device
        out_sim = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq == 0:
            return out_sim, out_idx, (tail(out_sim, out_idx) if tail else None)
        if mode not in ("auto", "tensor", "exact"):
            raise ValueError(f"unknown mode {mode!r}")
        tensor = mode == "tensor" or (mode == "auto" and self.use_tensor_path(nq, k))
        q32, qbf, qdl = l2_normalize(q, want_bf16=tensor, want_delta=tensor, pad_rows_to=128)
        self.launches += 1
        st = _stream_ptr()
        if not tensor:
            self._exact(q32, None, nq, k, out_sim, out_idx)
            self.last_stats = {"path": "exact", "uncertified": 0}
            return out_sim, out_idx, (tail(out_sim, out_idx) if tail else None)
        kc = self.choose_kc(k)
        plan = Plan()
        _lib.check(lib.hcir_simtopk_plan(nq, self.n, self.ld, kc, self.sm_count, plan), "simtopk_plan")
        plan.q_rows = -(-nq // 128) * 128  # l2_normalize(pad_rows_to=128) allocated them
        ws = torch.empty((int(plan.bytes),), dtype=torch.uint8, device=dev)
        if self.kernel_events is not None and plan.sample_rows > 0:
            # measurement only: enqueue the two phases separately so each gets its own events
            for name, flag, nk in (("simtopk_sample", 2, 2), ("simtopk", 4, 1)):
                plan.flags = flag
                _lib.check(self._timed(name, lambda: lib.hcir_simtopk(
                    qbf.data_ptr(), nq, self.gbf.data_ptr(), self.n, self.ld, plan, ws.data_ptr(), st),
                    n_kernels=nk), name)
            plan.flags = 0
        else:
            _lib.check(self._timed("simtopk", lambda: lib.hcir_simtopk(
                qbf.data_ptr(), nq, self.gbf.data_ptr(), self.n, self.ld, plan, ws.data_ptr(), st),
                n_kernels=plan.kernels()), "simtopk")
        unc_list = torch.empty((nq,), dtype=torch.int32, device=dev)
        unc_state = torch.zeros((4,), dtype=torch.int32, device=dev)   # [1] = result (include/hcir_b200.h)
        plan.flags = self.k3_width << 8
        _lib.check(self._timed("select_rescore", lambda: lib.hcir_select_rescore(
            q32.data_ptr(), self.g32.data_ptr(), self.ld, nq, self.n, k, self.idx_offset, plan, ws.data_ptr(),
            qdl.data_ptr(), self.g_delta_max, self.eps_acc, out_sim.data_ptr(), out_idx.data_ptr(),
            unc_list.data_ptr(), unc_state.data_ptr(), None, st)), "select_rescore")
        res = tail(out_sim, out_idx) if tail else None
        n_unc = int(unc_state[1].item())  # 4-byte readback: decides whether the exact fallback runs
        if n_unc > 0:
            self._finish_uncertified(q32, qbf, qdl, unc_list, n_unc, k, out_sim, out_idx)
            res = tail(out_sim, out_idx) if tail else None
        self.last_stats = {"path": "tensor", "uncertified": n_unc, "nsplit": int(plan.nsplit),
                           "kc": int(plan.kc), "cap": int(plan.cap), "workspace_bytes": int(plan.bytes),
                           "sample_rows": int(plan.sample_rows), "chunk_w": int(plan.chunk_w)}
        return out_sim, out_idx, res

    def _finish_uncertified(self, q32, qbf, qdl, unc_list, n_unc, k, out_sim, out_idx):
        """Complete the queries K3 could not certify (their kc best bf16 candidates do not provably
        contain the fp32 top-k).  First a SECOND tensor pass over just these queries with an explicit
        threshold: K3 left each one's best-so-far fp32 k-th score s_k, a true top-k row scores
        >= s_k in fp32, hence > s_k - eps in bf16, so the main pass with thr = s_k - eps (no sample
        pass) and a 4x wider kc collects every row that can matter, and K3 certifies against that
        very threshold.  Costs one gallery stream (the streaming regime) instead of an fp32 brute
        force.  What even that cannot certify (more than 4*kc rows inside the eps band: duplicated
        galleries) goes to the exact CUDA-core kernel."""
        lib, dev = self.lib, self.device
        kc = self.choose_kc(k)
        plan = Plan()
        for mult in (4, 3, 2, 1):   # widest kc whose plan still carries per-query thresholds
            kc2 = max(kc, min(mult * kc, 2048))
            _lib.check(lib.hcir_simtopk_plan(n_unc, self.n, self.ld, kc2, self.sm_count, plan), "simtopk_plan")
            if plan.sample_rows > 0:
                break
        if plan.sample_rows <= 0 or qbf is None:
            self._exact(q32, unc_list, n_unc, k, out_sim, out_idx)
            self.retry_stats = {"second_pass": 0, "exact": int(n_unc)}
            return
        sel = unc_list[:n_unc].long()
        dq = qdl[sel]
        eps = self.g_delta_max * (1.0 + dq) + dq * (1.0 + 1e-6) + self.eps_acc
        thr = (out_sim[sel, k - 1] - eps * 1.001 - 1e-6).float().contiguous()
        rows = -(-n_unc // 128) * 128
        qbf2 = torch.zeros((rows, self.ld), dtype=torch.bfloat16, device=dev)
        qbf2[:n_unc] = qbf[sel]
        q32_2 = q32[sel].contiguous()
        dq = dq.contiguous()
        plan.q_rows = rows
        ws = torch.empty((int(plan.bytes),), dtype=torch.uint8, device=dev)
        for off in (int(plan.thr0_off), int(plan.thr_hi_off)):   # the thresholds the sample pass would write
            ws[off: off + 4 * n_unc].view(torch.float32).copy_(thr)
        st = _stream_ptr()
        plan.flags = 4  # HCIR_FLAG_MAIN_ONLY: thresholds are given
        _lib.check(self._timed("simtopk_retry", lambda: lib.hcir_simtopk(
            qbf2.data_ptr(), n_unc, self.gbf.data_ptr(), self.n, self.ld, plan, ws.data_ptr(), st)), "simtopk(retry)")
        plan.flags = 0
        o_s = torch.empty((n_unc, k), dtype=torch.float32, device=dev)
        o_i = torch.empty((n_unc, k), dtype=torch.int64, device=dev)
        unc2 = torch.empty((n_unc,), dtype=torch.int32, device=dev)
        cnt2 = torch.zeros((4,), dtype=torch.int32, device=dev)
        _lib.check(self._timed("select_rescore_retry", lambda: lib.hcir_select_rescore(
            q32_2.data_ptr(), self.g32.data_ptr(), self.ld, n_unc, self.n, k, self.idx_offset, plan, ws.data_ptr(),
            dq.data_ptr(), self.g_delta_max, self.eps_acc, o_s.data_ptr(), o_i.data_ptr(), unc2.data_ptr(),
            cnt2.data_ptr(), None, st)), "select_rescore(retry)")
        n2 = int(cnt2[1].item())
        if n2 < n_unc:
            ok = torch.ones((n_unc,), dtype=torch.bool, device=dev)
            ok[unc2[:n2].long()] = False
            out_sim[sel[ok]] = o_s[ok]
            out_idx[sel[ok]] = o_i[ok]
        if n2 > 0:
            left = unc_list[:n_unc][unc2[:n2].long()].contiguous()
            self._exact(q32, left, n2, k, out_sim, out_idx)
        self.retry_stats = {"second_pass": int(n_unc - n2), "exact": int(n2)}

    def _exact(self, q32, qlist, nlist, k, out_sim, out_idx):
        lib = self.lib
        nbytes = int(lib.hcir_exact_workspace_bytes(nlist, self.n, k, self.sm_count))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        _lib.check(self._timed("exact_topk", lambda: lib.hcir_exact_topk(
            q32.data_ptr(), self.g32.data_ptr(), self.ld, self.n, k, self.idx_offset,
            qlist.data_ptr() if qlist is not None else None, nlist, out_sim.data_ptr(), out_idx.data_ptr(),
            ws.data_ptr(), nbytes, self.sm_count, _stream_ptr()), n_kernels=2), "exact_topk")

    # ------------------------------------------------------------------ vote
    def neighbour_labels(self, idx: torch.Tensor) -> torch.Tensor:
        if self.labels is None:
            raise ValueError("this GalleryBank was built without labels")
        out = torch.empty(idx.shape, dtype=torch.int32, device=self.device)
        self.launches += 1
        _lib.check(self.lib.hcir_gather_labels(idx.data_ptr(), idx.numel(), self.labels.data_ptr(), self.n,
                                               self.idx_offset, out.data_ptr(), _stream_ptr()), "gather_labels")
        return out

    def vote(self, sims: torch.Tensor, nbr_labels: torch.Tensor, *, T=None, return_scores: bool = False):
        """K4 on device tensors: class INDEX [Q] int32 (+ optional [Q, C] scores)."""
        nq, k = sims.shape
        c = len(self.classes_)
        pred = torch.empty((nq,), dtype=torch.int32, device=self.device)
        scores = torch.empty((nq, c), dtype=torch.float32, device=self.device) if return_scores else None
        self.launches += 1
        _lib.check(self.lib.hcir_vote(sims.data_ptr(), nbr_labels.data_ptr(), nq, k, c,
                                      float(T) if T is not None else 0.0, pred.data_ptr(),
                                      scores.data_ptr() if return_scores else None, _stream_ptr()), "vote")
        return (pred, scores) if return_scores else pred

    def vote_from_idx(self, sims: torch.Tensor, idx: torch.Tensor, *, T=None, nbr_out: torch.Tensor | None = None):
        """Fused K4 tail on device tensors: neighbour-label gather -> vote -> ORIGINAL label value
        [Q] int64 (one launch; ``nbr_out`` optionally receives the [Q, k] neighbour class indices)."""
        if self.labels is None:
            raise ValueError("this GalleryBank was built without labels")
        nq, k = sims.shape
        cls = self._classes_device()
        pred = torch.empty((nq,), dtype=torch.int64, device=self.device)
        self.launches += 1
        _lib.check(self.lib.hcir_vote_idx(sims.data_ptr(), idx.data_ptr(), self.labels.data_ptr(), self.n,
                                          self.idx_offset, nq, k, len(self.classes_),
                                          float(T) if T is not None else 0.0, cls.data_ptr(), pred.data_ptr(),
                                          nbr_out.data_ptr() if nbr_out is not None else None, _stream_ptr()),
                   "vote_idx")
        return pred

    def vote_from_labels(self, sims: torch.Tensor, nbr_labels: torch.Tensor, *, T=None):
        """K4 on gathered neighbour class indices -> ORIGINAL label value [Q] int64 (one launch)."""
        nq, k = sims.shape
        cls = self._classes_device()
        pred = torch.empty((nq,), dtype=torch.int64, device=self.device)
        self.launches += 1
        _lib.check(self.lib.hcir_vote_classes(sims.data_ptr(), nbr_labels.data_ptr(), nq, k, len(self.classes_),
                                              float(T) if T is not None else 0.0, cls.data_ptr(), pred.data_ptr(),
                                              _stream_ptr()), "vote_classes")
        return pred

    def predict(self, queries, k: int, *, T=None, mode: str = "auto", return_neighbors: bool = False):
        """kNN classification: predicted ORIGINAL label values [Q] int64.  ``T=None`` is the
        reference's uniform vote; ``T>0`` the temperature-weighted extension."""
        q, kind = _as_2d_f32(queries, "queries")
        with torch.cuda.device(self.device):
            if not q.is_cuda:
                q = q.contiguous().to(self.device, non_blocking=True)
            elif q.device != self.device:
                q = q.to(self.device)
            self._classes_device()

            def tail(s, i):
                return self.vote_from_idx(s, i, T=T)

            sims, idx, pred = self._search(q, int(k), mode, tail)
            if return_neighbors:
                return _to_host(pred, kind), _to_host(sims, kind), _to_host(idx, kind)
            return _to_host(pred, kind)

    def _classes_device(self) -> torch.Tensor:
        if self.labels is None:
            raise ValueError("this GalleryBank was built without labels")
        if getattr(self, "_cls_dev", None) is None or self._cls_dev.shape[0] != len(self.classes_):
            self._cls_dev = torch.from_numpy(np.asarray(self.classes_).astype(np.int64)).to(self.device)
        return self._cls_dev

    def predict_multi_k(self, queries, ks, *, T=None, mode: str = "auto"):
        """One search at max(ks), one prefix vote per k (the reference recomputes the whole
        distance matrix per k: classification_engine.py:79-82).  Returns {k: pred [Q]}."""
        ks = [int(k) for k in ks]
        q, kind = _as_2d_f32(queries, "queries")
        with torch.cuda.device(self.device):
            if not q.is_cuda:
                q = q.contiguous().to(self.device, non_blocking=True)
            elif q.device != self.device:
                q = q.to(self.device)
            sims, idx = self._topk_device(q, max(ks), mode)
            nl = self.neighbour_labels(idx)
            cls = torch.from_numpy(self.classes_.astype(np.int64)).to(self.device)
            out = {}
            for k in ks:
                p = self.vote(sims[:, :k].contiguous(), nl[:, :k].contiguous(), T=T)
                out[k] = _to_host(cls[p.long()], kind)
            return out


# ---------------------------------------------------------------------------------------------
# encoder -> bank without leaving the device
# ---------------------------------------------------------------------------------------------
class FeatureBankBuilder:
    """Device-side replacement of the feature-bank loop of ``Classifier.extracting_features``
    (HairPretraining/src/classification_engine.py:42-53 and :66-67: per batch ``F.normalize`` on
    the GPU, ``.cpu()``, and one ``torch.cat`` at the end).  ``append`` takes the encoder's
    [B, D] output where it is -- on the device, un-normalised, fp32 / fp16 / bf16 -- and K1 writes
    the unit fp32 and bf16 rows straight into the growing bank: no per-batch D2H sync, no host
    copy, no concatenation, no second normalisation pass at ``fit``.  ``finish()`` returns the
    :class:`GalleryBank` (bit-identical to ``GalleryBank(torch.cat(batches), labels)``)."""

    def __init__(self, d: int, *, capacity: int = 1 << 16, device=None, idx_offset: int = 0):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.d = int(d)
        self.ld = self.lib.hcir_padded_dim(self.d)
        self.idx_offset = int(idx_offset)
        self.n = 0
        self._labels = []
        with torch.cuda.device(self.device):
            self._alloc(max(1, int(capacity)))
            self._dmax = torch.zeros((), dtype=torch.float32, device=self.device)

    def _alloc(self, cap: int):
        g32 = torch.empty((cap, self.ld), dtype=torch.float32, device=self.device)
        gbf = torch.empty((cap, self.ld), dtype=torch.bfloat16, device=self.device)
        if self.n:
            g32[: self.n].copy_(self.g32[: self.n])
            gbf[: self.n].copy_(self.gbf[: self.n])
        self.g32, self.gbf, self.capacity = g32, gbf, cap

    def append(self, features: torch.Tensor, labels=None):
        """features: [B, D] CUDA tensor (the encoder output of one batch); labels: [B] ints (any
        device / numpy / list), kept on the host like the reference keeps them."""
        if not isinstance(features, torch.Tensor) or not features.is_cuda or features.dim() != 2:
            raise ValueError("FeatureBankBuilder.append expects a 2-D CUDA tensor [batch, dim]")
        if features.shape[1] != self.d:
            raise ValueError(f"feature dim {features.shape[1]} != builder dim {self.d}")
        b = int(features.shape[0])
        if labels is not None:
            y = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
            if y.reshape(-1).shape[0] != b:
                raise ValueError(f"{y.reshape(-1).shape[0]} labels for {b} feature rows")
            self._labels.append(y.reshape(-1))
        elif self._labels:
            raise ValueError("labels were given for earlier batches but not for this one")
        if b == 0:
            return self
        with torch.cuda.device(self.device):
            x = features.detach()
            if x.device != self.device:
                x = x.to(self.device)
            if x.dtype != torch.float32:
                x = x.float()
            if x.stride(1) != 1:
                x = x.contiguous()
            if self.n + b > self.capacity:
                self._alloc(max(self.n + b, 2 * self.capacity))
            dl = torch.empty((b,), dtype=torch.float32, device=self.device)
            a = self.n
            _lib.check(self.lib.hcir_l2norm_cast(x.data_ptr(), b, self.d, x.stride(0), self.g32[a:a + b].data_ptr(),
                                                 self.gbf[a:a + b].data_ptr(), self.ld, dl.data_ptr(), _stream_ptr()),
                       "l2norm_cast")
            self._dmax = torch.maximum(self._dmax, dl.max())
            self.n += b
        return self

    def finish(self, classes=None) -> "GalleryBank":
        if self.n == 0:
            raise ValueError("FeatureBankBuilder.finish: no rows were appended")
        labels = np.concatenate(self._labels) if self._labels else None
        if labels is not None and labels.shape[0] != self.n:
            raise ValueError("some batches were appended without labels")
        return GalleryBank._from_parts(self.g32[: self.n], self.gbf[: self.n], float(self._dmax.item()), self.d,
                                       labels, classes, self.device, self.idx_offset)


def _gallery_from_parts(cls, g32, gbf, g_delta_max, d, labels, classes, device, idx_offset=0):
    """A GalleryBank over already normalised device rows (FeatureBankBuilder.finish)."""
    self = cls.__new__(cls)
    self.lib = _lib.load()
    self.device = device
    self.n, self.d = int(g32.shape[0]), int(d)
    self.ld = int(g32.shape[1])
    self.idx_offset = int(idx_offset)
    self.sm_count = _sm_count(device)
    self.g32, self.gbf, self.g_delta_max = g32, gbf, float(g_delta_max)
    self.eps_acc = float(self.ld) * 2.0 ** -22
    self.labels = None
    self.classes_ = None
    self._cls_dev = None
    if labels is not None:
        self.set_labels(labels, classes)
    self.last_stats, self.retry_stats = {}, {}
    self.kernel_events = None
    self.launches = 0
    self.k3_width = 0
    return self


GalleryBank._from_parts = classmethod(_gallery_from_parts)


# ---------------------------------------------------------------------------------------------
# functional surface
# ---------------------------------------------------------------------------------------------
def knn_topk(bank, query, k, *, normalized: bool = False, mode: str = "auto", use_graph: bool = False):
    """``torch.mm(query_n, bank_n.t()).topk(k)`` (qualitative_test.py:79-84;
    dual_view_model.py:317-335) without materialising the similarity matrix.
    ``bank`` is a [N, D] array/tensor or a prepared :class:`GalleryBank`.  ``normalized`` states
    that the rows are already unit vectors, as at the reference's call sites (F.normalize at
    qualitative_test.py:57,76); K1 normalises either way -- idempotent up to fp32 rounding -- because
    the bf16 bank and the certification bound are produced by the same pass."""
    del normalized
    gb = bank if isinstance(bank, GalleryBank) else GalleryBank(bank)
    return gb.topk(query, k, mode=mode, use_graph=use_graph)


def knn_predict(query, bank, labels, k, *, T=None, mode: str = "auto"):
    """kNN classification of ``query`` against (``bank``, ``labels``): uniform vote (T=None,
    == classification_engine.py:80-82) or temperature-weighted vote (T>0, extension)."""
    gb = bank if isinstance(bank, GalleryBank) else GalleryBank(bank, labels)
    if gb.labels is None:
        gb.set_labels(labels)
    return gb.predict(query, k, T=T, mode=mode)


# ---------------------------------------------------------------------------------------------
# pipelined submission: the host-side check of a step is deferred behind the next step's launch
# ---------------------------------------------------------------------------------------------
class PendingStep:
    """A submitted search step.  The device work is enqueued; the few bytes that decide whether
    the (rare) uncertified completion is needed travel to pinned host memory asynchronously and are
    looked at only in ``result()`` -- so a caller that keeps one or two steps in flight (submit the
    next, then take the previous result) never leaves the GPU idle on a host round trip.
    ``result()`` returns the step's exact results (device tensors owned by the caller)."""

    def __init__(self, event, flags_host, needs_redo, results, redo):
        self._event, self._flags, self._needs_redo, self._results, self._redo = event, flags_host, needs_redo, results, redo
        self.redone = False

    @property
    def ready_event(self):
        """The CUDA event recorded behind the step's results on the stream that produced them (None once
        ``result()`` has run, or for a step that completed synchronously)."""
        return self._event

    @property
    def device_results(self):
        """The results as enqueued (device tensors; final unless ``result()`` finds the batch must be
        redone) -- for callers that queue further device work, e.g. the read-back on a copy stream."""
        return self._results

    def result(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
            if self._needs_redo(self._flags):
                self._results = self._redo()
                self.redone = True
            self._redo = self._needs_redo = None
        return self._results


# ---------------------------------------------------------------------------------------------
# CUDA-graph session: one fixed-shape predict / top-k step replayed with a single launch
# ---------------------------------------------------------------------------------------------
class SearchSession:
    """A fixed-shape search step (``nq`` queries, ``k`` neighbours, optional vote) captured ONCE
    into a CUDA graph of FIVE kernels: K1(queries) -> K2 sample pass -> thresholds -> K2 main pass ->
    K3 with its fused tail (neighbour labels, vote, stores into the peer regions).  No fill / copy
    nodes: K3 maintains its own counters.  ``run`` copies the queries into the static input buffer,
    replays the graph (one launch instead of a dozen launches and their host-side allocation /
    planning work) and reads back the 4-byte count of uncertified queries; those (rare) are
    finished eagerly by the second tensor pass / the exact kernel.

    Only the tensor path is captured; ``GalleryBank.session`` returns None when the bank would use
    the exact CUDA-core path for this shape.  ``profile=True`` adds external CUDA events around the
    sample pass, the main pass and K3 so their durations can be read after every replay.

    Multi-GPU (sharded.py): ``tail_hook(session, tail)`` fills the peer fields of K3's tail (the
    channel is allocated collectively on the eager warm-up pass); ``post(session)`` enqueues what
    follows K3 in the same graph (the fused wait + merge + vote kernel, or the wait kernel)."""

    RETRY_CAPACITY = 128   # queries the in-graph completion pass holds (one query tile)

    def __init__(self, bank: "GalleryBank", nq: int, k: int, *, T=None, vote: bool = True, profile: bool = False,
                 pack: bool = False, post=None, tail_hook=None, trailer: bool = False, device_completion=None):
        if not (1 <= k <= bank.n):
            raise ValueError(f"k={k} must be in [1, N={bank.n}]")
        if vote and bank.labels is None:
            raise ValueError("this GalleryBank was built without labels")
        self.bank, self.nq, self.k, self.T, self.vote = bank, int(nq), int(k), T, vote
        # pack: results live in ONE byte block [idx | sims | labels] (hcir_packed_block_bytes), the
        # unit of the NCCL candidate all-gather; out_sim / out_idx / out_lab are views of it
        self.pack_results = bool(pack)
        self.post, self.post_out, self.tail_hook = post, None, tail_hook
        # trailer (NCCL exchange only): 16 bytes behind the packed block carry this rank's uncertified count
        self.trailer = 16 if (pack and trailer) else 0
        dev = bank.device
        self.events = {}
        self._profile = profile
        with torch.cuda.device(dev):
            self.q_in = torch.zeros((self.nq, bank.d), dtype=torch.float32, device=dev)
            # K3's counters: zero once, the kernel leaves [0] and [2] at zero after every launch.
            # [0:4] first pass, [4:8] the completion pass (device_completion)
            self.counters = torch.zeros((8,), dtype=torch.int32, device=dev)
            self.unc_state = self.counters[0:4]
            self.final_state = self.counters[4:8]
            # Device-driven completion: the second tensor pass over the uncertified queries is part of the
            # graph (three more kernels that return at once when nothing is uncertified), so a step whose
            # queries the second pass can certify needs no host action at all.  On by default from k = 128
            # (at k = 200 some query of every 16k-query step is uncertified); the ~10 us of idle launches are
            # not worth it for small k, where uncertified queries are rare and steps short.
            self.dev_completion = (k >= 128) if device_completion is None else bool(device_completion)
            # what the host reads after a replay: queries that are STILL uncertified
            self.unc_cnt = self.final_state[1:2] if self.dev_completion else self.unc_state[1:2]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body(capture=False)  # warm-up: caches allocator blocks, sets func attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            # a captured collective: NCCL's watchdog thread polls CUDA events while we capture, which
            # only the thread-local capture mode tolerates
            mode = "thread_local" if post is not None else "global"
            # Python's cyclic collector must not run inside the capture: an earlier session is garbage only
            # through the bank <-> session cycle, and destroying its CUDA graph (cudaGraphExecDestroy) while
            # this thread captures in global mode invalidates the capture (seen as error 901 on the next
            # launch).  torch.cuda.graph collects before it begins; nothing may be collected until it ends.
            import gc
            gc_was = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(self.graph, capture_error_mode=mode):
                    self._body(capture=True)
            finally:
                if gc_was:
                    gc.enable()
        self.launches_per_run = 1  # one graph launch; the kernels inside: self.kernels_per_run

    def _mark(self, name, capture):
        if not (capture and self._profile):
            return
        ev = torch.cuda.Event(enable_timing=True, external=True)
        ev.record()
        self.events[name] = ev

    def _body(self, capture: bool):
        b, lib, dev = self.bank, self.bank.lib, self.bank.device
        nq, k = self.nq, self.k
        st = _stream_ptr()
        self.q32, self.qbf, self.qdl = l2_normalize(self.q_in, pad_rows_to=128)
        kc = b.choose_kc(k)
        plan = Plan()
        _lib.check(lib.hcir_simtopk_plan(nq, b.n, b.ld, kc, b.sm_count, plan), "simtopk_plan")
        plan.q_rows = -(-nq // 128) * 128
        self.plan = plan
        self.ws = torch.empty((int(plan.bytes),), dtype=torch.uint8, device=dev)
        self.out_lab = None
        with_lab = b.labels is not None
        if self.pack_results:
            e = nq * k
            self.block_bytes = int(lib.hcir_packed_block_bytes(nq, k, int(with_lab)))
            self.pack = torch.zeros((self.block_bytes + self.trailer,), dtype=torch.uint8, device=dev)
            self.out_idx = self.pack[: e * 8].view(torch.int64).view(nq, k)
            self.out_sim = self.pack[e * 8: e * 12].view(torch.float32).view(nq, k)
            if with_lab:
                self.out_lab = self.pack[e * 12: e * 16].view(torch.int32).view(nq, k)
        else:
            self.out_sim = torch.empty((nq, k), dtype=torch.float32, device=dev)
            self.out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        self.unc_list = torch.empty((nq,), dtype=torch.int32, device=dev)
        kernels = 1  # l2norm
        self._mark("t0", capture)
        if plan.sample_rows > 0:
            plan.flags = 2
            _lib.check(lib.hcir_simtopk(self.qbf.data_ptr(), nq, b.gbf.data_ptr(), b.n, b.ld, plan,
                                        self.ws.data_ptr(), st), "simtopk(sample)")
            self._mark("t1", capture)
            plan.flags = 4
            kernels += 2
        # measurement aid: HCIR_MAIN_FLAGS=16 runs the main pass on CTA pairs (cta_group::2), 32 rotates tiles
        plan.flags |= int(os.environ.get("HCIR_MAIN_FLAGS", "0")) & (16 | 32)
        _lib.check(lib.hcir_simtopk(self.qbf.data_ptr(), nq, b.gbf.data_ptr(), b.n, b.ld, plan,
                                    self.ws.data_ptr(), st), "simtopk(main)")
        plan.flags = b.k3_width << 8
        self._mark("t2", capture)
        # K3 + tail: labels / vote / peers happen inside the query's own CTA
        tail = Tail()
        self.pred = None
        if with_lab and (self.vote or self.out_lab is not None):
            tail.labels, tail.n_labels = b.labels.data_ptr(), b.n
            tail.num_classes = len(b.classes_)
        if self.vote:
            self.pred = torch.empty((nq,), dtype=torch.int64, device=dev)
            tail.pred = self.pred.data_ptr()
            tail.T = float(self.T) if self.T is not None else 0.0
            tail.classes = b._classes_device().data_ptr()
        if self.out_lab is not None:
            tail.out_lab = self.out_lab.data_ptr()
        if self.tail_hook is not None:
            self.tail_hook(self, tail)
        if self.dev_completion and self._completion_plan() is None:
            self.dev_completion = False           # (no sample-pass layout for the batch: cannot happen on
            self.unc_cnt = self.unc_state[1:2]    #  the tensor path, but never leave the peers unsignalled)
        if self.dev_completion and tail.world > 0:
            tail.no_signal = 1   # the completion pass may still correct rows: IT signals the peers
        _lib.check(lib.hcir_select_rescore(self.q32.data_ptr(), b.g32.data_ptr(), b.ld, nq, b.n, k, b.idx_offset,
                                           plan, self.ws.data_ptr(), self.qdl.data_ptr(), b.g_delta_max, b.eps_acc,
                                           self.out_sim.data_ptr(), self.out_idx.data_ptr(),
                                           self.unc_list.data_ptr(), self.unc_state.data_ptr(), tail, st),
                   "select_rescore")
        plan.flags = 0
        self._mark("t3", capture)
        kernels += 2
        self.final_list = self.unc_list
        if self.dev_completion:
            kernels += self._completion_pass(tail, st)
        if self.trailer:
            self.pack[self.block_bytes: self.block_bytes + 4].view(torch.int32).copy_(self.unc_cnt)
        if self.post is not None:
            self.post_out = self.post(self)
        self.kernels_per_run = kernels

    def _completion_plan(self):
        """Launch plan of the completion batch (RETRY_CAPACITY queries, up to 4x kc), or None."""
        b = self.bank
        kc = b.choose_kc(self.k)
        plan2 = Plan()
        for mult in (4, 3, 2, 1):   # widest kc whose plan still carries per-query thresholds
            kc2 = max(kc, min(mult * kc, 2048))
            _lib.check(b.lib.hcir_simtopk_plan(self.RETRY_CAPACITY, b.n, b.ld, kc2, b.sm_count, plan2), "simtopk_plan")
            if plan2.sample_rows > 0:
                return plan2
        return None

    def _completion_pass(self, tail, st) -> int:
        """In-graph second pass over the queries K3 could not certify (GalleryBank._finish_uncertified has
        the reasoning): gather them into a compact batch with explicit thresholds (hcir_retry_setup), stream
        the gallery once for that batch if there is one (hcir_simtopk_gated), and let K3 -- same tail, rows
        mapped back to the original queries -- overwrite what it certifies, signal the peers, and list what
        is STILL uncertified (near-duplicate galleries) for the exact kernel on the host's initiative."""
        b, lib, dev = self.bank, self.bank.lib, self.bank.device
        R, k = self.RETRY_CAPACITY, self.k
        plan2 = self._completion_plan()
        plan2.q_rows = R
        self.plan2 = plan2
        if getattr(self, "ws2", None) is None:
            # allocated once, on the eager warm-up pass (nothing here needs initialising: the setup kernel
            # writes every threshold, rows beyond the live count get +inf and are never read back) -- a
            # torch.zeros inside the capture would put a fill node of the whole workspace into the graph
            self.ws2 = torch.empty((int(plan2.bytes),), dtype=torch.uint8, device=dev)
            self.q2bf = torch.empty((R, b.ld), dtype=torch.bfloat16, device=dev)
            self.q2f = torch.empty((R, b.ld), dtype=torch.float32, device=dev)
            self.q2d = torch.empty((R,), dtype=torch.float32, device=dev)
            self.qmap = torch.empty((R,), dtype=torch.int32, device=dev)
            self.retry_active = torch.zeros((1,), dtype=torch.int32, device=dev)
            self._final_list = torch.empty((self.nq,), dtype=torch.int32, device=dev)
        self.final_list = self._final_list
        _lib.check(lib.hcir_retry_setup(self.qbf.data_ptr(), self.q32.data_ptr(), self.qdl.data_ptr(), b.ld, self.nq, k,
                                        self.out_sim.data_ptr(), self.unc_list.data_ptr(), self.unc_state.data_ptr(),
                                        b.g_delta_max, b.eps_acc, R, self.q2bf.data_ptr(), self.q2f.data_ptr(),
                                        self.q2d.data_ptr(), self.ws2.data_ptr() + int(plan2.thr0_off),
                                        self.ws2.data_ptr() + int(plan2.thr_hi_off), self.qmap.data_ptr(),
                                        self.retry_active.data_ptr(), self.final_list.data_ptr(),
                                        self.final_state.data_ptr(), st), "retry_setup")
        plan2.flags = 4  # HCIR_FLAG_MAIN_ONLY: the thresholds are in the workspace
        _lib.check(lib.hcir_simtopk_gated(self.q2bf.data_ptr(), R, b.gbf.data_ptr(), b.n, b.ld, plan2,
                                          self.ws2.data_ptr(), self.retry_active.data_ptr(), st), "simtopk(completion)")
        plan2.flags = 0
        t2 = Tail()
        for name, _ in Tail._fields_:   # same labels / vote / peers as the first pass ...
            setattr(t2, name, getattr(tail, name))
        t2.no_signal = 0                # ... but this launch signals, commits only what it certifies,
        t2.commit_certified_only = 1    # and writes to the rows of the original queries
        t2.qmap = self.qmap.data_ptr()
        t2.active = self.retry_active.data_ptr()
        t2.out_rows = self.nq
        _lib.check(lib.hcir_select_rescore(self.q2f.data_ptr(), b.g32.data_ptr(), b.ld, R, b.n, k, b.idx_offset, plan2,
                                           self.ws2.data_ptr(), self.q2d.data_ptr(), b.g_delta_max, b.eps_acc,
                                           self.out_sim.data_ptr(), self.out_idx.data_ptr(),
                                           self.final_list.data_ptr(), self.final_state.data_ptr(), t2, st),
                   "select_rescore(completion)")
        return 3

    def _gather_packed_labels(self):
        b = self.bank
        b.launches += 1
        _lib.check(b.lib.hcir_gather_labels(self.out_idx.data_ptr(), self.out_idx.numel(), b.labels.data_ptr(), b.n,
                                            b.idx_offset, self.out_lab.data_ptr(), _stream_ptr()), "gather_labels")

    def _tail(self):
        return self.bank.vote_from_idx(self.out_sim, self.out_idx, T=self.T, nbr_out=self.out_lab)

    def kernel_ms(self):
        """Durations (ms) of the profiled phases of the LAST replay (profile=True only)."""
        e = self.events
        if not e:
            return {}
        out = {"simtopk": e["t1" if "t1" in e else "t0"].elapsed_time(e["t2"]), "select_rescore": e["t2"].elapsed_time(e["t3"])}
        if "t1" in e:
            out["simtopk_sample"] = e["t0"].elapsed_time(e["t1"])
        return out

    def finish_uncertified(self, n_unc: int):
        """Eager completion of the (rare) queries the tensor path could not certify: the second tensor pass
        and then the exact kernel -- or, when the second pass already ran inside the graph
        (device_completion), only the exact kernel for what it left."""
        b = self.bank
        if self.dev_completion:
            b._exact(self.q32, self.final_list, n_unc, self.k, self.out_sim, self.out_idx)
            b.retry_stats = {"second_pass": "in graph", "exact": int(n_unc)}
        else:
            b._finish_uncertified(self.q32, self.qbf, self.qdl, self.unc_list, n_unc, self.k, self.out_sim,
                                  self.out_idx)
        if self.out_lab is not None:
            self._gather_packed_labels()

    @property
    def input(self) -> torch.Tensor:
        """The step's static input buffer [nq, d] fp32 on the device.  A producer that writes its queries
        straight into it (the encoder's output, an H2D copy) calls ``run()`` / ``submit()`` without an
        argument and saves the device-to-device copy in front of every replay."""
        return self.q_in

    def submit(self, queries=None) -> PendingStep:
        """Pipelined ``run``: enqueue the step and return at once.  ``result()`` of the returned
        handle gives (pred | None, sims, idx) as fresh device tensors; uncertified queries (rare) are
        completed there by re-running this batch through the eager path.  ``queries`` must stay
        unmodified until then.  At most 8 steps may be pending per session."""
        b = self.bank
        if queries is not None and tuple(queries.shape) != (self.nq, b.d):
            raise ValueError(f"session was built for queries of shape {(self.nq, b.d)}, got {tuple(queries.shape)}")
        with torch.cuda.device(b.device):
            ring = self.__dict__.get("_flag_ring")
            if ring is None:
                ring = self._flag_ring = torch.zeros((8, 1), dtype=torch.int32, pin_memory=True)
                self._submitted = 0
            slot = ring[self._submitted % 8]
            self._submitted += 1
            if queries is not None:
                self.q_in.copy_(queries, non_blocking=True)
            else:
                queries = self.q_in.clone()   # the redo path needs the batch after q_in has been overwritten
            self.graph.replay()
            b.launches += self.kernels_per_run
            res = (self.pred.clone() if self.pred is not None else None, self.out_sim.clone(), self.out_idx.clone())
            slot.copy_(self.unc_cnt, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()

        def redo():
            with torch.cuda.device(b.device):
                q = queries if queries.is_cuda else queries.to(b.device)
                if self.vote:
                    return b.predict(q, self.k, T=self.T, mode="tensor", return_neighbors=True)
                s_, i_ = b._topk_device(q, self.k, "tensor")
                return None, s_, i_

        return PendingStep(ev, slot, lambda f: int(f[0]) > 0, res, redo)

    def run(self, queries=None, check: bool = True):
        """queries: [nq, d] fp32 tensor (device, or host -- pinned for an async copy), or None when the
        caller has filled ``self.input`` itself.  Returns
        (pred [nq] int64 | None, sims [nq, k], idx [nq, k]) as DEVICE tensors owned by the session
        (valid until the next run).  ``check=False``: only replay; the caller reads the uncertified
        count (multi-GPU: from the gathered trailers) and calls ``finish_uncertified``."""
        b = self.bank
        if queries is not None and tuple(queries.shape) != (self.nq, b.d):
            raise ValueError(f"session was built for queries of shape {(self.nq, b.d)}, got {tuple(queries.shape)}")
        with torch.cuda.device(b.device):
            if queries is not None:
                self.q_in.copy_(queries, non_blocking=True)
            self.graph.replay()
            b.launches += self.kernels_per_run
            if not check:
                return self.pred, self.out_sim, self.out_idx
            cnt = self.counters.tolist()   # one 32-byte read-back: [1] first pass, [5] after the in-graph completion
            n_unc = cnt[5] if self.dev_completion else cnt[1]
            pred = self.pred
            if n_unc > 0:
                self.finish_uncertified(n_unc)
                if self.vote:
                    pred = self._tail()
            b.last_stats = {"path": "tensor+graph", "uncertified": n_unc, "uncertified_first_pass": cnt[1],
                            "completion": "device" if self.dev_completion else "host",
                            "nsplit": int(self.plan.nsplit),
                            "kc": int(self.plan.kc), "cap": int(self.plan.cap),
                            "workspace_bytes": int(self.plan.bytes), "sample_rows": int(self.plan.sample_rows),
                            "chunk_w": int(self.plan.chunk_w)}
        return pred, self.out_sim, self.out_idx


def _bank_session(self, nq: int, k: int, *, T=None, vote: bool = True, profile: bool = False, pack: bool = False,
                  post=None, post_key=None, tail_hook=None, trailer: bool = False, device_completion=None):
    """Cached :class:`SearchSession` for this shape, or None if the exact path would be used."""
    if nq < 1 or not self.use_tensor_path(nq, k):
        return None
    key = (int(nq), int(k), None if T is None else float(T), bool(vote), bool(profile), bool(pack), post_key,
           device_completion)
    # Sessions whose graph ends in a multi-GPU exchange own a collectively constructed channel: they
    # live in their own cache, whose keys hold rank-independent values only, so every rank creates
    # and evicts them at the same calls (rank-local sessions can never shift that order).
    cache = self.__dict__.setdefault("_sessions_collective" if post is not None else "_sessions", {})
    s = cache.get(key)
    if s is None:
        if len(cache) >= 8:  # each session owns a workspace: keep a handful
            old = cache.pop(next(iter(cache)))
            xc = getattr(old, "xchg", None)
            if xc is not None:   # every rank evicts the same session at the same call -> collective close
                xc.close()
        s = cache[key] = SearchSession(self, nq, k, T=T, vote=vote, profile=profile, pack=pack, post=post,
                                       tail_hook=tail_hook, trailer=trailer, device_completion=device_completion)
    return s


def _bank_drop_sessions(self):
    """Forget every cached SearchSession (their CUDA graphs hold raw pointers into this bank's
    tensors).  COLLECTIVE when sessions with a multi-GPU exchange exist: all ranks must call it."""
    self.__dict__.pop("_sessions", None)
    for sess in self.__dict__.pop("_sessions_collective", {}).values():
        xc = getattr(sess, "xchg", None)
        if xc is not None:
            xc.close()


GalleryBank.drop_sessions = _bank_drop_sessions
GalleryBank.session = _bank_session
