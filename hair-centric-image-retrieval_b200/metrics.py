"""Consumers of the [Q, k] neighbour lists.

* ``recall_ap_at_k``  -- Recall@K / AP@K exactly as the per-query Python loop of
  experiments/DualViewHair/scripts/quantitative_eval.py:195-209 computes them, vectorised over
  the batch on the device the indices live on (pure index bookkeeping: torch is plumbing here).
* ``kth_neighbour``   -- ``NegSamplerStatic`` (HairPretraining/src/neg_sampling.py:26-53, cosine
  branch): index of the k-th most similar row of the batch to every row, the row itself being
  rank 1.  Same kernels as retrieval with gallery = queries = the batch.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import GalleryBank


def recall_ap_at_k(neighbour_idx, ground_truth, ks=(10, 20, 50)):
    """``neighbour_idx`` [Q, >=max(ks)] int; ``ground_truth`` list of Q index collections.
    Returns {"Recall": {k: r}, "mAP": {k: m}, "total_queries": Q} (quantitative_eval.py:228-232):
      Recall@K = fraction of queries with any ground-truth item in the top K;
      AP@K = sum_i [hit_i] * (hits so far / (i+1)) / min(|gt|, K), 0 when gt is empty."""
    idx = torch.as_tensor(neighbour_idx)
    q = idx.shape[0]
    if len(ground_truth) != q:
        raise ValueError("one ground-truth set per query")
    kmax = max(ks)
    if idx.shape[1] < kmax:
        raise ValueError(f"need at least {kmax} neighbours per query")
    dev = idx.device
    gmax = max((len(g) for g in ground_truth), default=0)
    gt = torch.full((q, max(gmax, 1)), -1, dtype=torch.int64, device=dev)
    glen = torch.zeros(q, dtype=torch.int64, device=dev)
    for r, g in enumerate(ground_truth):
        g = list(g)
        glen[r] = len(g)
        if g:
            gt[r, : len(g)] = torch.as_tensor(g, dtype=torch.int64, device=dev)
    hit = (idx[:, :kmax, None].long() == gt[:, None, :]).any(dim=2)          # [Q, kmax]
    cum = torch.cumsum(hit.to(torch.float64), dim=1)
    prec = cum / torch.arange(1, kmax + 1, device=dev, dtype=torch.float64)  # hits / (i+1)
    out = {"Recall": {}, "mAP": {}, "total_queries": q}
    for k in ks:
        h = hit[:, :k]
        out["Recall"][k] = float(h.any(dim=1).double().mean()) if q else 0.0
        denom = torch.minimum(glen, torch.tensor(k, device=dev)).clamp(min=1).double()
        ap = (prec[:, :k] * h).sum(dim=1) / denom
        ap = torch.where(glen > 0, ap, torch.zeros_like(ap))
        out["mAP"][k] = float(ap.mean()) if q else 0.0
    return out


def kth_neighbour(embeddings, k: int = 7, *, metric: str = "cosine", mode: str = "auto"):
    """``NegSamplerStatic`` (neg_sampling.py:26-53): ``sorted_indices[:, k-1]`` of the row-wise
    descending sort of the batch similarity matrix -- cosine similarity of the rows (:34-37) or negative
    euclidean distance of the RAW rows (:38-41).  Returns int64 [B] (same device kind as the input).
    Exact ties are ordered by ascending index (torch.sort leaves them unspecified).

    euclidean: for a fixed row a, ascending ||a - b|| == descending a.b - ||b||^2 / 2, an inner product
    of the augmented rows [a, 1] . [b, -||b||^2 / 2]; the exact fp32 CUDA-core kernel ranks by that
    inner product directly (it normalises nothing), so no second code path is needed."""
    x = embeddings if isinstance(embeddings, torch.Tensor) else torch.as_tensor(np.asarray(embeddings))
    b = x.shape[0]
    if k < 1 or k > b:
        raise ValueError(f"k must be between 1 and {b}")
    if metric == "cosine":
        bank = GalleryBank(x.detach().float())
        _, idx = bank.topk(x.detach().float(), k, mode=mode)
        return idx[:, k - 1]
    if metric != "euclidean":
        raise ValueError("Unsupported metric. Choose 'cosine' or 'euclidean'.")
    from . import _lib
    from .engine import _require_cuda, _sm_count, _stream_ptr
    lib = _lib.load()
    dev = x.device if x.is_cuda else _require_cuda(None)
    with torch.cuda.device(dev):
        xd = x.detach().to(dev, torch.float32)
        d = xd.shape[1]
        ld = lib.hcir_padded_dim(d + 1)
        qa = torch.zeros((b, ld), dtype=torch.float32, device=dev)
        ga = torch.zeros((b, ld), dtype=torch.float32, device=dev)
        qa[:, :d], ga[:, :d] = xd, xd
        qa[:, d] = 1.0
        ga[:, d] = -0.5 * (xd.double() ** 2).sum(dim=1).float()
        o_s = torch.empty((b, k), dtype=torch.float32, device=dev)
        o_i = torch.empty((b, k), dtype=torch.int64, device=dev)
        sms = _sm_count(dev)
        nbytes = int(lib.hcir_exact_workspace_bytes(b, b, k, sms))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.hcir_exact_topk(qa.data_ptr(), ga.data_ptr(), ld, b, k, 0, None, b, o_s.data_ptr(),
                                       o_i.data_ptr(), ws.data_ptr(), nbytes, sms, _stream_ptr()), "exact_topk")
        out = o_i[:, k - 1]
        return out if x.is_cuda else out.cpu()
