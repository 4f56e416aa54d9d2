"""Retrieval-shaped drop-ins.

* ``retrieve_similar_images``  == HairEncoder.retrieve_similar_images
  (src/models/hair_encoder.py:180-198; twin src/models/face_encoder.py:197-215)
* ``HairRetrievalB200``        == HairRetrieval._build_gallery / retrieve_similar
  (experiments/DualViewHair/scripts/qualitative_test.py:43-103)
* ``compute_similarity_topk``  == HairstyleRetrievalModel.compute_similarity followed by
  topk (experiments/DualViewHair/src/models/dual_view_model.py:317-335) without the matrix
* ``FlatIndex``                == faiss.normalize_L2 + IndexFlatL2.add/search call sites
  (HairPretraining/app/inference.py:75,90-96,108; quantitative_eval.py:143-147,185)
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import GalleryBank, _as_2d_f32

_BANK_CACHE: dict = {}


def clear_bank_cache():
    _BANK_CACHE.clear()


def _content_digest(a) -> tuple:
    """Digest of the FULL content of an embeddings array: shape + 64-bit sum and xor of its fp32 bit
    patterns (CUDA tensors: sum of the bit patterns and of the squares).  One streaming pass (numpy:
    ~0.2 s per GB on the host; CUDA tensors: microseconds) --
    still far cheaper than what the reference does on every call (re-normalise all N rows, full
    argsort: hair_encoder.py:193-194), and unlike a sampled probe it cannot miss an in-place edit."""
    if isinstance(a, torch.Tensor) and a.is_cuda:
        t = a.detach()
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        s1 = int(t.view(torch.int32).sum(dtype=torch.int64).item())          # sum of the bit patterns
        s2 = float((t.double() if t.numel() <= (1 << 24) else t).square().sum(dtype=torch.float64).item())
        return ("t", tuple(t.shape), s1, s2)
    if isinstance(a, torch.Tensor):
        a = a.detach().numpy()
    arr = np.ascontiguousarray(a, dtype=np.float32)
    bits = arr.view(np.uint32).reshape(-1)
    return ("n", arr.shape, int(bits.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(bits)) if bits.size else 0)


def _cached_bank(all_embeddings) -> GalleryBank:
    """The reference re-normalises all N rows on every call (hair_encoder.py:193).  Here the
    normalised bank is built once per embeddings CONTENT and reused: the cache key is a digest of
    every element (``_content_digest``), so an array edited in place -- anywhere -- is rebuilt and a
    stale bank can never answer.  Callers that own the lifetime pass a prepared :class:`GalleryBank`
    instead (no digest pass at all); ``clear_bank_cache()`` drops everything."""
    if isinstance(all_embeddings, GalleryBank):
        return all_embeddings
    a = all_embeddings if isinstance(all_embeddings, torch.Tensor) else np.asarray(all_embeddings)
    key = _content_digest(a)
    bank = _BANK_CACHE.get(key)
    if bank is None:
        if len(_BANK_CACHE) >= 4:
            _BANK_CACHE.pop(next(iter(_BANK_CACHE)))
        bank = GalleryBank(a)
        _BANK_CACHE[key] = bank
    return bank


def retrieve_similar_images(query_embedding, all_embeddings, all_paths, top_k=5):
    """Drop-in for hair_encoder.py:180-198: un-normalised numpy in, list of
    ``{'path': str, 'similarity': np.float32}`` out, descending similarity."""
    bank = _cached_bank(all_embeddings)
    sims, idx = bank.topk(np.asarray(query_embedding, dtype=np.float32).reshape(1, -1), int(top_k))
    return [{"path": all_paths[int(i)], "similarity": np.float32(s)} for i, s in zip(idx[0], sims[0])]


def compute_similarity_topk(query_embeddings, gallery_embeddings, k):
    """``compute_similarity(q, g).topk(k)`` (dual_view_model.py:317-335) without [N_q, N_g]."""
    return _cached_bank(gallery_embeddings).topk(query_embeddings, int(k))


class HairRetrievalB200:
    """qualitative_test.py:22-103 with the gallery held on the GPU.  ``gallery_embeddings``
    is what ``_build_gallery`` concatenates (one embedding per dataset item)."""

    def __init__(self, gallery_embeddings, gallery_ids=None):
        self.bank = GalleryBank(gallery_embeddings)
        self.gallery_embeddings, _ = _as_2d_f32(gallery_embeddings, "gallery_embeddings")
        self.gallery_ids = list(gallery_ids) if gallery_ids is not None else list(range(self.bank.n))

    def retrieve_similar(self, query_idx: int, top_k: int = 10):
        """top-(k+1), drop the first hit (self), like qualitative_test.py:82-84."""
        q = self.gallery_embeddings[query_idx: query_idx + 1]
        sims, idx = self.bank.topk(q, top_k + 1)
        sims = sims[0][1:]
        idx = idx[0][1:]
        results = [{"gallery_idx": int(i), "image_id": self.gallery_ids[int(i)], "similarity": float(s)}
                   for i, s in zip(idx, sims)]
        return {"query_idx": query_idx, "query_id": self.gallery_ids[query_idx], "results": results}


class FlatIndex:
    """faiss.IndexFlatL2 over L2-normalised vectors, as the reference uses it
    (``faiss.normalize_L2(x); index.add(x); D, I = index.search(q, k)``).  Rows are
    normalised on add/search (idempotent for already-normalised input); ``search`` returns
    faiss's convention: squared L2 = 2 - 2*cos ascending, int64 indices."""

    def __init__(self, d: int):
        self.d = int(d)
        self._chunks: list[np.ndarray] = []
        self._bank: GalleryBank | None = None

    @property
    def ntotal(self) -> int:
        return sum(c.shape[0] for c in self._chunks)

    def add(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).reshape(-1, self.d)
        self._chunks.append(x)
        self._bank = None

    def search(self, q, k: int):
        if self._bank is None:
            if not self._chunks:
                raise RuntimeError("FlatIndex is empty")
            self._bank = GalleryBank(np.concatenate(self._chunks, axis=0))
        sims, idx = self._bank.topk(np.asarray(q, dtype=np.float32).reshape(-1, self.d), int(k))
        return (2.0 - 2.0 * sims).astype(np.float32), idx
