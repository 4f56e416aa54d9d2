"""Gallery on-disk format and result wire format of the reference.

* ``embeddings.npy`` ([N, D] fp32, NOT normalised) + ``image_paths.txt`` (one path per line), as
  written / read by ``HairEncoder.extract_dataset_features`` / ``load_embeddings``
  (src/models/hair_encoder.py:136-139,154-157).  ``load_gallery`` memory-maps the .npy and streams it
  to the device in row chunks through K1 (normalise + bf16 cast), so a 10M-row file never needs
  a second host copy; ``rows=(start, stop)`` loads one shard of it.
* The Visualizer's top-100 JSON ``[{"query": "<id>_hair.png", "top100": [names...]}, ...]`` as dumped
  by experiments/DualViewHair/scripts/quantitative_eval.py:189-192,214-217 and consumed by
  Visualizer/app/models/data_loader.py:15-27.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .engine import GalleryBank


def save_embeddings(save_dir: str, embeddings, paths) -> None:
    """hair_encoder.py:134-139."""
    os.makedirs(save_dir, exist_ok=True)
    np.save(os.path.join(save_dir, "embeddings.npy"), np.asarray(embeddings, dtype=np.float32))
    with open(os.path.join(save_dir, "image_paths.txt"), "w") as f:
        for p in paths:
            f.write(p + "\n")


def load_paths(save_dir: str):
    with open(os.path.join(save_dir, "image_paths.txt"), "r") as f:
        return [line.strip() for line in f.readlines()]


def load_gallery(save_dir: str, *, device=None, rows=None, labels=None, chunk_rows: int = 1 << 18):
    """(GalleryBank, paths) from ``embeddings.npy`` + ``image_paths.txt``.  ``rows=(start, stop)``
    builds the bank of one contiguous shard with global indices (idx_offset = start)."""
    emb = np.load(os.path.join(save_dir, "embeddings.npy"), mmap_mode="r")
    if emb.ndim != 2:
        raise ValueError(f"embeddings.npy must be [N, D], got {emb.shape}")
    paths = load_paths(save_dir)
    if len(paths) != emb.shape[0]:
        raise ValueError(f"{len(paths)} paths for {emb.shape[0]} embeddings")
    start, stop = (0, emb.shape[0]) if rows is None else rows
    bank = GalleryBank(emb[start:stop], labels, device=device, idx_offset=start, chunk_rows=chunk_rows)
    return bank, paths


def top100_records(query_names, neighbour_idx, all_paths, width: int = 100):
    """The Visualizer records for a batch of queries: basename of the query and of its neighbours
    (quantitative_eval.py:189-192 keeps ``retrieved[:100]`` of whatever depth was searched)."""
    out = []
    for name, row in zip(query_names, np.asarray(neighbour_idx)):
        out.append({"query": os.path.basename(name),
                    "top100": [os.path.basename(all_paths[int(i)]) for i in row[:width]]})
    return out


def write_top100_json(path: str, records) -> None:
    """quantitative_eval.py:214-217 (``json.dump(..., indent=2)``)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(list(records), f, indent=2)


def read_top100_json(path: str):
    """Visualizer/app/models/data_loader.py:15-27: {query: top100 list}."""
    with open(path, "r") as f:
        data = json.load(f)
    return {rec["query"]: rec["top100"] for rec in data}
