"""Seeded synthetic embeddings of the benchmark shapes.

bank = randn(N, D) + 2.0 * centroid[label], ``C`` random unit centroids, so neighbours are
label-correlated (pure iid Gaussians make the vote meaningless).  Rows are returned
UN-normalised, like the encoder output the reference normalises at
HairPretraining/src/classification_engine.py:50.  There is no network for datasets or
checkpoints, so every benchmark and parity case uses these generators.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CENTROID_SCALE = 2.0

# name -> (N, D, Q, k, C)  (BASELINE.json configs / SURVEY.md section 8d)
CONFIGS = {
    "C1": dict(n=10_000, d=512, q=1_000, k=20, classes=27, T=0.07),
    "C2": dict(n=200_000, d=768, q=10_000, k=20, classes=61, T=0.07),
    "C3": dict(n=1_000_000, d=768, q=4_096, k=100, classes=61, T=None),
    "C4": dict(n=10_000_000, d=768, q=64, k=20, classes=61, T=None),
    "C5": dict(n=10_000_000, d=2048, q=16_384, k=200, classes=61, T=None),
}


def centroids(n_classes: int, d: int, seed: int, device="cpu") -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return F.normalize(torch.randn(n_classes, d, generator=g), dim=1).to(device)


def make_clustered(n: int, d: int, n_classes: int, seed: int, *, centroid_seed: int = 7,
                   labels: torch.Tensor | None = None, device="cpu", chunk: int = 1 << 18):
    """Return (features [n,d] fp32 un-normalised, labels [n] int64 in [0, n_classes)).

    CPU generation is bit-reproducible for a given torch version (used by the golden
    fixtures); device generation (``device='cuda'``) is seeded per chunk and used only for
    the large benchmark banks that would not fit through host memory quickly."""
    dev = torch.device(device)
    cen = centroids(n_classes, d, centroid_seed, dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    if labels is None:
        labels = torch.randint(0, n_classes, (n,), generator=g, device=dev)
    else:
        labels = labels.to(dev)
    out = torch.empty(n, d, dtype=torch.float32, device=dev)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        out[a:b] = torch.randn(b - a, d, generator=g, device=dev)
        out[a:b] += CENTROID_SCALE * cen[labels[a:b]]
    return out, labels


def make_config(name: str, *, device="cpu", n=None, q=None):
    """(bank, bank_labels, queries, query_labels, cfg) for a named config; ``n``/``q`` override
    the sizes (parity tests run reduced sizes of the same recipe)."""
    cfg = dict(CONFIGS[name])
    if n is not None:
        cfg["n"] = n
    if q is not None:
        cfg["q"] = q
    tag = int(name[1:])
    bank, bl = make_clustered(cfg["n"], cfg["d"], cfg["classes"], 1234 + tag, device=device)
    qs, ql = make_clustered(cfg["q"], cfg["d"], cfg["classes"], 4321 + tag, device=device)
    return bank, bl, qs, ql, cfg
