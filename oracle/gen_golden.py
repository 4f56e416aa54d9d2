"""Generate the committed golden fixtures under tests/golden/ (TEST INFRASTRUCTURE).

Run in the build container (``python oracle/gen_golden.py``).  It executes the reference's
own call sequences -- the installed sklearn / torch / numpy code the reference calls, via
``oracle.oracle`` -- on seeded synthetic inputs and stores the OUTPUTS (plus a checksum of
the inputs, and the inputs themselves for the tiny case).  /root/reference is read only for
the real 27-class label column of HairPretraining/data/data_*_combination3.csv (labels are
data, not code); nothing at test time reads /root/reference.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "hcir_synth", os.path.join(ROOT, "hair-centric-image-retrieval_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

OUT = os.path.join(ROOT, "tests", "golden")


def digest(t) -> str:
    a = np.ascontiguousarray(t.numpy() if isinstance(t, torch.Tensor) else t)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def gen_tiny():
    """Fully self-contained case: inputs stored.  Non-contiguous labels, one zero row, one
    exact duplicate pair (exact tie), N not a multiple of any tile size."""
    g = torch.Generator().manual_seed(99)
    n, d, q, k = 301, 64, 17, 5
    classes = np.array([2, 4, 5, 7, 9, 10, 11, 13, 52])
    feats = torch.randn(n, d, generator=g)
    lab_i = torch.randint(0, len(classes), (n,), generator=g).numpy()
    feats += 1.5 * torch.randn(len(classes), d, generator=g)[lab_i]
    feats[17] = 0.0                      # zero row
    feats[200] = feats[100]              # exact duplicate -> exact tie
    labels = classes[lab_i]
    queries = torch.randn(q, d, generator=g) + 1.5 * torch.randn(len(classes), d, generator=g)[
        torch.randint(0, len(classes), (q,), generator=g)]
    queries[3] = feats[100] * 3.0        # query collinear with the duplicated pair
    bn, qn = O.normalize(feats), O.normalize(queries)
    pred, dist, ind = O.sklearn_knn(bn.numpy(), labels, qn.numpy(), k)
    v, i = O.mm_topk(qn, bn, k)
    cs, ci = O.canonical_topk(qn, bn, k)
    a_idx, a_sim = [], []
    for r in range(q):                   # hair_encoder.py path: UN-normalised inputs
        ii, ss = O.cosine_argsort(queries[r].numpy(), feats.numpy(), k)
        a_idx.append(ii)
        a_sim.append(ss)
    tpred, tscores = O.vote_temperature(v.numpy(), labels[i.numpy()], classes, 0.07)
    np.savez_compressed(
        os.path.join(OUT, "tiny.npz"), feats=feats.numpy(), labels=labels, queries=queries.numpy(),
        classes=classes, k=k, bank_unit=bn.numpy(), queries_unit=qn.numpy(),
        sk_pred=pred, sk_dist=dist, sk_ind=ind, mm_sims=v.numpy(), mm_idx=i.numpy(),
        canon_sims=cs, canon_idx=ci, argsort_idx=np.stack(a_idx), argsort_sims=np.stack(a_sim),
        temp_pred=tpred, temp_scores=tscores, T=0.07)
    print("tiny:", ind.shape, "sk==mm idx frac", (ind == i.numpy()).mean())


def gen_c1():
    """BASELINE.json configs[0]: 10k x 512 gallery, 1k queries, k=20, T=0.07 (outputs only)."""
    bank, bl, qs, ql, cfg = synth.make_config("C1")
    k = cfg["k"]
    bn, qn = O.normalize(bank), O.normalize(qs)
    classes = np.arange(cfg["classes"])
    pred, dist, ind = O.sklearn_knn(bn.numpy(), bl.numpy(), qn.numpy(), k)
    v, i = O.mm_topk(qn, bn, k + 12)     # 12 extra columns sharpen the boundary test
    tpred, _ = O.vote_temperature(v[:, :k].numpy(), bl.numpy()[i[:, :k].numpy()], classes, cfg["T"])
    upred = O.vote_uniform(bl.numpy()[i[:, :k].numpy()], classes)
    np.savez_compressed(
        os.path.join(OUT, "c1.npz"), bank_digest=digest(bank), queries_digest=digest(qs),
        labels_digest=digest(bl), k=k, T=cfg["T"], sk_pred=pred.astype(np.int16),
        sk_ind=ind.astype(np.int32), sk_dist=dist, mm_idx=i.numpy().astype(np.int32),
        mm_sims=v.numpy(), temp_pred=tpred.astype(np.int16), uni_pred=upred.astype(np.int16))
    print("c1: sk pred == uniform vote on mm idx:", (pred == upred).mean(),
          " sk_ind==mm_idx", (ind == i[:, :k].numpy()).mean())


def gen_real27():
    """Repo-real shape: 11,269 x 768 bank / 6,088 queries with the REAL 27-class label column
    (non-contiguous ids 2..52) of data_{train,test}_combination3.csv; sklearn prediction for
    every k of ``Classifier.knn_eval`` (classification_engine.py:71)."""
    import pandas as pd
    ref = "/root/reference/HairPretraining/data"
    ytr = pd.read_csv(os.path.join(ref, "data_train_combination3.csv"))["class"].to_numpy()
    yte = pd.read_csv(os.path.join(ref, "data_test_combination3.csv"))["class"].to_numpy()
    classes = np.unique(ytr)
    d = 768
    tr_i = torch.from_numpy(np.searchsorted(classes, ytr))
    te_i = torch.from_numpy(np.searchsorted(classes, yte))
    bank, _ = synth.make_clustered(len(ytr), d, len(classes), 2027, labels=tr_i)
    qs, _ = synth.make_clustered(len(yte), d, len(classes), 2028, labels=te_i)
    bn, qn = O.normalize(bank), O.normalize(qs)
    ks = (5, 10, 20, 27, 30, 40, 642)
    preds = {}
    for k in ks:
        pred, _, _ = O.sklearn_knn(bn.numpy(), ytr, qn.numpy(), k)
        preds[f"sk_pred_k{k}"] = pred.astype(np.int16)
        print("real27 k", k, "acc", (pred == yte).mean())
    sub = 512                            # neighbour lists for the first 512 queries only
    v, i = O.mm_topk(qn[:sub], bn, 52)
    np.savez_compressed(
        os.path.join(OUT, "real27.npz"), train_labels=ytr.astype(np.int16),
        test_labels=yte.astype(np.int16), classes=classes.astype(np.int16), ks=np.array(ks),
        bank_digest=digest(bank), queries_digest=digest(qs), mm_idx=i.numpy().astype(np.int32),
        mm_sims=v.numpy(), **preds)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    which = sys.argv[1:] or ["tiny", "c1", "real27"]
    for w in which:
        {"tiny": gen_tiny, "c1": gen_c1, "real27": gen_real27}[w]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
