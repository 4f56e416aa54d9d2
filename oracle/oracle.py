"""CPU oracle for the HSimCLR exact-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the reference's own call sequences for the path
(L2-normalise -> cosine similarity -> exact top-k -> kNN vote).  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path (``hcir_b200``) never does.

The arithmetic of the reference path lives in third-party libraries that are not vendored
in /root/reference but ARE installed in this image: torch 2.11 (CPU ATen/MKL), scikit-learn
1.9.0 (reference leaves it unpinned: ``requirements.txt:6``), numpy 2.3.  Each function below
is the reference's lines re-typed with the unavailable imports (timm, lightly, umap,
matplotlib) stripped, and cites the file:line it follows.

Parity pin: the reference ships NO tests, golden vectors or embeddings for this path
(SURVEY.md section 4 / 8c), so the pin is "outputs of the reference's own call sequence run
here" -- ``oracle/gen_golden.py`` executes these functions (i.e. the real sklearn / torch /
numpy code the reference calls) on seeded inputs and commits the results under
``tests/golden/``.  The temperature-weighted vote (``vote_temperature``) is an EXTENSION
that the reference does not contain; its parity is unpinned by the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# Near-tie window of the documented tie policy: 4 * ulp(1.0f).
TAU = 4.0 * float(np.finfo(np.float32).eps)


# --------------------------------------------------------------------------------------
# (a1) feature-bank normalisation
# --------------------------------------------------------------------------------------
def normalize(features) -> torch.Tensor:
    """``torch.nn.functional.normalize(f, dim=1)`` exactly as called at
    HairPretraining/src/classification_engine.py:50,62 and
    experiments/DualViewHair/scripts/qualitative_test.py:57,76 (eps=1e-12 clamp)."""
    f = torch.as_tensor(features, dtype=torch.float32)
    return F.normalize(f, dim=1)


# --------------------------------------------------------------------------------------
# (a2 + a6) sklearn brute cosine kNN + uniform vote  -- the authoritative classifier oracle
# --------------------------------------------------------------------------------------
def sklearn_knn(bank, labels, queries, k):
    """HairPretraining/src/classification_engine.py:80-82:

        knn = KNeighborsClassifier(n_neighbors=k, metric="cosine")
        knn.fit(self.training_features, self.training_labels)
        y_pred = knn.predict(self.testing_features)

    Returns (y_pred [Q] int64, neigh_dist [Q,k] float32 ascending, neigh_ind [Q,k] int64).
    ``kneighbors`` is the public accessor of the same internal neighbour search."""
    from sklearn.neighbors import KNeighborsClassifier

    bank = np.asarray(bank, dtype=np.float32)
    queries = np.asarray(queries, dtype=np.float32)
    labels = np.asarray(labels)
    knn = KNeighborsClassifier(n_neighbors=k, metric="cosine")
    knn.fit(bank, labels)
    y_pred = knn.predict(queries)
    dist, ind = knn.kneighbors(queries)
    return y_pred.astype(np.int64), dist.astype(np.float32), ind.astype(np.int64)


# --------------------------------------------------------------------------------------
# (a3) single-query retrieval on un-normalised embeddings
# --------------------------------------------------------------------------------------
def cosine_argsort(query_embedding, all_embeddings, top_k=5):
    """src/models/hair_encoder.py:193-194 (== src/models/face_encoder.py:210-211):

        similarities = cosine_similarity([query_embedding], all_embeddings)[0]
        top_indices = np.argsort(similarities)[::-1][:top_k]

    Returns (top_indices [k] int64, similarities[top_indices] [k] float32)."""
    from sklearn.metrics.pairwise import cosine_similarity

    similarities = cosine_similarity([np.asarray(query_embedding)], np.asarray(all_embeddings))[0]
    top_indices = np.argsort(similarities)[::-1][:top_k]
    return top_indices.astype(np.int64), similarities[top_indices].astype(np.float32)


def retrieve_similar_images(query_embedding, all_embeddings, all_paths, top_k=5):
    """src/models/hair_encoder.py:180-198 in full (list of {'path','similarity'} dicts)."""
    idx, sims = cosine_argsort(query_embedding, all_embeddings, top_k)
    return [{"path": all_paths[i], "similarity": s} for i, s in zip(idx.tolist(), sims)]


# --------------------------------------------------------------------------------------
# (a4 / a5) fp32 torch.mm + torch.topk  -- the oracle formulation the north star names
# --------------------------------------------------------------------------------------
def similarity_matrix(query_embeddings, gallery_embeddings) -> torch.Tensor:
    """experiments/DualViewHair/src/models/dual_view_model.py:317-335 /
    qualitative_test.py:79: ``torch.mm(q, G.t())`` on unit fp32 rows (CPU)."""
    q = torch.as_tensor(query_embeddings, dtype=torch.float32)
    g = torch.as_tensor(gallery_embeddings, dtype=torch.float32)
    return torch.mm(q, g.t())


def mm_topk(query_unit, gallery_unit, k):
    """experiments/DualViewHair/scripts/qualitative_test.py:79-84 batched over queries:
    ``similarities = torch.mm(q, G.t()); _, indices = torch.topk(similarities, k)``.
    Returns (sims [Q,k] float32 descending, idx [Q,k] int64)."""
    s = similarity_matrix(query_unit, gallery_unit)
    v, i = torch.topk(s, k, dim=1)
    return v, i


def mm_topk_drop_self(query_unit, gallery_unit, k):
    """qualitative_test.py:82-84: top-(k+1), drop the first (self) hit."""
    s = similarity_matrix(query_unit, gallery_unit)
    _, i = torch.topk(s, k + 1, dim=1)
    i = i[:, 1:]
    return torch.gather(s, 1, i), i


def mm_topk_chunked(query_unit, gallery_unit, k, chunk=1024):
    """Same operator as :func:`mm_topk`, evaluated in query chunks so a [Q,N] matrix that does
    not fit in host memory is never materialised (used for the larger parity cases)."""
    q = torch.as_tensor(query_unit, dtype=torch.float32)
    g = torch.as_tensor(gallery_unit, dtype=torch.float32)
    vs, is_ = [], []
    for a in range(0, q.shape[0], chunk):
        v, i = torch.topk(torch.mm(q[a:a + chunk], g.t()), k, dim=1)
        vs.append(v)
        is_.append(i)
    return torch.cat(vs), torch.cat(is_)


def canonical_topk(query_unit, gallery_unit, k):
    """fp32 ``torch.mm`` similarities ordered by the build's canonical rule
    (descending similarity, ties -> ascending gallery index).  numpy lexsort, exact."""
    s = similarity_matrix(query_unit, gallery_unit).numpy()
    n = s.shape[1]
    idx = np.empty((s.shape[0], k), dtype=np.int64)
    for r in range(s.shape[0]):
        order = np.lexsort((np.arange(n), -s[r]))
        idx[r] = order[:k]
    return np.take_along_axis(s, idx, axis=1), idx


# --------------------------------------------------------------------------------------
# (a6) uniform vote and (a7, extension) temperature-weighted vote
# --------------------------------------------------------------------------------------
def vote_uniform(neigh_labels, classes=None):
    """sklearn ``_mode(_y[neigh_ind])`` (neighbors/_classification.py:299-307 of the installed
    1.9.0): per-row histogram over the k neighbour labels, arg-max, ties -> smallest class.
    ``neigh_labels`` holds original label values; returns original label values."""
    neigh_labels = np.asarray(neigh_labels)
    if classes is None:
        classes = np.unique(neigh_labels)
    classes = np.asarray(classes)
    cls_idx = np.searchsorted(classes, neigh_labels)
    q = neigh_labels.shape[0]
    hist = np.zeros((q, classes.shape[0]), dtype=np.int64)
    np.add.at(hist, (np.arange(q)[:, None], cls_idx), 1)
    return classes[np.argmax(hist, axis=1)].astype(np.int64)


def vote_temperature(sims, neigh_labels, classes, T=0.07):
    """EXTENSION (not in the reference; BASELINE.json config 1 'T=0.07 weighted vote').
    InstDisc / lightly ``knn_predict`` form: ``score[c] = sum_{j: y_j = c} exp(s_j / T)``,
    arg-max, ties -> smallest class.  Evaluated as ``exp((s_j - s_0) / T)`` (s_0 = the query's
    best similarity), which leaves the arg-max unchanged and cannot overflow for small T.
    fp32, neighbours accumulated in rank order j = 0..k-1.  Returns (pred [Q], scores [Q,C])."""
    sims = np.asarray(sims, dtype=np.float32)
    neigh_labels = np.asarray(neigh_labels)
    classes = np.asarray(classes)
    cls_idx = np.searchsorted(classes, neigh_labels)
    q, k = sims.shape
    inv_t = np.float32(1.0) / np.float32(T)
    w = np.exp(((sims - sims[:, :1]) * inv_t).astype(np.float32)).astype(np.float32)
    scores = np.zeros((q, classes.shape[0]), dtype=np.float32)
    rows = np.arange(q)
    for j in range(k):  # rank order, fp32 accumulate
        scores[rows, cls_idx[:, j]] += w[:, j]
    return classes[np.argmax(scores, axis=1)].astype(np.int64), scores


def knn_predict_torch(bank_unit, labels, queries_unit, k, classes=None, T=None):
    """torch CPU restatement of the whole path used as the second CPU baseline
    (BASELINE.md section 3 item 2): ``mm`` -> ``topk`` -> vote(labels[idx])."""
    labels = np.asarray(labels)
    if classes is None:
        classes = np.unique(labels)
    v, i = mm_topk(queries_unit, bank_unit, k)
    nl = labels[i.numpy()]
    if T is None:
        return vote_uniform(nl, classes), v, i
    pred, _ = vote_temperature(v.numpy(), nl, classes, T)
    return pred, v, i


# --------------------------------------------------------------------------------------
# parity checker implementing the documented tie policy
# --------------------------------------------------------------------------------------
def check_topk_against_sims(our_idx, our_sims, oracle_sims_full, k, *, tau=TAU, rtol=1e-5,
                            atol=2e-7):
    """Compare a [Q,k] result with the oracle's full fp32 similarity matrix [Q,N].

    Policy: the index SET must equal the oracle's except for
    candidates within ``tau`` of the k-th boundary; the ORDER must be descending in the
    oracle's similarities up to ``tau``; the returned similarities must match the oracle's
    value for the same index within ``rtol`` relative.  Returns a dict of violation counts
    (all zero == parity)."""
    our_idx = np.asarray(our_idx)
    our_sims = np.asarray(our_sims, dtype=np.float32)
    s = np.asarray(oracle_sims_full, dtype=np.float32)
    q, n = s.shape
    assert our_idx.shape == (q, k), (our_idx.shape, (q, k))
    bad = {"range": 0, "dup": 0, "value": 0, "missing": 0, "intruder": 0, "order": 0}
    if ((our_idx < 0) | (our_idx >= n)).any():
        bad["range"] = int(((our_idx < 0) | (our_idx >= n)).sum())
        return bad
    kth = np.partition(s, n - k, axis=1)[:, n - k]  # k-th largest per row
    got = np.take_along_axis(s, our_idx, axis=1)
    bad["value"] = int((np.abs(got - our_sims) > rtol * np.abs(got) + atol).sum())
    bad["intruder"] = int((got < (kth[:, None] - tau)).sum())
    for r in range(q):
        if len(np.unique(our_idx[r])) != k:
            bad["dup"] += 1
        must = np.nonzero(s[r] > kth[r] + tau)[0]
        bad["missing"] += int(len(np.setdiff1d(must, our_idx[r])))
    bad["order"] = int((got[:, 1:] > got[:, :-1] + tau).sum())
    return bad


def check_topk_against_topk(our_idx, our_sims, ora_idx, ora_sims, *, tau=TAU, rtol=1e-5,
                            atol=2e-7):
    """Same policy when only the oracle's top-(k+m) lists are available (large cases).
    ``ora_*`` must be at least as wide as ``our_*``; extra oracle columns sharpen the
    boundary test.  Returns violation counts."""
    our_idx = np.asarray(our_idx)
    our_sims = np.asarray(our_sims, dtype=np.float32)
    ora_idx = np.asarray(ora_idx)
    ora_sims = np.asarray(ora_sims, dtype=np.float32)
    q, k = our_idx.shape
    bad = {"value": 0, "missing": 0, "intruder": 0, "order": 0}
    kth = ora_sims[:, k - 1]
    for r in range(q):
        lut = dict(zip(ora_idx[r].tolist(), ora_sims[r].tolist()))
        for j in range(k):
            i = int(our_idx[r, j])
            if i in lut:
                if abs(lut[i] - our_sims[r, j]) > rtol * abs(lut[i]) + atol:
                    bad["value"] += 1
            elif our_sims[r, j] < kth[r] - tau - atol:
                bad["intruder"] += 1
        must = ora_idx[r][ora_sims[r] > kth[r] + tau]
        bad["missing"] += int(len(np.setdiff1d(must, our_idx[r])))
    bad["order"] = int((our_sims[:, 1:] > our_sims[:, :-1] + tau).sum())
    return bad


def labels_agree_except_vote_ties(our_pred, neigh_labels, classes, sims=None, T=None,
                                  rel_margin=1e-5):
    """Predicted labels must equal the oracle vote on the SAME neighbour lists except where
    the vote itself is a (near-)tie.  Returns the number of unexplained mismatches."""
    our_pred = np.asarray(our_pred)
    if T is None:
        ref = vote_uniform(neigh_labels, classes)
        return int((ref != our_pred).sum())
    ref, scores = vote_temperature(sims, neigh_labels, classes, T)
    mism = np.nonzero(ref != our_pred)[0]
    unexplained = 0
    cls = np.asarray(classes)
    for r in mism:
        top = scores[r].max()
        mine = scores[r][np.searchsorted(cls, our_pred[r])]
        if not (top - mine <= rel_margin * top):
            unexplained += 1
    return unexplained
