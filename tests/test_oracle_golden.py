"""CPU: the oracle (reference call sequences on sklearn / torch / numpy) against the committed
golden fixtures, and the three reference formulations against each other."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
import hcir_b200
from hcir_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_tiny_golden_reproduces(golden_dir):
    g = _load(golden_dir, "tiny.npz")
    k = int(g["k"])
    bn, qn = O.normalize(g["feats"]), O.normalize(g["queries"])
    np.testing.assert_array_equal(bn.numpy(), g["bank_unit"])
    np.testing.assert_array_equal(qn.numpy(), g["queries_unit"])
    pred, dist, ind = O.sklearn_knn(bn.numpy(), g["labels"], qn.numpy(), k)
    np.testing.assert_array_equal(pred, g["sk_pred"])
    np.testing.assert_array_equal(ind, g["sk_ind"])
    np.testing.assert_allclose(dist, g["sk_dist"], rtol=0, atol=1e-6)
    v, i = O.mm_topk(qn, bn, k)
    np.testing.assert_array_equal(i.numpy(), g["mm_idx"])
    np.testing.assert_allclose(v.numpy(), g["mm_sims"], rtol=0, atol=1e-6)


def test_three_formulations_agree_on_tiny(golden_dir):
    """sklearn kNN, torch mm+topk and cosine_similarity+argsort are the same operator; they may
    differ only inside the documented near-tie window (exact duplicate rows 100/200)."""
    g = _load(golden_dir, "tiny.npz")
    k = int(g["k"])
    s = O.similarity_matrix(g["queries_unit"], g["bank_unit"]).numpy()
    for idx, sims, atol in ((g["sk_ind"], 1.0 - g["sk_dist"], 2e-6), (g["mm_idx"], g["mm_sims"], 2e-7),
                            (g["argsort_idx"], g["argsort_sims"], 2e-6), (g["canon_idx"], g["canon_sims"], 2e-7)):
        bad = O.check_topk_against_sims(idx, sims, s, k, atol=atol)
        assert not any(bad.values()), bad


def test_canonical_tie_order_on_duplicates(golden_dir):
    g = _load(golden_dir, "tiny.npz")
    # query 3 is collinear with the duplicated rows 100 and 200: exact tie, ascending index first
    assert list(g["canon_idx"][3][:2]) == [100, 200]


def test_c1_golden_reproduces(golden_dir):
    g = _load(golden_dir, "c1.npz")
    bank, bl, qs, ql, cfg = synth.make_config("C1")
    import hashlib
    dig = lambda t: hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()[:16]
    assert dig(bank) == str(g["bank_digest"]) and dig(qs) == str(g["queries_digest"])
    k = int(g["k"])
    bn, qn = O.normalize(bank), O.normalize(qs)
    v, i = O.mm_topk(qn, bn, k + 12)
    np.testing.assert_array_equal(i.numpy().astype(np.int32), g["mm_idx"])
    up = O.vote_uniform(bl.numpy()[i[:, :k].numpy()], np.arange(cfg["classes"]))
    np.testing.assert_array_equal(up, g["uni_pred"])
    # sklearn's own prediction == uniform vote over its own neighbour lists
    np.testing.assert_array_equal(O.vote_uniform(bl.numpy()[g["sk_ind"]], np.arange(cfg["classes"])), g["sk_pred"])
    tp, _ = O.vote_temperature(v[:, :k].numpy(), bl.numpy()[i[:, :k].numpy()], np.arange(cfg["classes"]), float(g["T"]))
    np.testing.assert_array_equal(tp, g["temp_pred"])


def test_vote_uniform_matches_sklearn_mode_ties():
    # ties -> smallest class (sklearn _mode / argmax convention)
    nl = np.array([[5, 2, 2, 5], [9, 9, 4, 4], [7, 3, 1, 0]])
    assert O.vote_uniform(nl, np.array([0, 1, 2, 3, 4, 5, 7, 9])).tolist() == [2, 4, 0]


def test_vote_temperature_basic():
    sims = np.array([[0.9, 0.8, 0.1]], dtype=np.float32)
    nl = np.array([[3, 1, 1]])
    pred, sc = O.vote_temperature(sims, nl, np.array([1, 3]), T=0.07)
    assert pred.tolist() == [3] and sc.shape == (1, 2)
    pred, _ = O.vote_temperature(sims, nl, np.array([1, 3]), T=100.0)  # ~uniform: class 1 has 2 votes
    assert pred.tolist() == [1]


def test_checker_detects_errors():
    torch.manual_seed(0)
    b, q = O.normalize(torch.randn(500, 32)), O.normalize(torch.randn(7, 32))
    s = O.similarity_matrix(q, b).numpy()
    v, i = O.mm_topk(q, b, 10)
    ok = O.check_topk_against_sims(i.numpy(), v.numpy(), s, 10)
    assert not any(ok.values())
    bad_i = i.numpy().copy()
    bad_i[0, 0] = int(np.argmin(s[0]))
    bad = O.check_topk_against_sims(bad_i, v.numpy(), s, 10)
    assert bad["missing"] >= 1 and bad["intruder"] >= 1
    sw = i.numpy().copy()
    sw[1, [0, 9]] = sw[1, [9, 0]]
    svals = v.numpy().copy()
    svals[1, [0, 9]] = svals[1, [9, 0]]
    assert O.check_topk_against_sims(sw, svals, s, 10)["order"] >= 1
