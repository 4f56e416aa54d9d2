"""GPU parity tests (run with ``-m gpu`` on a B200): every CUDA kernel behind the C ABI against
the CPU oracle on the same seeded inputs, the committed golden fixtures, and size-independent
properties.  Index / label results are compared bit-exactly under the documented tie policy
(oracle.check_topk_*: canonical order = descending similarity, ties -> ascending index;
near-tie window tau = 4 ulp(1)); similarities within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

import hcir_b200
from hcir_b200 import GalleryBank, KNeighborsClassifierB200, _lib, synth
from hcir_b200.engine import l2_normalize
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a B200; there is no CPU fallback"
    assert _lib.load().hcir_device_supported() == 1
    torch.cuda.set_device(0)


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _clean(bad):
    return not any(bad.values())


# ------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("n,d", [(1, 8), (257, 100), (1000, 512), (4097, 768), (33, 2048), (5, 7)])
def test_l2norm_matches_f_normalize(n, d):
    g = torch.Generator().manual_seed(n * 31 + d)
    x = torch.randn(n, d, generator=g) * 3.0
    if n > 3:
        x[2] = 0.0  # zero row: F.normalize clamps the norm at 1e-12 -> stays zero
    ref = O.normalize(x)
    o32, obf, dl = l2_normalize(x.cuda())
    ld = o32.shape[1]
    assert ld % 64 == 0 and ld >= d
    got = o32[:, :d].cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=3e-7, atol=1e-9)
    if ld > d:
        assert torch.count_nonzero(o32[:, d:]) == 0 and torch.count_nonzero(obf[:, d:].float()) == 0
    # bf16 bank is the round-to-nearest of the fp32 unit rows; delta bounds the rounding error
    assert torch.equal(obf, o32.to(torch.bfloat16))
    err = (o32 - obf.float()).norm(dim=1)
    assert (dl >= err * 0.9999).all() and (dl <= err * 1.001 + 1e-9).all()


def test_l2norm_strided_and_unaligned_input():
    x = torch.randn(100, 200).cuda()
    view = x[:, 3:103]  # row stride 200, 12-byte offset: scalar-load path
    o32, _, _ = l2_normalize(view)
    ref = O.normalize(view.cpu())
    np.testing.assert_allclose(o32[:, :100].cpu().numpy(), ref.numpy(), rtol=3e-7, atol=1e-9)


# ------------------------------------------------------------------------------------ K2 (raw tiles)
@pytest.mark.parametrize("flags", [0, 16, 48])  # 1-CTA (default); CTA pairs (cta_group::2); pairs + rotation
@pytest.mark.parametrize("nq,ng,d", [(128, 256, 64), (1, 300, 64), (130, 1000, 128), (257, 2049, 768),
                                     (64, 5000, 512), (300, 777, 2048)])
def test_simtopk_accumulator_tiles_match_fp32_matmul(nq, ng, d, flags):
    """The tcgen05 contraction itself: dump every accumulator value and compare with an fp32
    matmul of the same bf16 operands (bf16 products are exact in fp32; only the accumulation
    order differs)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(nq + ng + d)
    q = torch.randn(nq, d, generator=g).cuda()
    b = torch.randn(ng, d, generator=g).cuda()
    _, qbf, _ = l2_normalize(q, want_f32=False, want_delta=False)
    _, gbf, _ = l2_normalize(b, want_f32=False, want_delta=False)
    ld = qbf.shape[1]
    kc = 40
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(nq, ng, ld, kc, 148, plan))
    plan.flags = flags
    ws = torch.zeros(int(plan.bytes), dtype=torch.uint8, device="cuda")
    scores = torch.full((nq, ng), float("nan"), device="cuda")
    _lib.check(lib.hcir_simtopk_debug(qbf.data_ptr(), nq, gbf.data_ptr(), ng, ld, plan, ws.data_ptr(),
                                      scores.data_ptr(), torch.cuda.current_stream().cuda_stream), "simtopk_debug")
    torch.cuda.synchronize()
    ref = qbf.float().double() @ gbf.float().double().t()
    err = (scores.double() - ref).abs().max().item()
    assert err < 2e-6, f"max |tile - ref| = {err}"
    # candidate lists: a superset of the top-kc by these scores
    counts = ws[: nq * plan.nlists * 4].view(torch.int32).view(nq, plan.nlists).cpu().numpy()
    keys = ws[plan.keys_off:].view(torch.int64).view(nq, plan.nlists, plan.cap).cpu().numpy()
    sc = scores.cpu().numpy()
    kk = min(kc, ng)
    for r in range(0, nq, max(1, nq // 16)):
        cand = set()
        for s in range(plan.nlists):
            ks = keys[r, s, : counts[r, s]].astype(np.uint64)
            cand |= set((0xFFFFFFFF - (ks & np.uint64(0xFFFFFFFF))).astype(np.int64).tolist())
        kth = np.sort(sc[r])[::-1][kk - 1]
        must = set(np.nonzero(sc[r] > kth)[0].tolist())
        assert must <= cand, f"row {r}: {len(must - cand)} top-{kk} rows missing from the candidate lists"
        assert all(0 <= c < ng for c in cand)


# ------------------------------------------------------------------------------------ exact path + tensor path
def _check_bank(bank_feats, queries, k, mode, *, atol=2e-7):
    gb = GalleryBank(bank_feats)
    sims, idx = gb.topk(queries, k, mode=mode)
    bn, qn = O.normalize(bank_feats), O.normalize(queries)
    s = O.similarity_matrix(qn, bn).numpy()
    bad = O.check_topk_against_sims(np.asarray(idx), np.asarray(sims), s, k, atol=atol)
    assert _clean(bad), (mode, bad, gb.last_stats)
    return gb, np.asarray(sims), np.asarray(idx)


@pytest.mark.parametrize("mode", ["exact", "tensor"])
def test_tiny_golden(golden_dir, mode):
    g = _golden(golden_dir, "tiny.npz")
    k = int(g["k"])
    gb = GalleryBank(g["feats"], g["labels"])
    sims, idx = gb.topk(g["queries"], k, mode=mode)
    s = O.similarity_matrix(g["queries_unit"], g["bank_unit"]).numpy()
    assert _clean(O.check_topk_against_sims(idx, sims, s, k))
    # exact tie (duplicate rows 100 / 200, collinear query 3): canonical order = ascending index
    assert list(idx[3][:2]) == [100, 200]
    assert sims[3][0] == sims[3][1]
    # uniform vote == sklearn's prediction (non-contiguous labels 2..52)
    pred = gb.predict(g["queries"], k, mode=mode)
    nl = g["labels"][idx]
    assert O.labels_agree_except_vote_ties(pred, nl, g["classes"]) == 0
    agree = (pred == g["sk_pred"]).mean()
    assert agree >= 16 / 17, agree  # sklearn's own tie order may differ on the duplicate pair only
    predT = gb.predict(g["queries"], k, T=float(g["T"]), mode=mode)
    assert O.labels_agree_except_vote_ties(predT, nl, g["classes"], sims, float(g["T"])) == 0
    np.testing.assert_array_equal(predT, g["temp_pred"])


@pytest.mark.parametrize("mode", ["exact", "tensor"])
def test_c1_golden(golden_dir, mode):
    """BASELINE.json configs[0]: 10k x 512 gallery, 1k queries, k=20, T=0.07."""
    g = _golden(golden_dir, "c1.npz")
    bank, bl, qs, _, cfg = synth.make_config("C1")
    k = int(g["k"])
    gb = GalleryBank(bank, bl)
    sims, idx = gb.topk(qs, k, mode=mode)
    sims, idx = sims.numpy(), idx.numpy()
    bad = O.check_topk_against_topk(idx, sims, g["mm_idx"].astype(np.int64), g["mm_sims"])
    assert _clean(bad), bad
    assert (idx == g["mm_idx"][:, :k]).mean() > 0.9995  # identical except near-tie swaps
    classes = np.arange(cfg["classes"])
    pu = gb.predict(qs, k, mode=mode).numpy()
    pt = gb.predict(qs, k, T=float(g["T"]), mode=mode).numpy()
    nl = bl.numpy()[idx]
    assert O.labels_agree_except_vote_ties(pu, nl, classes) == 0
    assert O.labels_agree_except_vote_ties(pt, nl, classes, sims, float(g["T"])) == 0
    assert (pu == g["sk_pred"]).mean() >= 0.998 and (pu == g["uni_pred"]).mean() >= 0.998
    assert (pt == g["temp_pred"]).mean() >= 0.998


def test_real27_golden_k_sweep(golden_dir):
    """Repo-real shape (11,269 x 768 bank, 6,088 queries, real 27-class labels 2..52) and the
    reference's whole k sweep (classification_engine.py:71) from ONE search at k=642."""
    g = _golden(golden_dir, "real27.npz")
    classes = g["classes"].astype(np.int64)
    ytr, yte = g["train_labels"].astype(np.int64), g["test_labels"].astype(np.int64)
    tr_i = torch.from_numpy(np.searchsorted(classes, ytr))
    te_i = torch.from_numpy(np.searchsorted(classes, yte))
    bank, _ = synth.make_clustered(len(ytr), 768, len(classes), 2027, labels=tr_i)
    qs, _ = synth.make_clustered(len(yte), 768, len(classes), 2028, labels=te_i)
    knn = KNeighborsClassifierB200(n_neighbors=642, metric="cosine").fit(bank, ytr)
    np.testing.assert_array_equal(knn.classes_, classes)
    ks = [int(k) for k in g["ks"]]
    preds = knn.predict_multi_k(qs, ks)
    for k in ks:
        agree = (preds[k] == g[f"sk_pred_k{k}"].astype(np.int64)).mean()
        assert agree >= 0.998, (k, agree)  # sklearn near-tie neighbour order only
    dist, ind = knn.kneighbors(qs[:512], n_neighbors=52)
    bad = O.check_topk_against_topk(ind, 1.0 - dist, g["mm_idx"].astype(np.int64), g["mm_sims"], atol=5e-7)
    assert _clean(bad), bad
    # single-k predict == prefix vote of the k-sweep
    one = KNeighborsClassifierB200(n_neighbors=20).fit(bank, ytr).predict(qs)
    np.testing.assert_array_equal(one, preds[20])


@pytest.mark.parametrize("mode", ["exact", "tensor"])
@pytest.mark.parametrize("n,d,q,k", [(5000, 100, 33, 7), (4099, 768, 129, 20), (9000, 512, 1, 1),
                                     (12000, 256, 260, 100), (6000, 64, 50, 642),
                                     (30000, 2048, 300, 200), (8000, 2048, 40, 50)])   # C5's width and k
def test_random_shapes(mode, n, d, q, k):
    bank, _ = synth.make_clustered(n, d, 13, 5 + n)
    qs, _ = synth.make_clustered(q, d, 13, 6 + n)
    _check_bank(bank, qs, k, mode)


@pytest.mark.parametrize("mode", ["exact", "tensor"])
def test_adversarial_duplicates_zero_rows_and_sorted_gallery(mode):
    """Exact ties (a block of duplicated rows), zero rows, k larger than the number of distinct
    rows, and a gallery ordered by ascending similarity to query 0 (worst case for a running
    threshold: every row beats the threshold seen so far)."""
    n, d, k = 7000, 128, 50
    bank, _ = synth.make_clustered(n, d, 5, 77)
    bank[1000:1040] = bank[999]          # 41 identical rows
    bank[10] = 0.0
    bank[6999] = 0.0
    qs, _ = synth.make_clustered(40, d, 5, 78)
    qs[1] = bank[999] * 0.5
    order = torch.argsort(O.similarity_matrix(O.normalize(qs[:1]), O.normalize(bank))[0])
    bank = bank[order]
    gb, sims, idx = _check_bank(bank, qs, k, mode)
    # the 41 duplicates of query 1 are an exact tie: ascending index order among them
    dup = np.nonzero((bank == qs[1] * 2.0).all(dim=1).numpy())[0]
    assert len(dup) == 41
    np.testing.assert_array_equal(idx[1][:41], np.sort(dup))


def test_tensor_and_exact_paths_are_bit_identical():
    bank, _ = synth.make_clustered(20000, 768, 27, 11)
    qs, _ = synth.make_clustered(500, 768, 27, 12)
    gb = GalleryBank(bank)
    s1, i1 = gb.topk(qs, 20, mode="tensor")
    st = dict(gb.last_stats)
    s2, i2 = gb.topk(qs, 20, mode="exact")
    assert torch.equal(i1, i2) and torch.equal(s1, s2), st
    assert st["path"] == "tensor"


def test_certification_fallback_is_exercised_and_exact():
    """A gallery of near-duplicates defeats the bf16 contraction (all scores within its error
    bound): those queries must be flagged uncertified and resolved by the exact fp32 kernel."""
    d, n = 256, 8192
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, d, generator=g)
    bank = base + 1e-4 * torch.randn(n, d, generator=g)
    qs = base + 1e-4 * torch.randn(64, d, generator=g)
    gb = GalleryBank(bank)
    s1, i1 = gb.topk(qs, 10, mode="tensor")
    assert gb.last_stats["uncertified"] > 0
    s2, i2 = gb.topk(qs, 10, mode="exact")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_properties_permutation_and_k_prefix():
    n, d = 30000, 768
    bank, _ = synth.make_clustered(n, d, 61, 21)
    qs, _ = synth.make_clustered(256, d, 61, 22)
    gb = GalleryBank(bank)
    s100, i100 = gb.topk(qs, 100)
    s20, i20 = gb.topk(qs, 20)
    assert torch.equal(i100[:, :20], i20) and torch.equal(s100[:, :20], s20)  # top-k is a prefix of top-k'
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))
    gp = GalleryBank(bank[perm])
    sp, ip = gp.topk(qs, 20)
    assert torch.equal(sp, s20)                       # same similarities
    mapped = perm[ip]
    same = (mapped == i20)
    # indices permute accordingly (exact fp32 ties could reorder; there are none in this data)
    assert same.all(), int((~same).sum())


# ------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("c,k", [(2, 1), (27, 20), (61, 642), (100, 5)])
def test_vote_kernel_against_oracle(c, k):
    g = torch.Generator().manual_seed(c * k)
    nq = 777
    sims, _ = torch.sort(torch.rand(nq, k, generator=g) * 0.5 + 0.3, dim=1, descending=True)
    nl = torch.randint(0, c, (nq, k), generator=g, dtype=torch.int32)
    nl[5] = 3 % c
    gb = GalleryBank(torch.randn(64, 16), torch.arange(64) % c) if c <= 64 else \
        GalleryBank(torch.randn(c, 16), torch.arange(c))
    gb.classes_ = np.arange(c)
    pu = gb.vote(sims.cuda(), nl.cuda()).cpu().numpy()
    np.testing.assert_array_equal(pu, O.vote_uniform(nl.numpy(), np.arange(c)))
    pt, sc = gb.vote(sims.cuda(), nl.cuda(), T=0.07, return_scores=True)
    ref_p, ref_s = O.vote_temperature(sims.numpy(), nl.numpy(), np.arange(c), 0.07)
    np.testing.assert_allclose(sc.cpu().numpy(), ref_s, rtol=2e-5, atol=1e-30)
    assert O.labels_agree_except_vote_ties(pt.cpu().numpy(), nl.numpy(), np.arange(c), sims.numpy(), 0.07) == 0


# ------------------------------------------------------------------------------------ K5 + sharding on one GPU
@pytest.mark.parametrize("G", [2, 3, 8])
def test_fake_shards_merge_equals_single_bank(G):
    """Multi-rank logic without a cluster: G shard banks on one GPU, K5 merge of their exact
    local lists == the single-bank result, bit for bit."""
    from hcir_b200.sharded import ShardPlan, merge_topk
    n, d, k = 25001, 768, 20
    bank, bl = synth.make_clustered(n, d, 27, 31)
    qs, _ = synth.make_clustered(300, d, 27, 32)
    full = GalleryBank(bank, bl)
    s_ref, i_ref = full.topk(qs, k, return_device=True)
    sp = ShardPlan(n, G)
    ss, ii, ll = [], [], []
    for r in range(G):
        sh = GalleryBank(bank[sp.start(r):sp.stop(r)], bl[sp.start(r):sp.stop(r)], idx_offset=sp.start(r),
                         classes=np.arange(27))
        s, i = sh.topk(qs, k, return_device=True)
        ss.append(s), ii.append(i), ll.append(sh.neighbour_labels(i))
    o_s, o_i, o_l = merge_topk(torch.stack(ss), torch.stack(ii), torch.stack(ll), k)
    assert torch.equal(o_i, i_ref) and torch.equal(o_s, s_ref)
    assert torch.equal(o_l, full.neighbour_labels(i_ref))


def test_merge_with_short_shards():
    from hcir_b200.sharded import merge_topk
    # shard 1 holds only 2 rows: padded with (-inf, -1)
    s = torch.tensor([[[0.9, 0.5, 0.1]], [[0.7, 0.2, float("-inf")]]]).cuda()
    i = torch.tensor([[[4, 2, 0]], [[11, 10, -1]]]).cuda()
    o_s, o_i, _ = merge_topk(s, i, None, 3)
    assert o_i.cpu().tolist() == [[4, 11, 2]] and o_s.cpu().tolist() == [[pytest.approx(0.9), pytest.approx(0.7), 0.5]]


# ------------------------------------------------------------------------------------ reference-shaped surfaces
def test_retrieve_similar_images_matches_reference_call(golden_dir):
    g = _golden(golden_dir, "tiny.npz")
    paths = [f"img_{i:04d}.jpg" for i in range(g["feats"].shape[0])]
    for r in (0, 3, 9):
        ours = hcir_b200.retrieve_similar_images(g["queries"][r], g["feats"], paths, top_k=5)
        ref = O.retrieve_similar_images(g["queries"][r], g["feats"], paths, top_k=5)
        assert [o["path"] for o in ours] == [paths[i] for i in g["canon_idx"][r]]
        assert {o["path"] for o in ours} == {x["path"] for x in ref}
        for o, s in zip(ours, g["canon_sims"][r]):
            assert isinstance(o["similarity"], np.float32) and abs(o["similarity"] - s) < 2e-6


def test_flat_index_and_drop_self():
    bank, _ = synth.make_clustered(5000, 256, 9, 41)
    bn = O.normalize(bank)
    index = hcir_b200.FlatIndex(256)
    index.add(bn.numpy()[:3000])
    index.add(bn.numpy()[3000:])
    assert index.ntotal == 5000
    D, I = index.search(bn.numpy()[:7], 5)
    assert (I[:, 0] == np.arange(7)).all() and np.abs(D[:, 0]).max() < 1e-5 and (np.diff(D, axis=1) >= 0).all()
    hr = hcir_b200.HairRetrievalB200(bn)
    out = hr.retrieve_similar(17, top_k=10)
    v, i = O.mm_topk_drop_self(bn[17:18], bn, 10)
    assert [r["gallery_idx"] for r in out["results"]] == i[0].tolist()


def test_errors_are_loud():
    gb = GalleryBank(torch.randn(100, 32))
    with pytest.raises(ValueError):
        gb.topk(torch.randn(3, 32), 101)
    with pytest.raises(ValueError):
        gb.topk(torch.randn(3, 31), 5)
    with pytest.raises(ValueError):
        gb.predict(torch.randn(3, 32), 5)  # no labels
    s, i = gb.topk(torch.randn(0, 32), 5)
    assert s.shape == (0, 5) and i.shape == (0, 5)


# ------------------------------------------------------------------------------------ full-size properties
def test_c2_full_size_properties():
    """BASELINE.json configs[1] at full size (200k x 768 gallery, 10k queries, k=20): checked
    through size-independent properties -- every query's own (planted) row is rank 1 with
    similarity 1, results sorted, the exact-path result of a query subsample is bit-identical,
    and sharding the gallery in 4 leaves the result unchanged."""
    from hcir_b200.sharded import ShardPlan, merge_topk
    dev = "cuda"
    n, d, q, k = 200_000, 768, 10_000, 20
    bank, bl = synth.make_clustered(n, d, 61, 1236, device=dev)
    qs, _ = synth.make_clustered(q, d, 61, 4323, device=dev)
    planted = torch.randperm(n, device=dev)[:512]
    qs[:512] = bank[planted] * 1.7
    gb = GalleryBank(bank, bl)
    sims, idx = gb.topk(qs, k)
    assert gb.last_stats["path"] == "tensor"
    assert torch.equal(idx[:512, 0], planted) and (sims[:512, 0] - 1).abs().max() < 1e-6
    assert (sims[:, 1:] <= sims[:, :-1]).all()
    assert len(torch.unique(idx[7])) == k
    sub = torch.arange(0, q, 97, device=dev)
    s2, i2 = gb.topk(qs[sub], k, mode="exact")
    assert torch.equal(i2, idx[sub]) and torch.equal(s2, sims[sub])
    sp = ShardPlan(n, 4)
    parts = [GalleryBank(bank[sp.start(r):sp.stop(r)], idx_offset=sp.start(r)).topk(qs, k) for r in range(4)]
    o_s, o_i, _ = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), None, k)
    assert torch.equal(o_i, idx) and torch.equal(o_s, sims)
    pred = gb.predict(qs[:2000], k)
    ref = O.vote_uniform(bl[idx[:2000]].cpu().numpy(), np.arange(61))
    np.testing.assert_array_equal(pred.cpu().numpy(), ref)


# ------------------------------------------------------------------------------------ next rows (8f)
def test_load_gallery_from_npy_and_shard(tmp_path):
    from hcir_b200 import formats
    bank, _ = synth.make_clustered(6000, 96, 7, 51)
    qs, _ = synth.make_clustered(20, 96, 7, 52)
    paths = [f"db/{i:05d}_hair.png" for i in range(6000)]
    formats.save_embeddings(str(tmp_path), bank.numpy(), paths)
    gb, p2 = formats.load_gallery(str(tmp_path), chunk_rows=1024)
    assert p2 == paths and gb.n == 6000
    s_ref, i_ref = GalleryBank(bank).topk(qs, 10)
    s, i = gb.topk(qs, 10)
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref)
    shard, _ = formats.load_gallery(str(tmp_path), rows=(2000, 5000))
    s2, i2 = shard.topk(qs, 10)
    assert int(i2.min()) >= 2000 and int(i2.max()) < 5000
    recs = formats.top100_records([f"q{j}_hair.png" for j in range(20)], i.numpy(), paths)
    assert recs[3]["top100"][0] == paths[int(i[3, 0])].split("/")[-1]


def test_kth_neighbour_matches_neg_sampler_static():
    """HairPretraining/src/neg_sampling.py:26-53 (cosine): k-th entry of the descending sort."""
    from hcir_b200 import metrics
    g = torch.Generator().manual_seed(8)
    emb = torch.randn(256, 512, generator=g)
    en = emb / torch.norm(emb, dim=1, keepdim=True).clamp(min=1e-8)
    sim = torch.mm(en, en.t())
    _, order = torch.sort(sim, dim=1, descending=True)
    for k in (1, 7, 64):
        ours = metrics.kth_neighbour(emb, k)
        ref = order[:, k - 1]
        same = ours == ref
        # any disagreement must be a near-tie in the reference's own fp32 similarities
        for r in torch.nonzero(~same).flatten().tolist():
            assert abs(float(sim[r, ours[r]] - sim[r, ref[r]])) < O.TAU
        assert same.float().mean() > 0.99
    assert torch.equal(metrics.kth_neighbour(emb, 1), torch.arange(256))  # rank 1 is the row itself

@pytest.mark.parametrize("n,nq", [(8192, 256), (16384, 256), (8192, 64), (5000, 700)])
def test_near_duplicate_gallery_never_certifies_on_list_completeness_alone(n, nq):
    """Every row of a near-duplicate gallery passes any threshold, so the candidate lists can hold
    the WHOLE gallery; that alone must not certify a query -- only kc rows reach the re-score."""
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 256, generator=g)
    gb = GalleryBank(base + 1e-4 * torch.randn(n, 256, generator=g))
    qd = base + 1e-4 * torch.randn(nq, 256, generator=g)
    s1, i1 = gb.topk(qd, 10, mode="tensor")
    s2, i2 = gb.topk(qd, 10, mode="exact")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_uncertified_queries_are_finished_by_a_second_tensor_pass():
    """300 near-identical neighbours per query: the kc = 84 best bf16 candidates cannot certify the
    fp32 top-10, but the second pass (threshold = best-so-far k-th score - eps, kc x 4) collects the
    whole band and certifies it -- no fp32 brute force, results identical to the exact path."""
    d = 256
    g = torch.Generator().manual_seed(11)
    bases = torch.randn(8, d, generator=g)
    dense = (bases[:, None, :] + 1e-3 * torch.randn(8, 300, d, generator=g)).reshape(-1, d)
    bank = torch.cat([torch.randn(20000, d, generator=g), dense])[torch.randperm(22400, generator=g)]
    qs = bases.repeat_interleave(4, 0) + 1e-3 * torch.randn(32, d, generator=g)
    gb = GalleryBank(bank)
    s1, i1 = gb.topk(qs, 10, mode="tensor")
    assert gb.last_stats["uncertified"] > 0
    assert gb.retry_stats["second_pass"] > 0 and gb.retry_stats["exact"] == 0, gb.retry_stats
    s2, i2 = gb.topk(qs, 10, mode="exact")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    sess = gb.session(32, 10, vote=False)            # same completion after a graph replay
    _, s3, i3 = sess.run(qs.cuda())
    assert torch.equal(i3.cpu(), i2) and torch.equal(s3.cpu(), s2)


def test_pipelined_submit_result_equals_run_and_redoes_uncertified_batches():
    bank, bl = synth.make_clustered(30000, 256, 27, 61)
    gb = GalleryBank(bank, bl)
    sess = gb.session(300, 20)
    batches = [synth.make_clustered(300, 256, 27, 70 + i)[0].cuda() for i in range(4)]
    pend = [sess.submit(b) for b in batches[:2]]            # two steps in flight
    got = [pend[0].result()]
    pend.append(sess.submit(batches[2]))
    got += [pend[1].result(), pend[2].result()]
    for b, (p, s_, i_) in zip(batches, got):
        p_ref, s_ref, i_ref = gb.predict(b, 20, return_neighbors=True)
        assert torch.equal(p, p_ref) and torch.equal(s_, s_ref) and torch.equal(i_, i_ref)
    assert not any(x.redone for x in pend)
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 256, generator=g)
    dup = GalleryBank(base + 1e-4 * torch.randn(8192, 256, generator=g), torch.arange(8192) % 5)
    qd = (base + 1e-4 * torch.randn(64, 256, generator=g)).cuda()
    h = dup.session(64, 10).submit(qd)
    p, s_, i_ = h.result()
    assert h.redone
    s2, i2 = dup.topk(qd, 10, mode="exact")
    assert torch.equal(i_, i2) and torch.equal(s_, s2) and torch.equal(p, dup.predict(qd, 10, mode="exact"))


def test_feature_bank_builder_equals_concatenate_then_fit():
    """classification_engine.py:42-53,66-67 without the per-batch .cpu(): batches appended on the
    device (fp32 and fp16 encoder outputs, a growth step, an empty batch) give the very bank that
    GalleryBank(torch.cat(batches)) builds, hence identical predictions."""
    from hcir_b200 import FeatureBankBuilder
    feats, labels = synth.make_clustered(9000, 512, 27, 81)
    labels = labels * 2 + 2                                       # non-contiguous label values
    sizes = [256, 256, 0, 1000, 3000, 4488]
    fb = FeatureBankBuilder(512, capacity=1024)
    a = 0
    for sz in sizes:
        fb.append(feats[a:a + sz].cuda(), labels[a:a + sz])
        a += sz
    bank = fb.finish()
    ref = GalleryBank(feats, labels)
    assert bank.n == 9000 and torch.equal(bank.g32, ref.g32) and torch.equal(bank.gbf, ref.gbf)
    assert bank.g_delta_max == ref.g_delta_max and np.array_equal(bank.classes_, ref.classes_)
    qs, _ = synth.make_clustered(500, 512, 27, 82)
    clf = KNeighborsClassifierB200(20).fit_bank(bank)
    np.testing.assert_array_equal(clf.predict(qs.numpy()), KNeighborsClassifierB200(20).fit(feats, labels.numpy()).predict(qs.numpy()))
    half = FeatureBankBuilder(512).append(feats[:100].cuda().half(), labels[:100]).finish()   # AMP encoder output
    assert torch.equal(half.g32, GalleryBank(feats[:100].half().float()).g32)
    with pytest.raises(ValueError):
        FeatureBankBuilder(512).append(feats[:4], labels[:4])     # host tensor: the point is to stay on the device
    with pytest.raises(ValueError):
        FeatureBankBuilder(256).append(feats[:4].cuda())


# ------------------------------------------------------------------------------------ CUDA-graph sessions
def test_session_replay_equals_eager_and_handles_fallback():
    bank, bl = synth.make_clustered(30000, 768, 27, 61)
    gb = GalleryBank(bank, bl)
    sess = gb.session(300, 20)
    assert sess is not None and gb.session(300, 20) is sess          # cached
    for seed in (62, 63):                                            # replay with new inputs
        qs, _ = synth.make_clustered(300, 768, 27, seed)
        pred, sims, idx = sess.run(qs.cuda())
        assert gb.last_stats["path"] == "tensor+graph"
        p_ref, s_ref, i_ref = gb.predict(qs, 20, return_neighbors=True)
        assert torch.equal(idx.cpu(), i_ref) and torch.equal(sims.cpu(), s_ref) and torch.equal(pred.cpu(), p_ref)
    with pytest.raises(ValueError):
        sess.run(torch.zeros(299, 768))
    # classifier front end uses the session transparently (host tensors in, numpy out)
    clf = KNeighborsClassifierB200(20).fit(bank, bl.numpy())
    qs, _ = synth.make_clustered(300, 768, 27, 64)
    a = clf.predict(qs)
    b = KNeighborsClassifierB200(20, use_graph=False).fit(bank, bl.numpy()).predict(qs)
    np.testing.assert_array_equal(a, b)
    # near-duplicate gallery: uncertified queries are finished by the exact kernel after the replay
    g = torch.Generator().manual_seed(3)
    base = torch.randn(1, 256, generator=g)
    dup = GalleryBank(base + 1e-4 * torch.randn(8192, 256, generator=g), torch.arange(8192) % 5)
    qd = base + 1e-4 * torch.randn(64, 256, generator=g)
    sd = dup.session(64, 10)
    pred, sims, idx = sd.run(qd.cuda())
    assert dup.last_stats["uncertified"] > 0
    s2, i2 = dup.topk(qd, 10, mode="exact")
    assert torch.equal(idx.cpu(), i2) and torch.equal(sims.cpu(), s2)
    assert torch.equal(pred.cpu(), dup.predict(qd, 10, mode="exact"))


def test_c3_full_size_properties():
    """BASELINE.json configs[2] at full size (1M x 768 gallery, 4096 queries, top-100): planted rows
    are rank 1, lists are sorted and duplicate-free, a query subsample is bit-identical on the exact
    fp32 path, top-20 is the prefix of top-100, and an 8-way gallery sharding merges to the same result."""
    from hcir_b200.sharded import ShardPlan, merge_topk
    dev = "cuda"
    n, d, q, k = 1_000_000, 768, 4096, 100
    bank, _ = synth.make_clustered(n, d, 61, 1237, device=dev)
    qs, _ = synth.make_clustered(q, d, 61, 4324, device=dev)
    planted = torch.randperm(n, device=dev)[:256]
    qs[:256] = bank[planted] * 0.3
    gb = GalleryBank(bank)
    sims, idx = gb.topk(qs, k)
    assert gb.last_stats["path"] == "tensor" and gb.last_stats["uncertified"] == 0
    assert torch.equal(idx[:256, 0], planted) and (sims[:256, 0] - 1).abs().max() < 1e-6
    assert (sims[:, 1:] <= sims[:, :-1]).all()
    assert all(len(torch.unique(idx[r])) == k for r in (0, 300, 4095))
    sub = torch.arange(0, q, 257, device=dev)
    s2, i2 = gb.topk(qs[sub], k, mode="exact")
    assert torch.equal(i2, idx[sub]) and torch.equal(s2, sims[sub])
    s20, i20 = gb.topk(qs, 20)
    assert torch.equal(i20, idx[:, :20]) and torch.equal(s20, sims[:, :20])
    sp = ShardPlan(n, 8)
    qsub = qs[:512]
    parts = [GalleryBank(bank[sp.start(r):sp.stop(r)], idx_offset=sp.start(r)).topk(qsub, k) for r in range(8)]
    o_s, o_i, _ = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), None, k)
    assert torch.equal(o_i, idx[:512]) and torch.equal(o_s, sims[:512])


def test_c4_streaming_regime_properties():
    """BASELINE.json configs[3] regime (small query batch against a multi-million-row gallery; 4M rows
    here to keep the test short): every batch size 1..64 returns the planted row first and agrees
    bit for bit with the exact fp32 path."""
    dev = "cuda"
    n, d, k = 4_000_000, 768, 20
    bank, _ = synth.make_clustered(n, d, 61, 1238, device=dev)
    gb = GalleryBank(bank)
    del bank
    for q in (1, 8, 64):
        qs, _ = synth.make_clustered(q, d, 61, 4325 + q, device=dev)
        rows = torch.randint(0, n, (q,), device=dev)
        qs[:] = gb.g32[rows, :d] * 2.5 + 0.02 * qs      # noisy copies of gallery rows
        sims, idx = gb.topk(qs, k)
        assert gb.last_stats["path"] == "tensor"
        assert torch.equal(idx[:, 0], rows)
        s2, i2 = gb.topk(qs, k, mode="exact")
        assert torch.equal(i2, idx) and torch.equal(s2, sims)


def test_packed_merge_reads_gathered_blocks_in_place():
    """hcir_merge_topk_packed on G emulated rank blocks [idx | sims | labels] == the dense merge."""
    from hcir_b200.sharded import ShardPlan, merge_topk
    lib = _lib.load()
    n, d, k, G, nq = 40000, 256, 20, 4, 130
    bank, bl = synth.make_clustered(n, d, 9, 81)
    qs, _ = synth.make_clustered(nq, d, 9, 82)
    sp = ShardPlan(n, G)
    blocks, dense = [], []
    for r in range(G):
        sh = GalleryBank(bank[sp.start(r):sp.stop(r)], bl[sp.start(r):sp.stop(r)], idx_offset=sp.start(r),
                         classes=np.arange(9))
        sess = sh.session(nq, k, vote=False, pack=True)
        _, s, i = sess.run(qs.cuda())
        assert sess.pack.numel() == lib.hcir_packed_block_bytes(nq, k, 1)
        assert torch.equal(sess.out_lab, sh.neighbour_labels(i))
        blocks.append(sess.pack.clone())
        dense.append((s.clone(), i.clone(), sess.out_lab.clone()))
    gathered = torch.cat(blocks)
    o_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    o_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    o_l = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    _lib.check(lib.hcir_merge_topk_packed(gathered.data_ptr(), G, nq, k, 1, 0, o_s.data_ptr(), o_i.data_ptr(),
                                          o_l.data_ptr(), torch.cuda.current_stream().cuda_stream))
    d_s, d_i, d_l = merge_topk(torch.stack([x[0] for x in dense]), torch.stack([x[1] for x in dense]),
                               torch.stack([x[2] for x in dense]), k)
    assert torch.equal(o_i, d_i) and torch.equal(o_s, d_s) and torch.equal(o_l, d_l)
    s_ref, i_ref = GalleryBank(bank).topk(qs, k, return_device=True)
    assert torch.equal(o_i, i_ref) and torch.equal(o_s, s_ref)


def test_overlapped_host_predict_equals_single_shot():
    """Host batches >= 8192 rows: head chunk searched while the tail is still being copied."""
    bank, bl = synth.make_clustered(20000, 256, 13, 71)
    qs, _ = synth.make_clustered(9000, 256, 13, 72)
    a = KNeighborsClassifierB200(10).fit(bank, bl.numpy()).predict(qs.numpy())
    b = KNeighborsClassifierB200(10, use_graph=False).fit(bank, bl.numpy()).predict(qs.numpy())
    assert a.dtype == np.int64 and a.shape == (9000,)
    np.testing.assert_array_equal(a, b)
    c = KNeighborsClassifierB200(10).fit(bank, bl.numpy()).predict(qs.pin_memory())
    np.testing.assert_array_equal(c, b)
