"""CPU: the 'next' rows of SURVEY.md section 8f that are host logic -- result / gallery file formats
and the retrieval metrics -- against restatements of the reference's own loops."""
import json
import os

import numpy as np
import torch

from hcir_b200 import formats, metrics


def _reference_metrics(retrieved, gt_lists, Ks=(10, 20, 50)):
    """experiments/DualViewHair/scripts/quantitative_eval.py:195-209,228-232, loop for loop."""
    from collections import defaultdict
    recall_at_k, ap_at_k, total = defaultdict(int), defaultdict(list), 0
    for row, gt_list in zip(retrieved, gt_lists):
        for k in Ks:
            top_k_preds = list(row[:k])
            if any(gt in top_k_preds for gt in gt_list):
                recall_at_k[k] += 1
            hits, sum_precisions = 0, 0
            for i, p in enumerate(top_k_preds):
                if p in gt_list:
                    hits += 1
                    sum_precisions += hits / (i + 1)
            ap_at_k[k].append(sum_precisions / min(len(gt_list), k) if gt_list else 0.0)
        total += 1
    return {"mAP": {k: sum(ap_at_k[k]) / len(ap_at_k[k]) if ap_at_k[k] else 0 for k in Ks},
            "Recall": {k: recall_at_k[k] / total if total > 0 else 0 for k in Ks}, "total_queries": total}


def test_recall_ap_matches_reference_loop():
    rng = np.random.default_rng(0)
    q, n = 200, 500
    idx = np.stack([rng.permutation(n)[:50] for _ in range(q)])
    gts = [list(rng.choice(n, size=rng.integers(0, 12), replace=False)) for _ in range(q)]
    gts[0] = []                                 # empty ground truth -> AP 0
    gts[1] = list(idx[1][:3])                   # guaranteed hits at ranks 1-3
    ours = metrics.recall_ap_at_k(torch.from_numpy(idx), gts)
    ref = _reference_metrics(idx.tolist(), [[int(x) for x in g] for g in gts])
    assert ours["total_queries"] == ref["total_queries"] == q
    for k in (10, 20, 50):
        assert abs(ours["Recall"][k] - ref["Recall"][k]) < 1e-12
        assert abs(ours["mAP"][k] - ref["mAP"][k]) < 1e-12


def test_top100_json_wire_format(tmp_path):
    paths = [f"/data/db/{i:05d}_hair.png" for i in range(300)]
    idx = np.arange(150)[None, :].repeat(2, 0)
    idx[1] = idx[1][::-1]
    recs = formats.top100_records(["/q/00007_hair.png", "/q/00009_hair.png"], idx, paths)
    assert recs[0]["query"] == "00007_hair.png" and len(recs[0]["top100"]) == 100
    assert recs[0]["top100"][:2] == ["00000_hair.png", "00001_hair.png"]
    out = tmp_path / "log_json" / "top100.json"
    formats.write_top100_json(str(out), recs)
    raw = json.load(open(out))
    assert isinstance(raw, list) and set(raw[0]) == {"query", "top100"}  # what the Visualizer loads
    back = formats.read_top100_json(str(out))
    assert back["00009_hair.png"][0] == "00149_hair.png"


def test_embeddings_npy_roundtrip(tmp_path):
    emb = np.random.default_rng(1).standard_normal((37, 16)).astype(np.float32)
    paths = [f"img/{i}.jpg" for i in range(37)]
    formats.save_embeddings(str(tmp_path), emb, paths)
    assert sorted(os.listdir(tmp_path)) == ["embeddings.npy", "image_paths.txt"]
    np.testing.assert_array_equal(np.load(tmp_path / "embeddings.npy"), emb)
    assert formats.load_paths(str(tmp_path)) == paths


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) runs on the host cores alone and prints ONE JSON
    line with the contract's keys; the B200 arm's helpers import without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "1", "--warmup", "0", "--cpu-sample", "64"], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]
