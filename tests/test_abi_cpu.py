"""CPU: the C-ABI library builds/loads, exports every symbol include/hcir_b200.h declares,
argument validation works without a GPU, and compute fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import hcir_b200
from hcir_b200 import _lib, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    lib = _lib.load()
    assert os.path.exists(_build.LIB_PATH)
    assert lib.hcir_abi_version() == 5
    assert lib.hcir_padded_dim(768) == 768 and lib.hcir_padded_dim(512) == 512
    assert lib.hcir_padded_dim(100) == 128 and lib.hcir_padded_dim(2048) == 2048


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "hcir_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hcir_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hcir_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_sass_is_blackwell_native():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing: simtopk is not a tcgen05/TMA kernel"
    assert "HGMMA" not in sass


def _poisson_tail(lam, j):
    """P(Poisson(lam) >= j)"""
    import math
    return 1.0 - sum(math.exp(-lam) * lam ** i / math.factorial(i) for i in range(j))


def test_plan_is_consistent():
    lib = _lib.load()
    p = _lib.Plan()
    for nq, ng, ld, kc in [(10000, 200000, 768, 104), (1, 5000, 768, 104), (64, 10_000_000, 768, 104),
                           (16384, 1_250_000, 2048, 464), (129, 257, 64, 74), (4096, 1_000_000, 768, 264)]:
        assert lib.hcir_simtopk_plan(nq, ng, ld, kc, 148, p) == 0
        tiles = -(-ng // 256)
        tps = -(-tiles // p.nsplit)
        assert -(-tiles // tps) == p.nsplit, "empty split"
        assert p.cap >= p.kc + 64 and p.cap % 32 == 0
        assert p.cap <= max(2 * p.kc + 64, 8 * p.kc) or nq * p.nlists * p.cap * 8 <= (2 << 30) + 4096 * nq * p.nlists
        assert p.nlists == p.nsplit  # one list per (query, split): kColHalves == 1
        assert p.bytes >= p.keys_off + nq * p.nlists * p.cap * 8
        assert p.keys_off % 256 == 0 and p.cmax_off % 256 == 0
        if p.sample_rows:
            assert p.sample_rows % 256 == 0 and p.chunk_w in (8, 16, 32)
            assert p.num_chunks * p.chunk_w == p.sample_rows
            assert (p.sample_rows - 1) * p.sample_stride < ng and p.sample_rows * 4 <= ng + 1024
            # the main-pass threshold is the thr_rank-th largest chunk maximum: never looser than the
            # deterministic kc-th, never more than the chunks there are; the sample holds ~2 of the top-kc rows
            assert 1 <= p.hint_rank <= p.thr_rank <= min(p.kc, p.num_chunks - 1)
            lam = p.kc * p.sample_rows / ng
            assert p.thr_rank == p.kc or (lam <= 20 and _poisson_tail(lam, p.thr_rank) <= 1e-7 < _poisson_tail(lam, p.thr_rank - 1))
            st = -(-(p.sample_rows // 256) // p.sample_nsplit)
            assert -(-(p.sample_rows // 256) // st) == p.sample_nsplit
        else:
            assert ng < 64 * p.kc
    assert lib.hcir_simtopk_plan(0, 10, 64, 10, 148, p) == _lib.HCIR_EINVAL
    assert "bad shape" in _lib.last_error()


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # ld must equal padded dim -> EINVAL before any CUDA call
    assert lib.hcir_l2norm_cast(None, 4, 100, 100, None, None, 100, None, None) == _lib.HCIR_EINVAL
    assert lib.hcir_vote(None, None, 4, 0, 3, 0.0, None, None, None) == _lib.HCIR_EINVAL
    assert lib.hcir_merge_topk(None, None, None, 0, 4, 5, None, None, None, None) == _lib.HCIR_EINVAL
    with pytest.raises(ValueError):
        _lib.check(lib.hcir_exact_topk(None, None, 100, 10, 5, 0, None, 1, None, None, None, 0, 148, None), "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.hcir_device_supported() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hcir_b200.GalleryBank(np.random.randn(16, 8).astype(np.float32))
    with pytest.raises(RuntimeError):
        hcir_b200.KNeighborsClassifierB200(3).fit(np.random.randn(16, 8).astype(np.float32), np.arange(16))
    # a well-formed compute call on a box without a B200 must fail, not compute
    x = np.zeros((4, 64), np.float32)
    rc = lib.hcir_l2norm_cast(x.ctypes.data, 4, 64, 64, None, None, 64, None, None)
    assert rc in (_lib.HCIR_ECUDA, _lib.HCIR_EARCH)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hair-centric-image-retrieval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+sklearn", src, flags=re.M), f


def test_shard_plan():
    from hcir_b200 import ShardPlan
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (64, 2), (10_000_001, 8)]:
        sp = ShardPlan(n, w)
        assert sp.start(0) == 0 and sp.stop(w - 1) == n
        assert sum(sp.size(r) for r in range(w)) == n
        for r in range(w - 1):
            assert sp.stop(r) == sp.start(r + 1)
        for row in {0, n - 1, n // 2, n // 3}:
            r = sp.owner(row)
            assert sp.start(r) <= row < sp.stop(r)


def test_classifier_argument_validation():
    from hcir_b200 import KNeighborsClassifierB200
    with pytest.raises(ValueError):
        KNeighborsClassifierB200(5, metric="euclidean")
    with pytest.raises(ValueError):
        KNeighborsClassifierB200(5, weights="distance")
    with pytest.raises(RuntimeError):
        KNeighborsClassifierB200(5).predict(np.zeros((2, 4), np.float32))


def test_peer_region_layout_and_argument_validation():
    """Host-only arithmetic of the peer-memory exchange (csrc/peer.cu): header + 2 parities x world
    slots, slots 256-byte aligned, push CTA count a pure function of the block size."""
    lib = _lib.load()
    for world, slot in [(2, 1024), (8, 10000 * 20 * 16), (8, 16), (16, 4097)]:
        stride = -(-slot // 256) * 256
        assert lib.hcir_peer_region_bytes(world, slot) == 512 + 2 * world * stride
        seen = set()
        for par in (0, 1):
            for r in range(world):
                off = lib.hcir_peer_slot_offset(world, slot, par, r)
                assert off >= 512 and off % 256 == 0 and off + stride <= lib.hcir_peer_region_bytes(world, slot)
                seen.add(off)
        assert len(seen) == 2 * world
    assert lib.hcir_peer_region_bytes(17, 1024) == 0 and lib.hcir_peer_region_bytes(0, 1024) == 0
    assert lib.hcir_peer_push_ctas(16) == 1 and lib.hcir_peer_push_ctas(1 << 30) == 64
    assert 1 <= lib.hcir_peer_push_ctas(10000 * 8) <= 64
    # null pointers / bad ranks are rejected before any CUDA call
    assert lib.hcir_peer_push(None, 16, None, 2, 0, 16, None, None, None) == _lib.HCIR_EINVAL
    assert lib.hcir_peer_wait(None, 2, None, 0, None) == _lib.HCIR_EINVAL
    assert lib.hcir_peer_merge_vote(None, 2, 4, 3, 1, 1024, None, 0, None, None, None, 0, 0.0, None, None,
                                    None) == _lib.HCIR_EINVAL
    # K3's tail: a vote without labels, a peer tail without a step counter, a slot smaller than the block
    plan = _lib.Plan()
    assert lib.hcir_simtopk_plan(64, 100000, 768, 104, 148, plan) == 0
    one = ctypes.c_void_p(16)   # any non-null pointer: validation happens before any CUDA call
    def k3(tail):
        return lib.hcir_select_rescore(one, one, 768, 64, 100000, 20, 0, plan, one, None, 0.0, 0.0, one, one, one, one,
                                       tail, None)
    t = _lib.Tail()
    t.pred = 16
    assert k3(t) == _lib.HCIR_EINVAL and "vote needs labels" in _lib.last_error()
    t = _lib.Tail()
    t.world, t.rank, t.payload = 2, 0, _lib.PAYLOAD_BLOCK
    assert k3(t) == _lib.HCIR_EINVAL and "step counter" in _lib.last_error()
    t.step, t.slot_bytes = 16, 64
    assert k3(t) == _lib.HCIR_EINVAL and "smaller than" in _lib.last_error()
    t = _lib.Tail()
    t.world, t.rank = 2, 2
    assert k3(t) == _lib.HCIR_EINVAL and "bad tail rank" in _lib.last_error()
    assert lib.hcir_vote_idx(None, None, None, 4, 0, 4, 0, 3, 0.0, None, None, None, None) == _lib.HCIR_EINVAL


def test_pending_step_defers_the_check_and_redoes_once():
    from hcir_b200.engine import PendingStep

    class Ev:
        def __init__(self):
            self.synced = 0

        def synchronize(self):
            self.synced += 1

    ev, calls = Ev(), []
    ok = PendingStep(ev, [0], lambda f: f[0] > 0, "fast", lambda: calls.append(1) or "redone")
    assert ev.synced == 0                       # nothing is looked at before result()
    assert ok.result() == "fast" and ok.result() == "fast" and ev.synced == 1 and not ok.redone and not calls
    ev2 = Ev()
    bad = PendingStep(ev2, [3], lambda f: f[0] > 0, "fast", lambda: calls.append(1) or "redone")
    assert bad.result() == "redone" and bad.result() == "redone" and bad.redone and calls == [1]
    done = PendingStep(None, None, None, "sync", None)   # a step that completed synchronously
    assert done.result() == "sync" and not done.redone
