#!/usr/bin/env python
"""bench.py -- queries/sec of the exact cosine top-k / kNN-vote hot path (BASELINE.json metric:
"queries/sec exact top-k cosine kNN (1M-10M x 768 gallery) vs roofline at 1/2/4/8 GPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic queries against the resident,
pre-normalised gallery bank -- ONE CUDA-graph launch per rank: query L2-normalise (K1) -> tcgen05
similarity sample pass + thresholds -> tcgen05 similarity + fused candidate filter (K2) -> fp32
re-score / exact order / certification with the fused tail (K3: neighbour labels, vote, stores into
the peers' memory over NVLink) [-> N>1, gallery shards: fused wait + merge + vote (K5)].

HEADLINE = BASELINE.json configs[2] ("C3"): hair_retrieval top-100 over a 1M x 768 gallery, query
batch 4096 -- inside the metric's 1M-10M x 768 range and the largest retrieval config that the
CPU reference arm can run beside it.  N>1 (torchrun, one rank per GPU) keeps the TOTAL workload
fixed (strong scaling); the partition is named in config.workload: query replicas (every rank holds
the bank and answers a slice of the batch; HCIR_BENCH_SHARD / --shard = query) or gallery row shards
(the north star's layout: local exact top-k, candidate exchange over NVLink peer memory, merge;
= gallery).  auto = hcir_b200.sharded.choose_sharding.

`also` (same JSON line) carries the other driver-visible configurations, each with its own roofline:
  * C4 on the FULL 10M x 768 gallery, 64 queries/step, k=20 + vote -- the HBM-bound streaming regime;
    gallery-sharded for N>1 (the north star's 8-GPU layout), with efficiency_vs_n1;
  * N>1: the headline workload under the OTHER partition;
  * C2 (200k x 768, 10k queries, k=20 kNN vote: uniform = the reference's vote, and T=0.07).

value   = queries/s with inputs resident in HBM (CUDA events, max over ranks)
e2e     = queries/s from pinned HOST queries to host results, every batch's H2D and D2H copies inside
          the timed region (bank fitted once, as the reference builds its bank once): `value` through
          HostPipeline (two host batches in flight: copies on their own streams beside the searches),
          `one_call_at_a_time` through the synchronous reference-facing call (knn_topk /
          KNeighborsClassifierB200.predict / the sharded gallery's topk / predict)
roofline= the dominant kernel (simtopk main pass) timed with CUDA events inside the timed steps:
          tensor bound: 2*Q*N_local*D flops / mean launch duration vs MEASURED_PEAKS.json bf16 burst
          peak; HBM bound (Q below the ridge): (N*D*2 + Q*D*2 + Q*k*12) bytes vs the copy bandwidth
cpu_baseline = the reference's own call sequences on the host cores, bounded query sample: torch.mm +
          topk (qualitative_test.py:79-84) for retrieval workloads, sklearn KNeighborsClassifier
          (classification_engine.py:80-82) for vote workloads, + the Q=1 cosine_similarity + argsort
          call (hair_encoder.py:193-194)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="C1..C5 (synth.CONFIGS); default = C3 = BASELINE configs[2]")
    ap.add_argument("--n", "--gallery-rows", dest="n", type=int, default=None,
                    help="override gallery rows (total); spell it --gallery-rows under torchrun (--n is ambiguous there)")
    ap.add_argument("--q", type=int, default=None, help="override query batch")
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--want", default=None, choices=["topk", "pred"],
                    help="what a step returns: top-k lists (retrieval) or voted labels; default per workload")
    ap.add_argument("--temperature", type=float, default=None, help="T of the weighted vote (default: uniform)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--also", default="auto", choices=["auto", "none", "all"],
                    help="sub-records beside the headline (auto: only for the default workload)")
    ap.add_argument("--shard", default=os.environ.get("HCIR_BENCH_SHARD", "auto"), choices=["auto", "gallery", "query"],
                    help="multi-GPU partition of the HEADLINE: gallery rows (candidate exchange + merge) or query "
                         "replicas; also settable through HCIR_BENCH_SHARD")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="steps in flight for the device-resident measurement (submit / result API); 1 = every "
                         "step waits for its own host-side check before the next is launched")
    ap.add_argument("--pipeline-below-ms", type=float, default=2.0,
                    help="pipeline the submission only when the one-at-a-time step is shorter than this")
    ap.add_argument("--exchange", default=None, choices=["peer", "nccl"],
                    help="multi-GPU result exchange: the library's stores over NVLink peer memory or one NCCL all-gather "
                         "(default: hcir_b200.sharded.DEFAULT_EXCHANGE)")
    ap.add_argument("--k3-width", type=int, default=0, help="force K3's CTA width (1/2/3 = 128/256/1024 threads)")
    return ap.parse_args()


ARGS = parse() if __name__ == "__main__" else None
if ARGS is not None and (ARGS.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) == 1):
    # the CPU legs use every host core: torchrun exports OMP_NUM_THREADS=1, which would throttle the
    # BLAS under torch / sklearn (r1: the reference arm ran 1.8x slower under torchrun)
    _cores = str(os.cpu_count() or 1)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = _cores

import numpy as np  # noqa: E402
import torch  # noqa: E402

DEFAULT_WORKLOAD = "C3"
WANT = {"C1": "pred", "C2": "pred", "C3": "topk", "C4": "pred", "C5": "topk"}


# ---------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi's clocks line, via NVML) during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # NVML missing: report nulls rather than fail the bench
            self.nv = None
        self.period = period_s

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake",
            nv.nvmlClocksEventReasonApplicationsClocksSetting: "applications_clocks_setting",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's call sequences on the host cores
# ---------------------------------------------------------------------------------------------
def host_info():
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    info = {"cores": cores, "torch_threads": torch.get_num_threads()}
    try:
        pi = torch.__config__.parallel_info().splitlines()
        info["torch_parallel_info"] = "; ".join(x.strip() for x in pi if "thread" in x.lower() or "mkl" in x.lower())[:300]
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_info
        info["blas"] = [{"api": t.get("internal_api"), "threads": t.get("num_threads")} for t in threadpool_info()]
    except Exception:
        pass
    return info


class CpuReference:
    """The reference's own call sequences on a query sample of the workload.
      want == "topk":  torch.mm(q, G.t()) + torch.topk (experiments/DualViewHair/scripts/qualitative_test.py:79-84),
                       in query chunks of <= 256 so the [Q, N] fp32 matrix stays near 1 GiB per chunk at N = 1M
      want == "pred":  KNeighborsClassifier(n_neighbors=k, metric="cosine").fit(bank, y).predict(q)
                       (HairPretraining/src/classification_engine.py:80-82)"""

    def __init__(self, cfg, want):
        from hcir_b200 import synth
        from oracle import oracle as O
        self.cfg, self.want, self.O = cfg, want, O
        tag = cfg["tag"]
        bank, bl = synth.make_clustered(cfg["n"], cfg["d"], cfg["classes"], 1234 + tag)
        qs, _ = synth.make_clustered(min(cfg["q"], 16384), cfg["d"], cfg["classes"], 4321 + tag)
        self.bn_t, self.qn_t = O.normalize(bank), O.normalize(qs)
        del bank
        self.bn, self.qn, self.y = self.bn_t.numpy(), self.qn_t.numpy(), bl.numpy()
        self.chunk = max(16, min(256, (1 << 28) // max(1, cfg["n"])))

    def step(self, sample_q: int) -> float:
        t0 = time.perf_counter()
        if self.want == "topk":
            self.O.mm_topk_chunked(self.qn_t[:sample_q], self.bn_t, self.cfg["k"], chunk=self.chunk)
        else:
            from sklearn.neighbors import KNeighborsClassifier
            knn = KNeighborsClassifier(n_neighbors=self.cfg["k"], metric="cosine")
            knn.fit(self.bn, self.y)
            knn.predict(self.qn[:sample_q])
        return time.perf_counter() - t0

    def step_torch_vote(self, sample_q: int) -> float:
        """torch.mm + topk + vote(labels[idx]): the torch CPU kNN path the north star names."""
        t0 = time.perf_counter()
        v, i = self.O.mm_topk_chunked(self.qn_t[:sample_q], self.bn_t, self.cfg["k"], chunk=self.chunk)
        self.O.vote_uniform(self.y[i.numpy()], np.arange(self.cfg["classes"]))
        return time.perf_counter() - t0

    def step_sklearn_kneighbors(self, sample_q: int) -> float:
        from sklearn.neighbors import KNeighborsClassifier
        t0 = time.perf_counter()
        KNeighborsClassifier(n_neighbors=self.cfg["k"], metric="cosine").fit(self.bn, self.y).kneighbors(self.qn[:sample_q])
        return time.perf_counter() - t0

    def step_q1(self) -> float:
        """ONE query: cosine_similarity([q], G)[0] + np.argsort(...)[::-1][:k] (src/models/hair_encoder.py:193-194)."""
        t0 = time.perf_counter()
        self.O.cosine_argsort(self.qn[0], self.bn, self.cfg["k"])
        return time.perf_counter() - t0

    def calibrate(self, budget_s: float) -> int:
        """Query-sample size whose step takes about ``budget_s`` seconds on this host."""
        probe = min(128, self.qn.shape[0])
        self.step(probe)  # page in / thread-pool warm-up
        t = self.step(probe)
        want = int(probe * budget_s / max(t, 1e-6))
        return max(probe, min(self.qn.shape[0], want))

    def describe(self) -> str:
        if self.want == "topk":
            return ("torch.mm(q, G.t()) + torch.topk on the host cores = the reference's retrieval call sequence "
                    f"(qualitative_test.py:79-84), query chunks of {self.chunk}")
        return ("sklearn KNeighborsClassifier(metric='cosine').fit/predict = the reference's own call sequence "
                "(classification_engine.py:80-82)")


def workload_cfg(name, args=None):
    from hcir_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    cfg["tag"] = int(name[1:])
    cfg["name"] = name
    if args is not None:
        if args.n:
            cfg["n"] = args.n
        if args.q:
            cfg["q"] = args.q
        if args.k:
            cfg["k"] = args.k
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    cfg = workload_cfg(name, args)
    want = args.want or WANT[name]
    host = host_info()
    ref = CpuReference(cfg, want)
    # bounded sample: the whole --steps/--warmup run should end within ~2-3 minutes
    sample = args.cpu_sample or ref.calibrate(120.0 / max(1, args.steps + args.warmup))
    times = []
    for i in range(args.warmup + args.steps):
        t = ref.step(sample)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times)) if times else float("nan")
    val = sample / t
    what = "top-%d lists" % cfg["k"] if want == "topk" else "k=%d, uniform vote" % cfg["k"]
    line = {
        "impl": "reference", "metric": "queries/sec exact top-k cosine kNN", "value": val,
        "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['n']}x{cfg['d']} gallery, {what}; each step = {sample} of {cfg['q']} "
                               f"queries against the full gallery (bounded sample: a rate on the same gallery, "
                               f"not the same batch)",
                   "host": host},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": host["cores"], "kind": "port",
                         "sample": f"{sample} queries x full {cfg['n']}-row gallery per step; {ref.describe()}"},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_loop(self, fn, steps, warmup):
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))


def measure(ctx: Ctx, cfg, *, want: str, T, shard: str, steps: int, warmup: int, e2e: bool, probe: bool):
    """One workload on ctx.world GPUs -> the fields of a bench record (value, e2e, roofline, ...)."""
    import hcir_b200
    from hcir_b200 import synth
    from hcir_b200.sharded import QueryShardedGallery, ShardPlan, ShardedGallery

    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    n, d, q, k, C = cfg["n"], cfg["d"], cfg["q"], cfg["k"], cfg["classes"]
    tag = cfg["tag"]
    classes = np.arange(C)
    if world == 1:
        shard = "none"
    # ---- synthetic data, generated on device shard by shard (no network for datasets) ----
    sp = ShardPlan(n, world if shard == "gallery" else 1)
    n_local = sp.size(rank if shard == "gallery" else 0)
    # gallery sharding: every rank synthesises its own row range; query sharding: identical replicas
    bank, bl = synth.make_clustered(n_local, d, C, 1234 + tag + (1000 * rank if shard == "gallery" else 0), device=dev)
    qs, _ = synth.make_clustered(q, d, C, 4321 + tag, device=dev)  # same queries on every rank
    n_plant = min(16, q) if probe else 0
    planted = None
    if n_plant:   # correctness probe: the first queries are scaled copies of rows of rank 0's bank
        planted = torch.arange(n_plant, device=dev) * (n_local // n_plant) + 3
        qs[:n_plant] = bank[planted] * 1.7
        if world > 1:
            ctx.dist.broadcast(qs, src=0)
            ctx.dist.broadcast(planted, src=0)
    labels = bl if want == "pred" else None   # a retrieval gallery carries no labels
    cls = classes if labels is not None else None
    q_local = q if shard != "query" else ShardPlan(q, world).size(rank)
    if shard == "query":
        gal = QueryShardedGallery(bank, labels, device=dev, classes=cls, exchange=args.exchange)
        gb = gal.bank
    elif shard == "gallery":
        gal = ShardedGallery(bank, labels, n_total=n, device=dev, classes=cls, exchange=args.exchange)
        gb = gal.bank
    else:
        gb = hcir_b200.GalleryBank(bank, labels, device=dev, classes=cls)
        gal = None
    del bank
    torch.cuda.empty_cache()
    gb.k3_width = args.k3_width
    sess = gb.session(q, k, T=T, vote=(want == "pred")) if gal is None else None
    if gal is None and sess is None:
        raise SystemExit(f"workload {cfg['name']} is too small for the tensor path on one GPU")
    if sess is not None:
        sess.input.copy_(qs)   # resident input: the producer writes into the step's static buffer

    def step_resident():
        if sess is not None:
            pred, sims, idx = sess.run()
            return pred if want == "pred" else (sims, idx)
        return gal.predict(qs, k, T=T) if want == "pred" else gal.topk(qs, k)

    # ---- device-resident throughput ("value") + clocks ----
    for _ in range(warmup):
        step_resident()
    ctx.barrier()
    graph_sess = sess if sess is not None else gal.last_session
    sampler = ClockSampler(ctx.local)
    l0 = gb.launches
    sampler.start()
    total_ms = ctx.timed_loop(step_resident, steps, 0)
    launches = gb.launches - l0
    sync_ms_per_step = total_ms / steps
    ms_per_step = sync_ms_per_step

    # ---- the same K steps, pipelined: submit step i+1 before looking at step i's host-side check ----
    # (only worth it when the step is short enough for the host round trip to show)
    def submit():
        if sess is not None:
            return sess.submit()
        return gal.submit_predict(qs, k, T=T) if want == "pred" else gal.submit_topk(qs, k)

    can_submit = sess is not None or (gal.exchange == "peer" and (want == "pred" or hasattr(gal, "submit_topk")))
    pipelined = args.pipeline > 1 and sync_ms_per_step < args.pipeline_below_ms and can_submit
    if pipelined:
        import collections

        def pipelined_steps(count):
            pend = collections.deque()
            for _ in range(count):
                pend.append(submit())
                if len(pend) >= args.pipeline:
                    pend.popleft().result()
            while pend:
                pend.popleft().result()

        pipelined_steps(max(3, warmup))
        ctx.barrier()
        l0 = gb.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipelined_steps(steps)
        e1.record()
        ctx.barrier()
        launches = gb.launches - l0
        ms_per_step = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    clocks = sampler.stop()
    value = q / (ms_per_step * 1e-3)
    kernels_per_step = int(getattr(graph_sess, "kernels_per_run", 0)) + (1 if gal is not None else 0)

    # ---- per-kernel durations: the SAME step captured once more with CUDA events between its kernels
    # (the event nodes cost a few microseconds per step, so they are kept out of the timed graphs) ----
    kern = {}
    if gal is None:
        psess = gb.session(q, k, T=T, vote=(want == "pred"), profile=True)
        psess.input.copy_(qs)
    else:
        gal.profile = True
        psess = None
    cur = psess
    for i in range(3 + min(steps, 10)):
        if psess is not None:
            psess.run()
        else:
            step_resident()
            cur = gal.last_session
        if i >= 3 and cur is not None:
            for kname, ms in cur.kernel_ms().items():
                kern.setdefault(kname, []).append(ms)
    if gal is not None:
        gal.profile = False
    ctx.barrier()
    # the dominant kernel's duration on every rank (power-capped GPUs of one box do not run alike;
    # a synchronous sharded step waits for the slowest)
    by_rank = None
    if world > 1 and "simtopk" in kern:
        t = torch.tensor([float(np.mean(kern["simtopk"]))], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        ctx.dist.all_gather(allt, t)
        by_rank = [round(float(x.item()), 4) for x in allt]
    sim_ms = float(np.mean(kern["simtopk"])) if "simtopk" in kern else None
    stats = dict(gb.last_stats)
    if gal is not None:
        stats["exchange"] = gal.exchange
    if stats.get("uncertified"):
        stats["completion"] = dict(gb.retry_stats)
    # ---- correctness probe (outside every timed region): planted rows rank first, lists sorted ----
    probe_out = None
    if n_plant:
        if gal is None:
            s_, i_ = gb.topk(qs[:128], k, return_device=True)
        else:
            s_, i_ = gal.topk(qs[:128 * world] if shard == "query" else qs[:128], k)
        ok_rank1 = int((i_[:n_plant, 0] == planted).sum().item())
        ok_sim = int(((s_[:n_plant, 0] - 1.0).abs() < 1e-5).sum().item())
        srt = bool((s_[:, 1:] <= s_[:, :-1]).all().item())
        probe_out = {"planted": n_plant, "rank1": ok_rank1, "sim_is_1": ok_sim, "sorted": srt}
        if ok_rank1 != n_plant or not srt:
            raise SystemExit(f"bench correctness probe failed: {probe_out}")

    # ---- e2e through the reference-facing call with HOST buffers ----
    e2e_out = None
    if e2e:
        q_host = torch.empty((q, d), dtype=torch.float32, pin_memory=True)
        q_host.copy_(qs)
        torch.cuda.synchronize()
        if gal is None and want == "pred":
            clf = hcir_b200.KNeighborsClassifierB200(n_neighbors=k, metric="cosine", device=dev,
                                                     weights="temperature" if T else "uniform", T=T or 0.07)
            clf.fit_bank(gb)
            fn = lambda: clf.predict(q_host)  # noqa: E731  numpy predictions on the host
            call = "KNeighborsClassifierB200.predict(pinned host queries) -> host int64 labels; bank fitted once"
        elif gal is None:
            fn = lambda: hcir_b200.knn_topk(gb, q_host, k, use_graph=True)  # noqa: E731
            call = "knn_topk(bank, pinned host queries, k) -> host (sims fp32, idx int64); bank built once"
        elif want == "pred":
            fn = lambda: gal.predict(q_host, k, T=T)  # noqa: E731
            call = f"{type(gal).__name__}.predict(pinned host queries) -> host int64 labels on every rank"
        else:
            fn = lambda: gal.topk(q_host, k)  # noqa: E731
            call = f"{type(gal).__name__}.topk(pinned host queries) -> host (sims, idx) on every rank"
        e2e_ms = ctx.timed_loop(fn, steps, max(3, warmup)) / steps
        e2e_out = {"value": q / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q * d * 4,
                   "d2h_bytes_per_step": q * 8 if want == "pred" else q * k * 12, "ms_per_step": e2e_ms, "call": call,
                   "submission": "one call at a time"}
        # the same batches through the host-buffer serving loop: H2D of batch i+1 and D2H of batch i-1 on copy
        # streams beside the search of batch i (every step still copies its queries in and its results out)
        pipe = None
        if args.pipeline > 1:
            try:
                pipe = (hcir_b200.HostPipeline.for_bank(gb, q, k, want=want, T=T, depth=args.pipeline) if gal is None else
                        hcir_b200.HostPipeline.for_gallery(gal, q, k, want=want, T=T, depth=args.pipeline)
                        if gal.exchange == "peer" else None)
            except ValueError:
                pipe = None
        if pipe is not None:
            import collections

            def host_steps(count):
                pend = collections.deque()
                for _ in range(count):
                    pend.append(pipe.submit(q_host))
                    if len(pend) >= args.pipeline:
                        pend.popleft().result()
                while pend:
                    pend.popleft().result()

            host_steps(max(3, warmup))
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            host_steps(steps)
            e1.record()
            ctx.barrier()
            pipe_ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
            one = {"value": e2e_out["value"], "ms_per_step": e2e_ms, "call": call}
            e2e_out.update({"value": q / (pipe_ms * 1e-3), "ms_per_step": pipe_ms, "one_call_at_a_time": one,
                            "call": "HostPipeline.submit(pinned host queries).result() -> host "
                                    + ("int64 labels" if want == "pred" else "(sims fp32, idx int64)")
                                    + (" on every rank" if gal is not None else ""),
                            "submission": f"{args.pipeline} host batches in flight: each batch's H2D copy, search and "
                                          "D2H read-back run on three streams, so copies overlap the neighbouring "
                                          "batches' searches; every batch is copied in and read back inside the timed region"})

    # ---- roofline of the dominant kernel ----
    pk = peaks()
    roof = None
    if sim_ms:
        flops = 2.0 * q_local * n_local * d  # contraction only (this rank's share)
        gbytes = n_local * d * 2.0 + q_local * d * 2.0 + q_local * k * 12.0  # bf16 bank once + queries + results
        # arithmetic intensity of the contraction = q flops per gallery byte; ridge = peak flops / peak bytes
        ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        traffic, traffic_src = profiled_traffic(cfg, world, shard)
        common = {"kernel": "simtopk_kernel<main>", "kernel_ms_by_rank": by_rank,
                  "timed_in": "up to 10 replays of the same step captured with CUDA events between its kernels, right "
                              "after the K timed steps (the timed graphs carry no event nodes)",
                  "traffic": traffic, "traffic_unit": "bytes/launch",
                  "traffic_source": traffic_src, "algorithmic_bytes": gbytes, "algorithmic_flops": flops,
                  "kernel_ms": sim_ms,
                  "share_of_step": sim_ms / sync_ms_per_step,
                  "other_kernels_ms": {kname: float(np.mean(v)) for kname, v in kern.items() if kname != "simtopk"}}
        if q_local < ridge:
            ach = gbytes / (sim_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"],
                    "peak_source": f"{pk['source']} HBM copy bandwidth (MEASURED_PEAKS.json; a read-only stream can "
                                   "exceed a copy)", **common}
        else:
            ach = flops / (sim_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops"],
                    "peak_source": f"{pk['source']} bf16 burst (MEASURED_PEAKS.json)",
                    "frac_of_sustained": ach / pk["bf16_tflops_sustained"] if pk["bf16_tflops_sustained"] else None,
                    **common}

    part = {"gallery": "gallery rows sharded over %d GPUs, candidates exchanged over NVLink peer memory + merge" % world,
            "query": "gallery replicated on %d GPUs, query batch split, results exchanged over NVLink peer memory" % world,
            "none": "1 GPU"}[shard]
    what = f"top-{k} lists" if want == "topk" else f"k={k}, {'T=%g weighted' % T if T else 'uniform'} vote, {C} classes"
    ld = gb.ld
    rec = {
        "value": value, "unit": "queries/s", "ms_per_step": ms_per_step,
        "config": {"workload": f"{cfg['name']}: {n}x{d} gallery ({part}), {q} queries/step, {what}",
                   "partition": shard,
                   "arith": "bf16 tcgen05 contraction (fp32 accumulate) + fp32 re-score of candidates",
                   "l2": f"gallery stream {n_local * ld * 2 / 1e6:.0f} MB bf16 per step > 126 MB L2 (no flush needed)"
                         if n_local * ld * 2 > 126e6 else "gallery fits L2 (small workload)",
                   "path": stats,
                   "submission": (f"{args.pipeline} steps in flight (submit/result: a step's host-side check is "
                                  "read after the next step is launched)") if pipelined else "one step at a time",
                   "ms_per_step_one_at_a_time": sync_ms_per_step,
                   "kernels_per_step": kernels_per_step},
        "e2e": e2e_out, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "probe": probe_out,
    }
    # ---- release everything (collective for the peer channels) before the next workload ----
    if gal is not None:
        gal.close()
    gb.drop_sessions()
    psess = cur = None
    del sess, graph_sess, gb, gal, qs
    torch.cuda.empty_cache()
    ctx.barrier()
    return rec


def profiled_traffic(cfg, world: int, shard: str):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
    committed `ncu --set full` capture of this very workload (profiles/r02_ncu_*.json), or None when no
    capture of this workload / partition exists."""
    if world != 1:
        return None, None
    name = f"r02_ncu_full_{cfg['name'].lower()}_{cfg['n']}x{cfg['d']}_q{cfg['q']}.json"
    path = os.path.join(ROOT, "profiles", name)
    try:
        for krec in json.load(open(path)):
            kn = krec["Kernel Name"]
            if kn.startswith("void simtopk_kernel<0") or "simtopk_kernel<(int)0" in kn:
                unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = krec[key].split()
                    tot += float(v) * unit[u]
                return tot, f"profiles/{name} (ncu --set full, one launch of simtopk_kernel<main>)"
    except (OSError, KeyError, ValueError):
        pass
    return None, None


def n1_reference_point(key: str):
    """The 1-GPU value of an `also` workload from the committed 1-GPU bench line (profiles/), so that an
    N>1 line can state its efficiency; the driver's own N=1 run of the same command is the cross-check."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_n1_points.json")))
        return float(j[key]["value"]), j[key].get("source")
    except (OSError, KeyError, ValueError, TypeError):
        return None, None


def cpu_baseline_for(cfg, want, args, q):
    host = host_info()
    ref = CpuReference(cfg, want)
    sample = args.cpu_sample or ref.calibrate(12.0)
    t = ref.step(sample)
    out = {"value": sample / t, "unit": "queries/s", "cores": host["cores"], "kind": "port",
           "sample": f"{sample} of {q} queries x full {cfg['n']}-row gallery, {t:.2f} s; {ref.describe()} via oracle/",
           "host": host}
    if want == "pred":   # the torch CPU kNN path beside the sklearn one
        s2 = min(ref.qn.shape[0], max(64, sample))
        t2 = ref.step_torch_vote(s2)
        out["torch_mm_topk_vote"] = {"value": s2 / t2, "unit": "queries/s",
                                     "sample": f"{s2} queries, {t2:.2f} s; torch.mm + topk + vote(labels[idx]) "
                                               "(qualitative_test.py:79-84 + sklearn's _mode vote)"}
    else:                # the same neighbour search through sklearn's brute force
        s2 = min(ref.qn.shape[0], 128)
        t2 = ref.step_sklearn_kneighbors(s2)
        out["sklearn_kneighbors"] = {"value": s2 / t2, "unit": "queries/s",
                                     "sample": f"{s2} queries, {t2:.2f} s; KNeighborsClassifier(metric='cosine')"
                                               ".kneighbors (the neighbour search of classification_engine.py:80-82)"}
    t1 = min(ref.step_q1() for _ in range(2))
    out["q1_cosine_argsort"] = {"value": 1.0 / t1, "unit": "queries/s",
                                "sample": f"ONE query per call, {t1:.3f} s; cosine_similarity([q], G)[0] + np.argsort"
                                          "[::-1][:k] on the same gallery (hair_encoder.py:193-194: re-normalises "
                                          "all N rows per call)"}
    return out


def run_b200(args):
    from hcir_b200.sharded import choose_sharding

    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    name = args.workload or DEFAULT_WORKLOAD
    cfg = workload_cfg(name, args)
    want = args.want or WANT[name]
    T = args.temperature
    shard = args.shard if args.shard != "auto" else choose_sharding(cfg["n"], cfg["q"], world, d=cfg["d"])
    head = measure(ctx, cfg, want=want, T=T, shard=shard, steps=args.steps, warmup=args.warmup,
                   e2e=not args.no_e2e, probe=True)

    also = []
    do_also = args.also == "all" or (args.also == "auto" and args.workload is None and not (args.n or args.q or args.k))
    if do_also:
        def sub(label, cfg2, **kw):
            try:
                r = measure(ctx, cfg2, e2e=False, probe=True, **kw)
            except Exception as ex:  # a sub-record must never take the headline down with it
                if world > 1:
                    raise
                r = {"error": f"{type(ex).__name__}: {ex}"}
            r["label"] = label
            return r

        # (i) the streaming regime on the FULL 10M x 768 gallery (north star: gallery sharded over the GPUs)
        c4 = workload_cfg("C4")
        r = sub("C4-10M", c4, want="pred", T=None, shard="gallery", steps=max(20, args.steps), warmup=max(5, args.warmup))
        if "value" in r:
            v1, src = n1_reference_point("C4-10M")
            if world == 1:
                r["efficiency_vs_n1"] = 1.0
            elif v1:
                r["efficiency_vs_n1"] = r["value"] / (world * v1)
                r["n1_value"], r["n1_source"] = v1, src
        also.append(r)
        # (ii) N>1: the headline workload under the other partition
        if world > 1:
            other = "gallery" if shard == "query" else "query"
            also.append(sub(f"{name}-{other}-partition", cfg, want=want, T=T, shard=other, steps=args.steps,
                            warmup=args.warmup))
        # (iii) C2: the kNN-vote configuration (uniform = the reference's vote; T = 0.07 = BASELINE config text)
        c2 = workload_cfg("C2")
        shard2 = choose_sharding(c2["n"], c2["q"], world, d=c2["d"])
        r = sub("C2", c2, want="pred", T=None, shard=shard2, steps=args.steps, warmup=args.warmup)
        rt = sub("C2-T0.07", c2, want="pred", T=0.07, shard=shard2, steps=args.steps, warmup=args.warmup)
        if "value" in r and "value" in rt:
            r["value_T0.07_vote"] = rt["value"]
            r["ms_per_step_T0.07_vote"] = rt["ms_per_step"]
        also.append(r)

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_for(cfg, want, args, cfg["q"])

    if rank == 0:
        line = {
            "metric": "queries/sec exact top-k cosine kNN", "value": head["value"], "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "clocks": head["clocks"], "roofline": head["roofline"], "cpu_baseline": cpu, "probe": head["probe"],
            "also": also,
        }
        print(json.dumps(line))
    if world > 1:
        torch.cuda.synchronize()
        ctx.dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)   # (destroy_process_group was observed to hang after captured NCCL work in r1)


def main():
    args = ARGS
    if os.environ.get("HCIR_DEBUG_HANG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["HCIR_DEBUG_HANG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
