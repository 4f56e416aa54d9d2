#!/usr/bin/env python
"""bench.py -- queries/sec of the exact cosine top-k + kNN-vote hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic queries against the
resident, pre-normalised gallery bank: query L2-normalise (K1) -> tcgen05 similarity + fused
top-kc (K2) -> fp32 re-score / exact order / certification (K3) -> [N>1: all-gather of the
per-shard exact top-k + merge (K5)] -> neighbour labels + vote (K4).

N=1 workload = BASELINE.json configs[1] ("C2": 200k x 768 gallery, 10k queries, k=20 kNN vote).
N>1 (torchrun, one rank per GPU): the SAME total workload with the gallery row-sharded over the
ranks (strong scaling), one NCCL all-gather of k candidates per query per rank.

value   = queries/s with inputs resident in HBM (CUDA events, max over ranks)
e2e     = queries/s through the reference-facing call KNeighborsClassifierB200.predict(host
          queries) -> host predictions, H2D/D2H copies inside the timed region (bank fitted once,
          as sklearn's fit stores the bank once)
roofline= the dominant kernel (simtopk) timed with CUDA events inside the timed steps:
          2*Q*N_local*D flops / mean launch duration vs MEASURED_PEAKS.json bf16 burst peak
cpu_baseline = the reference's own call sequence (sklearn KNeighborsClassifier(metric="cosine")
          .fit/.predict, classification_engine.py:80-82) on the host cores, bounded query sample
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

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", help="C1..C5 (synth.CONFIGS); default = BASELINE configs[1]")
    ap.add_argument("--n", "--gallery-rows", dest="n", type=int, default=None,
                    help="override gallery rows (total); spell it --gallery-rows under torchrun (--n is ambiguous there)")
    ap.add_argument("--q", type=int, default=None, help="override query batch")
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (no CUDA graph)")
    ap.add_argument("--shard", default="auto", choices=["auto", "gallery", "query"],
                    help="multi-GPU partition: gallery rows (one candidate all-gather + merge) or query replicas")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="steps in flight for the device-resident measurement (submit / result API); 1 = every "
                         "step waits for its own host-side check before the next is launched")
    ap.add_argument("--pipeline-below-ms", type=float, default=1.0,
                    help="pipeline the submission only when the one-at-a-time step is shorter than this")
    ap.add_argument("--exchange", default=None, choices=["peer", "nccl"],
                    help="multi-GPU result exchange: the library's push over NVLink peer memory or one NCCL all-gather "
                         "(default: hcir_b200.sharded.DEFAULT_EXCHANGE)")
    return ap.parse_args()


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
# CPU baseline / reference arm: the reference's call sequence on the host cores
# ---------------------------------------------------------------------------------------------
class CpuReference:
    """KNeighborsClassifier(n_neighbors=k, metric="cosine").fit(bank, y).predict(q)
    (HairPretraining/src/classification_engine.py:80-82) on a query sample of the workload."""

    def __init__(self, cfg):
        from hcir_b200 import synth
        from oracle import oracle as O
        self.cfg = cfg
        tag = cfg["tag"]
        bank, bl = synth.make_clustered(cfg["n"], cfg["d"], cfg["classes"], 1234 + tag)
        qs, _ = synth.make_clustered(min(cfg["q"], 16384), cfg["d"], cfg["classes"], 4321 + tag)
        self.bn, self.qn, self.y = O.normalize(bank).numpy(), O.normalize(qs).numpy(), bl.numpy()

    def step(self, sample_q: int) -> float:
        from sklearn.neighbors import KNeighborsClassifier
        t0 = time.perf_counter()
        knn = KNeighborsClassifier(n_neighbors=self.cfg["k"], metric="cosine")
        knn.fit(self.bn, self.y)
        knn.predict(self.qn[:sample_q])
        return time.perf_counter() - t0

    def calibrate(self, budget_s: float) -> int:
        """Query-sample size whose step takes about ``budget_s`` seconds on this host."""
        probe = min(128, self.qn.shape[0])
        self.step(probe)  # page in / thread-pool warm-up
        t = self.step(probe)
        want = int(probe * budget_s / max(t, 1e-6))
        return max(probe, min(self.qn.shape[0], want))


def profiled_traffic(workload: str, world: int, cfg):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from
    the committed `ncu --set full` capture of this very command (profiles/), or None when no capture
    of this workload / partition exists."""
    if workload != "C2" or world != 1 or cfg != dict(__import__("hcir_b200").synth.CONFIGS["C2"], tag=2):
        return None, None
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    try:
        for k in json.load(open(path)):
            if k["Kernel Name"].startswith("void simtopk_kernel<0"):
                unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = k[key].split()
                    tot += float(v) * unit[u]
                return tot, "profiles/r01_ncu_full_summary.json (ncu --set full, one launch of simtopk_kernel<main>)"
    except (OSError, KeyError, ValueError):
        pass
    return None, None


def workload_cfg(args):
    from hcir_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    cfg["tag"] = int(args.workload[1:])
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
    cfg = workload_cfg(args)
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    ref = CpuReference(cfg)
    # bounded sample: the whole --steps/--warmup run should end within ~2-3 minutes
    sample = args.cpu_sample or ref.calibrate(150.0 / max(1, args.steps + args.warmup))
    times = []
    for i in range(args.warmup + args.steps):
        t = ref.step(sample)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times)) if times else float("nan")
    val = sample / t
    line = {
        "impl": "reference", "metric": "queries/sec exact top-k cosine kNN + vote", "value": val,
        "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['n']}x{cfg['d']} gallery, k={cfg['k']}, uniform vote; "
                               f"each step = {sample} of {cfg['q']} queries (bounded sample)"},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} queries x full {cfg['n']}-row gallery per step; sklearn "
                                   "KNeighborsClassifier(metric='cosine').fit/predict = the reference's own "
                                   "call sequence (classification_engine.py:80-82)"},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    import hcir_b200
    from hcir_b200 import synth
    from hcir_b200.sharded import QueryShardedGallery, ShardPlan, ShardedGallery, choose_sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload_cfg(args)
    n, d, q, k, C = cfg["n"], cfg["d"], cfg["q"], cfg["k"], cfg["classes"]
    tag = cfg["tag"]
    T = None  # uniform vote = the reference's KNeighborsClassifier (reference-parity mode)
    classes = np.arange(C)

    # ---- synthetic data, generated on device shard by shard (no network for datasets) ----
    shard = args.shard if args.shard != "auto" else choose_sharding(n, q, world, d=d)
    if world == 1:
        shard = "none"
    sp = ShardPlan(n, world if shard == "gallery" else 1)
    n_local = sp.size(rank if shard == "gallery" else 0)
    # gallery sharding: every rank synthesises its own row range; query sharding: identical replicas
    bank, bl = synth.make_clustered(n_local, d, C, 1234 + tag + (1000 * rank if shard == "gallery" else 0), device=dev)
    qs, _ = synth.make_clustered(q, d, C, 4321 + tag, device=dev)  # same queries on every rank
    q_local = q if shard != "query" else ShardPlan(q, world).size(rank)
    if shard == "query":
        gal = QueryShardedGallery(bank, bl, device=dev, classes=classes, exchange=args.exchange)
        gb = gal.bank
    elif shard == "gallery":
        gal = ShardedGallery(bank, bl, n_total=n, device=dev, classes=classes, exchange=args.exchange)
        gb = gal.bank
    else:
        gb = hcir_b200.GalleryBank(bank, bl, device=dev, classes=classes)
        gal = None
    del bank
    torch.cuda.empty_cache()

    # one fixed-shape step = one CUDA-graph launch (SearchSession); gallery-sharded mode runs eagerly
    use_graph = not args.no_graph
    sess = gb.session(q, k, T=T, profile=True) if (use_graph and shard == "none") else None
    if gal is not None and use_graph:
        gal.profile = True

    def step_resident():
        if sess is not None:
            pred, _, _ = sess.run(qs)
            return pred
        if gal is not None:
            return gal.predict(qs, k, T=T, mode="auto" if use_graph else "tensor")
        return gb.predict(qs, k, T=T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput ("value") + per-kernel events + clocks ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    graph_sess = sess if sess is not None else (gal.last_session if (gal is not None and use_graph) else None)
    kern = {}
    if graph_sess is None:
        gb.kernel_events = []
    l0 = gb.launches

    def timed_step():
        out = step_resident()
        if graph_sess is not None:  # events recorded inside the graph: read them after every replay
            for name, ms in graph_sess.kernel_ms().items():
                kern.setdefault(name, []).append(ms)
        return out

    sampler.start()
    total_ms = timed_loop(timed_step, args.steps, 0)
    launches = gb.launches - l0
    if graph_sess is None:
        for name, a, b in gb.kernel_events:
            kern.setdefault(name, []).append(a.elapsed_time(b))
        gb.kernel_events = None
    sync_ms_per_step = total_ms / args.steps
    ms_per_step = sync_ms_per_step

    # ---- the same K steps, pipelined: submit step i+1 before looking at step i's host-side check ----
    # (only worth it when the step is short enough for the host round trip to show: a multi-ms,
    # power-capped tensor-bound step gains nothing from losing its idle gaps)
    pipelined = (use_graph and args.pipeline > 1 and sync_ms_per_step < args.pipeline_below_ms
                 and (sess is not None or (gal is not None and gal.exchange == "peer")))
    if pipelined:
        import collections

        def submit():
            return sess.submit(qs) if sess is not None else gal.submit_predict(qs, k, T=T)

        def pipelined_steps(steps):
            pend = collections.deque()
            for _ in range(steps):
                pend.append(submit())
                if len(pend) >= args.pipeline:
                    pend.popleft().result()
            while pend:
                pend.popleft().result()

        pipelined_steps(max(3, args.warmup))
        barrier()
        l0 = gb.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipelined_steps(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = gb.launches - l0
        ms_per_step = float(t.item()) / args.steps
    clocks = sampler.stop()
    value = q / (ms_per_step * 1e-3)
    # the dominant kernel's duration on every rank (power-capped GPUs of one box do not run alike;
    # a synchronous sharded step waits for the slowest)
    by_rank = None
    if world > 1 and "simtopk" in kern:
        t = torch.tensor([float(np.mean(kern["simtopk"]))], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        by_rank = [round(float(x.item()), 4) for x in allt]
    sim_ms = float(np.mean(kern["simtopk"])) if "simtopk" in kern else None
    stats = dict(gb.last_stats)
    if gal is not None:
        stats["exchange"] = gal.exchange
    if stats.get("uncertified"):
        stats["completion"] = dict(gb.retry_stats)

    # ---- e2e through the reference-facing call with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        q_host = torch.empty((q, d), dtype=torch.float32, pin_memory=True)
        q_host.copy_(qs)
        torch.cuda.synchronize()
        if gal is None:
            clf = hcir_b200.KNeighborsClassifierB200(n_neighbors=k, metric="cosine", device=dev,
                                                     use_graph=not args.no_graph)
            clf._bank = gb
            clf.classes_ = gb.classes_
            fn = lambda: clf.predict(q_host)  # noqa: E731  numpy predictions on the host
        else:
            fn = lambda: gal.predict(q_host, k, T=T)  # noqa: E731
        e2e_ms = timed_loop(fn, args.steps, max(3, args.warmup)) / args.steps
        e2e = {"value": q / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q * d * 4,
               "d2h_bytes_per_step": q * 8, "ms_per_step": e2e_ms,
               "call": "KNeighborsClassifierB200.predict(pinned host queries) -> host int64 labels; bank fitted once"}

    # ---- roofline of the dominant kernel ----
    pk = peaks()
    roof = None
    if sim_ms:
        flops = 2.0 * q_local * n_local * d  # contraction only (this rank's share)
        gbytes = n_local * d * 2.0 + q_local * d * 2.0 + q_local * k * 12.0  # bf16 bank once + queries + results
        # arithmetic intensity of the contraction = q flops per gallery byte; ridge = peak flops / peak bytes
        ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        traffic, traffic_src = profiled_traffic(args.workload, world, cfg)
        common = {"kernel": "simtopk_kernel<main>", "kernel_ms_by_rank": by_rank,
                  "timed_in": "the synchronous pass of the K timed steps (CUDA events inside the step's graph)",
                  "traffic": traffic, "traffic_unit": "bytes/launch",
                  "traffic_source": traffic_src, "algorithmic_bytes": gbytes, "algorithmic_flops": flops,
                  "kernel_ms": sim_ms,
                  "share_of_step": sim_ms / ms_per_step,
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

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        ref = CpuReference(cfg)
        sample = args.cpu_sample or ref.calibrate(15.0)
        t = ref.step(sample)
        qps = sample / t
        cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{sample} of {q} queries x full {n}-row gallery, {t:.2f} s; sklearn "
                         "KNeighborsClassifier(metric='cosine').fit/predict (the reference's own call "
                         "sequence, classification_engine.py:80-82) via oracle/"}

    if rank == 0:
        line = {
            "metric": "queries/sec exact top-k cosine kNN + vote", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n}x{d} gallery ({({'gallery': 'rows sharded over %d GPUs, one candidate all-gather + merge' % world, 'query': 'replicated on %d GPUs, queries sharded, one result all-gather' % world, 'none': '1 GPU'})[shard]}), "
                                   f"{q} queries/step, k={k}, uniform vote, {C} classes",
                       "arith": "bf16 tcgen05 contraction (fp32 accumulate) + fp32 re-score of candidates",
                       "l2": f"gallery stream {n_local * gb.ld * 2 / 1e6:.0f} MB bf16 per step > 126 MB L2 (no flush needed)"
                             if n_local * gb.ld * 2 > 126e6 else "L2 flushed? no: gallery fits L2 (small workload)",
                       "path": stats,
                       "submission": (f"{args.pipeline} steps in flight (submit/result: a step's host-side check is "
                                      "read after the next step is launched)") if pipelined else "one step at a time",
                       "ms_per_step_one_at_a_time": sync_ms_per_step},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        # captured NCCL work keeps the communicator busy at teardown (destroy_process_group was
        # observed to hang with live CUDA graphs): drop the graphs, drain, and leave without it
        gb.__dict__.pop("_sessions", None)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if os.environ.get("HCIR_DEBUG_HANG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["HCIR_DEBUG_HANG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
