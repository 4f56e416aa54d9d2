"""Micro-benchmark of the simtopk call (sample pass + thresholds + main pass) with CUDA events per
iteration.  flags: HCIR_FLAG_* (1 = emit nothing, 4 = main pass only, 8 = never skip a chunk).
Prints min / median per-iteration time and the SM clock sampled while the loop runs."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hcir_b200
from hcir_b200 import _lib
from hcir_b200.engine import l2_normalize

try:
    import pynvml
    pynvml.nvmlInit()
    _H = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    _H = None

_DATA = {}

def data(nq, ng, d):
    key = (nq, ng, d)
    if key not in _DATA:
        _DATA.clear()
        g = torch.Generator(device="cuda").manual_seed(1)
        q = torch.randn(nq, d, device="cuda", generator=g); b = torch.randn(ng, d, device="cuda", generator=g)
        _, qbf, _ = l2_normalize(q, want_f32=False, want_delta=False, pad_rows_to=128)
        _, gbf, _ = l2_normalize(b, want_f32=False, want_delta=False)
        _DATA[key] = (qbf, gbf)
    return _DATA[key]

def run(nq, ng, d, kc, flags, iters=20):
    lib = _lib.load()
    qbf, gbf = data(nq, ng, d)
    ld = qbf.shape[1]
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(nq, ng, ld, kc, 148, plan))
    plan.q_rows = -(-nq // 128) * 128
    ws = torch.empty(int(plan.bytes), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _lib.check(lib.hcir_simtopk(qbf.data_ptr(), nq, gbf.data_ptr(), ng, ld, plan, ws.data_ptr(), st))
    call()                       # thresholds in the workspace (needed by MAIN_ONLY)
    plan.flags = flags
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    clocks, stop = [], threading.Event()
    def sample():
        while not stop.is_set():
            if _H is not None:
                clocks.append(pynvml.nvmlDeviceGetClockInfo(_H, pynvml.NVML_CLOCK_SM))
            time.sleep(0.002)
    th = threading.Thread(target=sample); th.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); call(); b.record()
    torch.cuda.synchronize(); stop.set(); th.join()
    ts = np.array([a.elapsed_time(b) for a, b in evs])
    tf = lambda ms: 2.0 * nq * ng * ld / ms / 1e9
    clk = int(np.median(clocks)) if clocks else -1
    print(f"nq={nq} ng={ng} d={d} kc={kc} flags={flags} nsplit={plan.nsplit} cap={plan.cap}: min {ts.min():.3f} ms "
          f"({tf(ts.min()):.0f} TF)  med {np.median(ts):.3f} ms ({tf(np.median(ts)):.0f} TF)  "
          f"gallery {ng*ld*2/np.median(ts)/1e6:.0f} GB/s  sm_clk {clk} MHz", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for flags in [int(x) for x in sys.argv[1:]]:
            run(10000, 200000, 768, 104, flags)
    else:
        for flags in (5, 4):
            run(10000, 200000, 768, 104, flags)
            run(4096, 1000000, 768, 264, flags)
            run(64, 2000000, 768, 104, flags)
            run(16384, 200000, 2048, 464, flags)


def ab(nq, ng, d, kc, flag_list, rounds=30):
    """Interleaved A/B/...: one launch of each variant per round, medians over rounds (robust to the
    slow clock drift of a power-capped part)."""
    lib = _lib.load()
    qbf, gbf = data(nq, ng, d)
    ld = qbf.shape[1]
    plans = []
    for f in flag_list:
        plan = _lib.Plan()
        _lib.check(lib.hcir_simtopk_plan(nq, ng, ld, kc, 148, plan))
        plan.q_rows = -(-nq // 128) * 128
        plans.append(plan)
    ws = torch.empty(int(max(p.bytes for p in plans)), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def call(plan):
        _lib.check(lib.hcir_simtopk(qbf.data_ptr(), nq, gbf.data_ptr(), ng, ld, plan, ws.data_ptr(), st))
    call(plans[0])
    for p, f in zip(plans, flag_list):
        p.flags = f
    times = {f: [] for f in flag_list}
    for r in range(rounds + 3):
        for p, f in zip(plans, flag_list):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(p); b.record()
            torch.cuda.synchronize()
            if r >= 3:
                times[f].append(a.elapsed_time(b))
    for f in flag_list:
        t = np.array(times[f])
        print(f"  flags={f:3d}: median {np.median(t):.3f} ms  min {t.min():.3f}  ({2.0*nq*ng*ld/np.median(t)/1e9:.0f} TF)", flush=True)
