"""Micro-benchmark of the simtopk kernel alone (CUDA events), optionally with the candidate
emission disabled (plan.reserved bit 0) to expose the pure tcgen05 contraction throughput."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hcir_b200
from hcir_b200 import _lib
from hcir_b200.engine import l2_normalize

def run(nq, ng, d, kc, flags, nsplit=None, iters=10):
    lib = _lib.load()
    q = torch.randn(nq, d, device="cuda"); g = torch.randn(ng, d, device="cuda")
    _, qbf, _ = l2_normalize(q, want_f32=False, want_delta=False)
    _, gbf, _ = l2_normalize(g, want_f32=False, want_delta=False)
    ld = qbf.shape[1]
    plan = _lib.Plan()
    _lib.check(lib.hcir_simtopk_plan(nq, ng, ld, kc, 148, plan))
    plan.flags = flags
    ws = torch.empty(int(plan.bytes), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(lib.hcir_simtopk(qbf.data_ptr(), nq, gbf.data_ptr(), ng, ld, plan, ws.data_ptr(), st))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        _lib.check(lib.hcir_simtopk(qbf.data_ptr(), nq, gbf.data_ptr(), ng, ld, plan, ws.data_ptr(), st))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    tf = 2.0 * nq * ng * ld / ms / 1e9
    print(f"nq={nq} ng={ng} d={d} kc={kc} flags={flags} nsplit={plan.nsplit}: {ms:.3f} ms  {tf:.1f} TFLOP/s  "
          f"gallery {ng*ld*2/ms/1e6:.0f} GB/s", flush=True)

if __name__ == "__main__":
    for flags in (1, 0):
        run(10000, 200000, 768, 104, flags)
        run(4096, 1000000, 768, 264, flags)
        run(64, 2000000, 768, 104, flags)
        run(16384, 200000, 2048, 464, flags)
