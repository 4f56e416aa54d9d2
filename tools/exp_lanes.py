"""Experiment: two SearchSessions of the same shape on two CUDA streams, submitted alternately, so that the
tail of step i (K3: DRAM/latency-bound) overlaps the main pass of step i+1 (tensor-bound).
    python tools/exp_lanes.py C3|C2|C4 [steps] [gallery rows] [lanes]
Prints ms/step for one lane (the bench's `value` path) and for two lanes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hcir_b200  # noqa: E402
from hcir_b200 import synth  # noqa: E402
from hcir_b200.engine import SearchSession  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    cfg = dict(synth.CONFIGS[name])
    if len(sys.argv) > 3:
        cfg["n"] = int(sys.argv[3])      # e.g. C4 60 1250000: one rank's share of the 10M gallery at 8 GPUs
    nlanes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    bank, bl = synth.make_clustered(cfg["n"], cfg["d"], cfg["classes"], 1234, device=dev)
    qs, _ = synth.make_clustered(cfg["q"], cfg["d"], cfg["classes"], 4321, device=dev)
    vote = name != "C3"
    gb = hcir_b200.GalleryBank(bank, bl if vote else None, device=dev)
    del bank
    lanes = [SearchSession(gb, cfg["q"], cfg["k"], vote=vote) for _ in range(nlanes)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nlanes)]
    for s in lanes:
        s.input.copy_(qs)
    torch.cuda.synchronize()

    def run(nl, count):
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for i in range(count):
            ln = i % nl
            with torch.cuda.stream(streams[ln]):
                lanes[ln].graph.replay()
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / count

    for nl in (1, nlanes, 1, nlanes, 2):
        run(nl, 5)
        ms = run(nl, steps)
        print(f"{name} lanes={nl}: {ms:.4f} ms/step  {cfg['q'] / ms * 1e3:,.0f} q/s", flush=True)


if __name__ == "__main__":
    main()
