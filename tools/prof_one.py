"""One simtopk main-pass launch sequence on C2 shapes for ncu (tools/prof_one.py [nq ng d kc flags])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_gemm import run
a = [int(x) for x in sys.argv[1:]] or [10000, 200000, 768, 104, 4]
run(a[0], a[1], a[2], a[3], a[4], iters=1)
