import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import hifir_b200 as hb
from bench import factorize, make_problem
from hifir_b200 import problems as P
A = make_problem("poisson", 128); n = A[0]
M = factorize(A, threads=os.cpu_count())
G = hb.GpuHif(M.levels()); G.set_matrix(A)
bk = P.csr_matvec(A, np.ones(n))
for rep in range(4):
    t0 = time.perf_counter(); x, flag, it, nmv = G.fgmres(bk); dt = time.perf_counter() - t0
    print(os.environ.get("HIFIR_B200_GRAPH"), os.environ.get("HIFIR_B200_MGS_FUSED"), "run", rep, round(dt * 1e3, 1), "ms", it, nmv, flush=True)
