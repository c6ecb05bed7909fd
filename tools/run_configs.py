"""BASELINE.json configs 1-3 end to end on one GPU (dev/measurement tool): device-timed iteration
rate of the whole run (CUDA-graph batches), iterations, and the kernel property M^T v == 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blk_lanczos_b200 as B

out = []
for k in (1, 2, 3):
    M, a = B.synth.baseline_config(k)
    p, n, right = a["p"], a["n"], a["right"]
    Mp = M.reduced(p)
    N = M.ncols if right else M.nrows
    Mc = M.nrows if right else M.ncols
    ctx = B.BlockLanczos(Mp, n=n, prime=p, right=right)
    v0 = B.synth.reference_start_block(N * n, p)
    ctx.set_state(v0); ctx.iterate(64)                # warm-up (graph capture, clocks)
    ctx.set_state(v0)
    t0 = time.perf_counter()
    it, stopped = 0, False
    while not stopped:
        it, stopped = ctx.iterate(4096)
    dt = time.perf_counter() - t0
    st = ctx.get_state(("v", "tmp"))
    ok = bool(st["v"].any()) and not st["tmp"][:Mc * n].any()
    rec = dict(config=k, rows=M.nrows, cols=M.ncols, nnz=M.nnz, n=n, p=p, right=right, iterations=it,
               seconds=dt, iters_per_s=it / dt, us_per_iter=dt / it * 1e6, kernel_ok=ok,
               launches=ctx.kernel_launches())
    print(json.dumps(rec), flush=True)
    out.append(rec)
    ctx.close()
json.dump(out, open(os.path.join("gpurun_out", "configs123.json"), "w"), indent=1)
