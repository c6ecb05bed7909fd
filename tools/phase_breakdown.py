"""Per-phase device time of one iteration on BASELINE configs 1-3 (dev tool; profiling mode = eager
launches bracketed by CUDA events, so launch gaps are visible too)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blk_lanczos_b200 as B
for k in (1, 2, 3):
    M, a = B.synth.baseline_config(k)
    p, n, right = a["p"], a["n"], a["right"]
    N = M.ncols if right else M.nrows
    ctx = B.BlockLanczos(M.reduced(p), n=n, prime=p, right=right)
    v0 = B.synth.reference_start_block(N * n, p)
    ctx.set_state(v0); ctx.iterate(32)
    ctx.set_profiling(True)
    t0 = time.perf_counter(); it, _ = ctx.iterate(500); dt = time.perf_counter() - t0
    ph = ctx.phase_times()
    print(f"cfg{k} n={n}: eager {dt/500*1e6:.1f} us/iter; device us/iter per phase: " +
          " ".join(f"{q}={v['ms']/500*1e3:.1f}" for q, v in ph.items()), ctx.info()["chunk_len"], flush=True)
    ctx.set_profiling(False)
    ctx.set_state(v0); ctx.iterate(64)
    t0 = time.perf_counter(); ctx.iterate(2048); dt = time.perf_counter() - t0
    print(f"      graph: {dt/2048*1e6:.1f} us/iter", flush=True)
    ctx.close()
