"""Phase clocks of the persistent loop kernel on BASELINE configs 1-3 (BLK_LOOP_PROF=1: block 0's SM-cycle counters,
each phase including the grid barrier that ends it) next to the event-timed phases of the kernel chain."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blk_lanczos_b200 as B

out = []
for mode in ("coop", "chain"):
    os.environ["BLK_LOOP"] = "auto" if mode == "coop" else "graph"
    os.environ["BLK_LOOP_PROF"] = "1"
    for k in (1, 2, 3):
        M, a = B.synth.baseline_config(k)
        p, n, right = a["p"], a["n"], a["right"]
        N = M.ncols if right else M.nrows
        with B.BlockLanczos(M.reduced(p), n=n, prime=p, right=right) as ctx:
            v0 = np.random.default_rng(5 + k).integers(0, p, size=N * n).astype(np.uint32)
            ctx.set_state(v0)
            ctx.iterate(64)
            ctx.set_profiling(True)
            it0 = ctx.iterate(0)[0] if False else 64
            iters = 2000
            ctx.iterate(iters)
            ph = ctx.phase_times()
            info = ctx.info()
            out.append({"config": k, "mode": mode, "loop_mode": info["loop_mode"], "tiles": info["tiles"], "chunk_len": info["chunk_len"],
                        "us_per_iter": {name: v["ms"] / iters * 1e3 for name, v in ph.items()}})
print(json.dumps(out))
