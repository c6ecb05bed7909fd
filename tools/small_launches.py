"""Eager iterations of one BASELINE small config for an ncu launch list (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blk_lanczos_b200 as B
k = int(sys.argv[1])
M, a = B.synth.baseline_config(k)
p, n, right = a["p"], a["n"], a["right"]
N = M.ncols if right else M.nrows
ctx = B.BlockLanczos(M.reduced(p), n=n, prime=p, right=right, use_graph=0)
ctx.set_state(B.synth.reference_start_block(N * n, p))
ctx.iterate(12)
ctx.close()
