"""Quick device-side SpMV timing on a synthetic power-law matrix generated on the GPU (dev tool)."""
import argparse, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blk_lanczos_b200 as B

def gen(N, Mc, mean, seed, dev):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    u = torch.rand(N, device=dev, generator=g, dtype=torch.float64)
    dmin = max(1, round(mean / 3))
    d = torch.clamp((dmin * (1 - u) ** (-1 / 1.5)).floor().to(torch.int64), max=min(1_000_000, 8 * Mc))
    rows = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int32), d)
    nnz = rows.numel()
    cols = torch.randint(0, Mc, (nnz,), device=dev, generator=g, dtype=torch.int32)
    vals = torch.randint(1, 100, (nnz,), device=dev, generator=g, dtype=torch.int32)
    return rows, cols, vals, nnz

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=5_000_000); ap.add_argument("--cols", type=int, default=5_000_000)
ap.add_argument("--mean", type=float, default=30.0); ap.add_argument("--n", type=int, nargs="+", default=[16])
ap.add_argument("--prime", type=int, nargs="+", default=[2147483647]); ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
t0 = time.time(); rows, cols, vals, nnz = gen(a.rows, a.cols, a.mean, 1, dev); torch.cuda.synchronize()
print(f"generated nnz={nnz} in {time.time()-t0:.2f}s", flush=True)
PEAK = 6449.1
for p in a.prime:
    for n in a.n:
        t0 = time.time()
        ctx = B.BlockLanczos(n=n, prime=p, right=False, chunk_len=a.chunk,
                             device_coo=(a.rows, a.cols, nnz, rows.data_ptr(), cols.data_ptr(), vals.data_ptr()))
        torch.cuda.synchronize(); tb = time.time() - t0
        inf = ctx.info()
        v0 = torch.randint(0, p, (a.rows * n,), dtype=torch.int64).to(torch.uint32).numpy()
        ctx.set_state(v0)
        for tr, R, C_ in ((True, a.cols, a.rows), (False, a.rows, a.cols)):
            ms = ctx.time_spmv(tr, a.reps)
            comp = 8 * nnz + 4 * (R + 1) + 4 * n * C_ + 4 * n * R
            gath = nnz * (8 + 4 * n) + 4 * (R + 1) + 4 * n * R
            print(f"p={p} n={n} transpose={int(tr)} Q={inf['chunk_len']} {ms:.3f} ms  {nnz*n/ms/1e6:.2f} Gnnz*n/s  "
                  f"compulsory {comp/ms/1e6:.0f} GB/s ({comp/ms/1e6/PEAK*100:.1f}%)  gather-model {gath/ms/1e6:.0f} GB/s ({gath/ms/1e6/PEAK*100:.1f}%)", flush=True)
        ctx.set_state(v0); ctx.set_profiling(True)
        t0 = time.time(); it, st = ctx.iterate(a.iters); dt = time.time() - t0
        ph = ctx.phase_times()
        print(f"   build {tb:.2f}s  {a.iters} iters in {dt*1e3:.1f} ms -> {it/dt:.2f} it/s; phases(ms/iter): " +
              " ".join(f"{k}={v['ms']/max(1,a.iters):.3f}" for k, v in ph.items()), flush=True)
        ctx.close()
