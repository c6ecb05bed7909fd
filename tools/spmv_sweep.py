"""BASELINE.json configs[4]: SpMV-only sweep, n in {1,2,4,8,16,32} x p in {65537, 2^31-1}, both
products, on the config-4 matrix (or a smaller twin with --rows).  Device-timed (CUDA events in
blk_time_spmv), >= 10 repetitions after a warm-up launch."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blk_lanczos_b200 as B
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/spmv_sweep.json")
ap.add_argument("--band", type=int, default=0, help="columns within +-band of the row (locality test); 0 = uniform")
ap.add_argument("--ns", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32])
a = ap.parse_args()
dev = torch.device("cuda:0")
w = dict(rows=a.rows, cols=a.rows, mean=30.0)
rows, cols, vals, nnz = bench.gen_device_coo(torch, w, dev)
if a.band:
    off = torch.randint(-a.band, a.band + 1, (nnz,), device=dev, dtype=torch.int32)
    cols = torch.clamp(rows + off, 0, a.rows - 1).to(torch.int32)
    del off
torch.cuda.synchronize()
peak, _ = bench.peaks()
res = []
for p in (65537, 2147483647):
    for n in a.ns:
        ctx = B.BlockLanczos(n=n, prime=p, right=False,
                             device_coo=(a.rows, a.rows, nnz, rows.data_ptr(), cols.data_ptr(), vals.data_ptr()))
        v0 = torch.randint(0, p, (a.rows * n,), dtype=torch.int64).to(torch.int32).numpy().view("uint32")
        ctx.set_state(v0)
        for tr in (True, False):
            ms = ctx.time_spmv(tr, a.reps)
            R = C_ = a.rows
            comp = 8 * nnz + 4 * (R + 1) + 4 * n * C_ + 4 * n * R
            gath = nnz * (8 + 4 * n) + 4 * (R + 1) + 4 * n * R
            line = nnz * (8 + 128) + 4 * n * R
            rec = dict(p=p, n=n, transpose=tr, ms=ms, gnnzn_per_s=nnz * n / ms / 1e6,
                       algorithmic_GBs=comp / ms / 1e6, algorithmic_frac=comp / ms / 1e6 / peak,
                       gather_model_GBs=gath / ms / 1e6, gather_model_frac=gath / ms / 1e6 / peak,
                       line_model_GBs=line / ms / 1e6, line_model_frac=line / ms / 1e6 / peak,
                       ggathers_per_s=nnz / ms / 1e6)
            res.append(rec)
            print(json.dumps(rec), flush=True)
        ctx.close()
json.dump(dict(rows=a.rows, nnz=nnz, band=a.band, peak_GBs=peak, results=res), open(a.out, "w"), indent=1)
