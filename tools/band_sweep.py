"""Column-banded products (SpOp::bands) on the config-4 matrix: ms per product for n in {1, 2, 4}, bands off / default / other
slice sizes (BLK_BANDS, BLK_BAND_BYTES are read at blk_create)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import blk_lanczos_b200 as B

w = bench.WORKLOADS["cfg4"]
dev = torch.device("cuda", 0)
rows, cols, vals, nnz = bench.gen_device_coo(torch, w, dev)
torch.cuda.synchronize()
coo = (w["rows"], w["cols"], nnz, rows.data_ptr(), cols.data_ptr(), vals.data_ptr())
out = []
settings = [("acc48MB", {}), ("acc64MB", {"BLK_BAND_BYTES": str(64 << 20)}), ("acc32MB", {"BLK_BAND_BYTES": str(32 << 20)}), ("partial48MB", {"BLK_BAND_ACC": "0"})]
for n in (1, 2, 4, 8):
    for name, env in settings:
        for k in ("BLK_BANDS", "BLK_BAND_BYTES", "BLK_BAND_ACC"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = time.time()
        ctx = B.BlockLanczos(n=n, prime=2147483647, right=False, device=0, device_coo=coo)
        torch.cuda.synchronize()
        build = time.time() - t0
        rec = {"n": n, "bands": name, "K": ctx.info()["bands"], "build_s": round(build, 2), "device_gb": round(ctx.info()["device_bytes"] / 1e9, 1)}
        for tr in (False, True):
            rec["Mt_x" if tr else "M_x"] = round(ctx.time_spmv(tr, 3), 3)
        ctx.close()
        out.append(rec)
        print(json.dumps(rec), flush=True)
