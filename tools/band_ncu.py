"""One banded product (n = 1, config-4 matrix) for an ncu capture: 4 band launches of k_spmv + k_band_combine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import blk_lanczos_b200 as B
w = bench.WORKLOADS["cfg4"]
dev = torch.device("cuda", 0)
rows, cols, vals, nnz = bench.gen_device_coo(torch, w, dev)
torch.cuda.synchronize()
coo = (w["rows"], w["cols"], nnz, rows.data_ptr(), cols.data_ptr(), vals.data_ptr())
ctx = B.BlockLanczos(n=int(os.environ.get("BAND_N", "1")), prime=2147483647, right=False, device=0, device_coo=coo)
print(ctx.info()["bands"], ctx.time_spmv(False, 1))
ctx.close()
