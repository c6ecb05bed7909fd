#!/bin/bash
# 2-GPU call (last of round 2): the bench line with the banded sweep contexts at full size under sharding
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --sweep-budget-s 35 > gpurun_out/r2_u_bench2.json 2> gpurun_out/r2_u_bench2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_u_bench2.json").read().strip().splitlines()[-1])
    sw = d["spmv_sweep"]
    print("bench", d["n_gpus"], round(d["value"], 2), d["state_sha256"][:16], "parity", (d.get("parity") or {}).get("ok"), "sweep", sw.get("truncated"), sw.get("error"),
          [(p["n"], p["p"], p.get("column_bands"), round(p["M_x"]["ms"], 2), round(p["Mt_x"]["ms"], 2)) for p in sw.get("points", [])])
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2_u_bench2.err").read()[-1500:])
PY
