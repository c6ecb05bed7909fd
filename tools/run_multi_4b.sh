#!/bin/bash
# 4-GPU call (round 2): the driver's bench launch at N = 4 and N = 2 with the final defaults (digests must equal the 1-GPU digest
# 028f9b3f858526ce... for --steps 10 --warmup 3), plus the world-4 parity tests of the dealt hot labels and the bulk push
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q --timeout=250 -x -k "hot-4 or bulk-4 or (sharded and 4])" > gpurun_out/r2_n_pytest4.log 2>&1
tail -3 gpurun_out/r2_n_pytest4.log
for G in 4 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2971$G \
    bench.py --gpus $G --steps 10 --warmup 3 --sweep-budget-s 8 > gpurun_out/r2_n_bench$G.json 2> gpurun_out/r2_n_bench$G.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_n_bench$G.json").read().strip().splitlines()[-1])
    print("bench", d["n_gpus"], round(d["value"], 2), "it/s", {k: round(v, 2) for k, v in d["phases_ms_per_step"].items()},
          d["state_sha256"][:16], "e2e", round(d["e2e"]["value"], 2), "parity", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2_n_bench$G.err").read()[-1500:])
PY
done
