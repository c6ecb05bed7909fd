#!/bin/bash
# 8-GPU call (round 2): charged 8 x wall time, so only what decides the default exchange and proves parity at 8.
#   1. tools/mg_check time on a config-4-sized matrix: default (push + dealt hot labels) | hot off | NCCL broadcasts | 4x2 grid
#   2. the bench line as the driver launches it (torchrun): state digest must equal the 1-GPU digest of the same step count
#   3. two parity tests at 8 GPUs
mkdir -p gpurun_out
G=${1:-8}
L=gpurun_out/r2_i_mgtime${G}.log
: > $L
SIZE="50000000 50000000 1473000000 16 2147483647 10"
ab() { name=$1; shift; echo "== $name ($*)" >> $L; env "$@" timeout 150 tools/mg_check time $G $SIZE 2>&1 | grep -v "^NCCL version" >> $L; }
ab push_hot BLK_NOP=1
ab push_nohot BLK_HOT_SINGLE_ONLY=1
ab nccl_hot BLK_EXCHANGE=nccl
ab grid BLK_EXPERIMENTAL=1 BLK_GRID=auto
cat $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus $G --steps 10 --warmup 3 --sweep-budget-s 15 > gpurun_out/r2_i_bench${G}.json 2> gpurun_out/r2_i_bench${G}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_i_bench${G}.json").read().strip().splitlines()[-1])
    print("bench", d["n_gpus"], round(d["value"], 2), "it/s", {k: round(v, 2) for k, v in d["phases_ms_per_step"].items()},
          d["state_sha256"][:16], "e2e", round(d["e2e"]["value"], 2), d["e2e"]["seconds"], "parity", d.get("parity"))
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2_i_bench${G}.err").read()[-1500:])
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -q --timeout=250 -x \
  -k "(group_context and hot-${G}) or (block_grid and 4x2) or (sharded and ${G}])" > gpurun_out/r2_i_pytest${G}.log 2>&1
tail -4 gpurun_out/r2_i_pytest${G}.log
