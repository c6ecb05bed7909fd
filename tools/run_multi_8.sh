#!/bin/bash
# 8-GPU measurement call (round 2): exchange-mode A/B on config 4 through bench.py (one process per GPU, the
# driver's launch), a subset of the multi-GPU parity tests, then the full bench line.
mkdir -p gpurun_out
G=${1:-8}
nvidia-smi topo -m > gpurun_out/r2_f_topo.txt 2>&1
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $G --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_f_ab_${name}.json 2> gpurun_out/r2_f_ab_${name}.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_f_ab_{name}.json").read().strip().splitlines()[-1])
    ph = {k: round(v, 2) for k, v in d["phases_ms_per_step"].items()}
    print(f"{name:14s} {d['value']:7.2f} it/s  {d['ms_per_step']:6.2f} ms  e2e {d['e2e']['value']:6.2f}  {ph}  sha {d['state_sha256'][:12]}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run push
run nccl BLK_EXCHANGE=nccl
run pushk BLK_PUSH_AV=kernel
run push_p8 BLK_PIECES=8
run push_c16 BLK_PUSH_CTAS=16
run push_c64 BLK_PUSH_CTAS=64
run pushk_p8 BLK_PUSH_AV=kernel BLK_PIECES=8
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout=240 -x \
  -k "(group_context and (8-push or 8-nccl or 4-pieces8)) or (sharded and 8) or (block_grid and (4x2 or 2x2)) or (arrival and 4-3) or cli_on_several" \
  > gpurun_out/r2_f_pytest8.log 2>&1
tail -4 gpurun_out/r2_f_pytest8.log
