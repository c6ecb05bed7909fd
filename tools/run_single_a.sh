#!/bin/bash
# 1-GPU call: look-back SpMV validation + A/B, sustained / in-place orthogonalize timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_g_pytest.log 2>&1; tail -4 gpurun_out/r2_g_pytest.log
python tools/small_configs_ab.py > gpurun_out/r2_g_small_lookback.json 2>&1
BLK_SPMV_FIX=kernel python tools/small_configs_ab.py > gpurun_out/r2_g_small_fixkernel.json 2>&1
cut -c1-900 gpurun_out/r2_g_small_lookback.json gpurun_out/r2_g_small_fixkernel.json
python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2_g_bench_lookback.json 2> gpurun_out/r2_g_bench_lookback.err
BLK_SPMV_FIX=kernel python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2_g_bench_fixkernel.json 2> gpurun_out/r2_g_bench_fixkernel.err
python - <<'PY'
import json
for n in ("lookback", "fixkernel"):
    try:
        d = json.loads(open(f"gpurun_out/r2_g_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 3), "it/s", {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["state_sha256"][:16], d["gpu_launches"])
    except Exception as e:
        print(n, "FAILED", e, open(f"gpurun_out/r2_g_bench_{n}.err").read()[-800:])
PY
timeout 300 tools/umma_dense_test ortho 50000000 2>&1 | grep -v "identical" > gpurun_out/r2_g_ortho_sustained.log
cat gpurun_out/r2_g_ortho_sustained.log
