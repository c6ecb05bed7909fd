#!/bin/bash
# 1-GPU call closing round 2: the full GPU suite, smoke(), and the driver's bench line with the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r2_y_pytest.log 2>&1; tail -4 gpurun_out/r2_y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_y_smoke.log 2>&1; tail -3 gpurun_out/r2_y_smoke.log
timeout 900 python bench.py > gpurun_out/r2_y_bench1.json 2> gpurun_out/r2_y_bench1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_y_bench1.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 3), "it/s", {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["state_sha256"][:16],
          "e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4), "line", round(d["roofline"]["line_model"]["frac"], 3),
          "parity", (d.get("parity") or {}).get("ok"), "small", [round(x["us_per_iter"], 1) for x in d.get("small_configs", [])],
          "cpu", d.get("cpu_baseline", {}).get("value"))
    sw = d["spmv_sweep"]
    print("sweep", sw.get("truncated"), sw.get("seconds"), [(p["n"], p["p"], p.get("column_bands"), round(p["M_x"]["ms"], 2), round(p["Mt_x"]["ms"], 2), round(p["M_x"]["frac_algorithmic"], 3)) for p in sw["points"]])
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2_y_bench1.err").read()[-1500:])
PY
