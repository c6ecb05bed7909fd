#!/bin/bash
# 1-GPU call at the end of round 2: the full GPU suite, the driver's bench line, the ncu launch list of the same command and one
# ncu --set full capture of the two k_spmv launches of an iteration (each only after its command has exited 0 without ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r2_z_pytest.log 2>&1; tail -4 gpurun_out/r2_z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_z_smoke.log 2>&1; tail -3 gpurun_out/r2_z_smoke.log
timeout 900 python bench.py > gpurun_out/r2_z_bench1.json 2> gpurun_out/r2_z_bench1.err
rc=$?
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_z_bench1.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 3), "it/s", {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["state_sha256"][:16],
          "e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4), "line", round(d["roofline"]["line_model"]["frac"], 3),
          "parity", (d.get("parity") or {}).get("ok"), "small", [round(x["us_per_iter"], 1) for x in d.get("small_configs", [])],
          "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2_z_bench1.err").read()[-1500:])
PY
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_spmv|k_dots|k_ortho|k_small' -s 18 -c 12 --csv \
      --log-file gpurun_out/r2_z_launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_z_ncu_launch.log 2>&1
  tail -14 gpurun_out/r2_z_launches_cfg4.csv | cut -c1-200
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_spmv<' -s 6 -c 2 -o gpurun_out/r2_z_spmv_cfg4 -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_z_ncu_full.log 2>&1
  ls -la gpurun_out/r2_z_spmv_cfg4.ncu-rep
fi
