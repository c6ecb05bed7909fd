#!/bin/bash
# 1-GPU call: accumulate-in-place column bands -- parity, then timing on the config-4 matrix (n = 1, 2, 4, 8)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=500 -k "column_banded" > gpurun_out/r2_s_pytest.log 2>&1; tail -4 gpurun_out/r2_s_pytest.log
timeout 600 python tools/band_sweep.py > gpurun_out/r2_s_band_sweep.log 2>&1; cat gpurun_out/r2_s_band_sweep.log | tail -18
