#!/bin/bash
# 1-GPU call: column-banded products for n_pad <= 4 -- parity, then timing on the config-4 matrix
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=500 -k "column_banded or spmv or relabelled or loop_state" > gpurun_out/r2_p_pytest.log 2>&1; tail -4 gpurun_out/r2_p_pytest.log
timeout 600 python tools/band_sweep.py > gpurun_out/r2_p_band_sweep.log 2>&1; cat gpurun_out/r2_p_band_sweep.log | tail -14
