#!/bin/bash
# 2-GPU call: column bands under sharding -- parity in both process models
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q --timeout=300 -x -k "bands" > gpurun_out/r2_r_pytest_multi.log 2>&1
tail -6 gpurun_out/r2_r_pytest_multi.log
