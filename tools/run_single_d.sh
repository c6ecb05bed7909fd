#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/band_sweep.py > gpurun_out/r2_q_band_sweep.log 2>&1; cat gpurun_out/r2_q_band_sweep.log | tail -10
