#!/bin/bash
mkdir -p gpurun_out
python tools/band_ncu.py > gpurun_out/r2_t_band_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none -k 'regex:^k_spmv$|k_band_combine' -s 5 -c 5 -o gpurun_out/r2_t_band_n1 -f python tools/band_ncu.py > gpurun_out/r2_t_band_ncu.log 2>&1
tail -2 gpurun_out/r2_t_band_plain.log gpurun_out/r2_t_band_ncu.log; ls -la gpurun_out/r2_t_band_n1.ncu-rep
