#!/bin/bash
# one ncu --set full capture of the two k_spmv launches of an iteration (config 4, 1 GPU); the same command exited 0 without ncu before
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_spmv$' -s 6 -c 2 -o gpurun_out/r2_z_spmv_cfg4 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_z_ncu_full.log 2>&1
tail -3 gpurun_out/r2_z_ncu_full.log; ls -la gpurun_out/r2_z_spmv_cfg4.ncu-rep
