#!/bin/bash
# 2-GPU call (round 2, second session): dealt degree-sorted labels + hot prefix at world > 1 -- parity, then A/B timing
mkdir -p gpurun_out
L=gpurun_out/r2_h_mgtime2.log
: > $L
timeout 700 python -m pytest tests/test_gpu_multi.py -q --timeout=300 -x -k "hot or (sharded and 2) or cli_on_several or runtime_check" > gpurun_out/r2_h_pytest_multi.log 2>&1
tail -12 gpurun_out/r2_h_pytest_multi.log
SIZE="20000000 20000000 600000000 16 2147483647 10"
ab() { name=$1; shift; echo "== $name ($*)" >> $L; env "$@" timeout 200 tools/mg_check time 2 $SIZE 2>&1 | grep -v "^NCCL version" >> $L; }
ab hot_default BLK_NOP=1
ab hot_off BLK_HOT_SINGLE_ONLY=1
echo "== check (hot on by default at this size)" >> $L
timeout 300 tools/mg_check check 2 20000000 20000000 600000000 16 2147483647 6 2>&1 | tail -8 >> $L
cat $L
