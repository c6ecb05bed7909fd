#!/bin/bash
# 2-GPU call: k_push_bulk (bulk-copy engine) against k_push_rows -- parity, then timing on a 20M-row matrix
mkdir -p gpurun_out
L=gpurun_out/r2_l_mgtime2.log
: > $L
timeout 500 python -m pytest tests/test_gpu_multi.py -q --timeout=300 -x -k "bulk" > gpurun_out/r2_l_pytest_multi.log 2>&1
tail -6 gpurun_out/r2_l_pytest_multi.log
SIZE="20000000 20000000 600000000 16 2147483647 10"
ab() { name=$1; shift; echo "== $name ($*)" >> $L; env "$@" timeout 120 tools/mg_check time 2 $SIZE 2>&1 | grep -v "^NCCL version" >> $L; }
ab pushk_rows BLK_PUSH_AV=kernel
ab pushk_bulk16 BLK_PUSH_AV=kernel BLK_PUSH_COPY=bulk
ab pushk_bulk4 BLK_PUSH_AV=kernel BLK_PUSH_COPY=bulk BLK_PUSH_CTAS=4
ab pushk_bulk32 BLK_PUSH_AV=kernel BLK_PUSH_COPY=bulk BLK_PUSH_CTAS=32
ab push_bulk16 BLK_PUSH_COPY=bulk
ab ce BLK_EXCHANGE=ce
cat $L
