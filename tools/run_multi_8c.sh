#!/bin/bash
# 8-GPU call (round 2): A/B of the exchange knobs around the default (push + dealt hot labels = 14.09 ms per iteration)
mkdir -p gpurun_out
G=${1:-8}
L=gpurun_out/r2_j_mgtime${G}.log
: > $L
SIZE="50000000 50000000 1473000000 16 2147483647 10"
ab() { name=$1; shift; echo "== $name ($*)" >> $L; env "$@" timeout 120 tools/mg_check time $G $SIZE 2>&1 | grep -v "^NCCL version" >> $L; }
ab pushk BLK_PUSH_AV=kernel
ab pushk_p8 BLK_PUSH_AV=kernel BLK_PIECES=8
ab push_p8 BLK_PIECES=8
ab pushk_p8_c64 BLK_PUSH_AV=kernel BLK_PIECES=8 BLK_PUSH_CTAS=64
ab pushk_p8_c16 BLK_PUSH_AV=kernel BLK_PIECES=8 BLK_PUSH_CTAS=16
cat $L
