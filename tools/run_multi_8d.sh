#!/bin/bash
# 8-GPU call (round 2): the bulk-copy push (k_push_bulk) against k_push_rows; references: push 14.09, push_p8 13.82 ms per iteration
mkdir -p gpurun_out
G=${1:-8}
L=gpurun_out/r2_m_mgtime${G}.log
: > $L
SIZE="50000000 50000000 1473000000 16 2147483647 10"
ab() { name=$1; shift; echo "== $name ($*)" >> $L; env "$@" timeout 120 tools/mg_check time $G $SIZE 2>&1 | grep -v "^NCCL version" >> $L; }
ab pushk_bulk_148x2x16k BLK_PUSH_AV=kernel BLK_PIECES=8 BLK_PUSH_COPY=bulk BLK_PUSH_CTAS=148 BLK_BULK_STAGES=2 BLK_BULK_CHUNK=16384
ab pushk_bulk_64x4x32k BLK_PUSH_AV=kernel BLK_PIECES=8 BLK_PUSH_COPY=bulk BLK_PUSH_CTAS=64
ab push_bulk_148x2x16k BLK_PIECES=8 BLK_PUSH_COPY=bulk BLK_PUSH_CTAS=148 BLK_BULK_STAGES=2 BLK_BULK_CHUNK=16384
ab ce_p8 BLK_EXCHANGE=ce BLK_PIECES=8
cat $L
