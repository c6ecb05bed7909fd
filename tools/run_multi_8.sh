#!/bin/bash
# Multi-GPU measurement call (round 2), sized for a slot that is charged G times its wall time:
#   1. exchange-mode A/B on a config-4-sized matrix with tools/mg_check (one process, one thread per GPU: starts in seconds)
#   2. a few parity checks at G GPUs (group context, P x Q grid, CLI, one torchrun run = CUDA-IPC path)
#   3. the bench line as the driver launches it (torchrun, one process per GPU)
mkdir -p gpurun_out
G=${1:-8}
L=gpurun_out/r2_f_mgtime${G}.log
: > $L
nvidia-smi topo -m > gpurun_out/r2_f_topo.txt 2>&1
SIZE="50000000 50000000 1473000000 16 2147483647 10"
ab() {   # name, env...
  name=$1; shift
  echo "== $name ($*)" >> $L
  env "$@" timeout 150 tools/mg_check time $G $SIZE 2>&1 | grep -v "^NCCL version" >> $L
}
ab push BLK_EXCHANGE=push
ab nccl BLK_EXCHANGE=nccl
ab pushk BLK_EXCHANGE=push BLK_PUSH_AV=kernel
ab push_p8 BLK_EXCHANGE=push BLK_PIECES=8
ab push_p2 BLK_EXCHANGE=push BLK_PIECES=2
ab push_c16 BLK_EXCHANGE=push BLK_PUSH_CTAS=16
ab push_c64 BLK_EXCHANGE=push BLK_PUSH_CTAS=64
ab grid BLK_EXPERIMENTAL=1 BLK_GRID=auto
ab colblocks4 BLK_EXPERIMENTAL=1 BLK_COLBLOCKS=4
cat $L
# parity at G GPUs
echo "== parity" >> $L
BLK_EXPERIMENTAL=1 BLK_GRID=auto timeout 120 tools/mg_check check $G 2>&1 | tail -7 >> $L
timeout 120 tools/mg_check check $G 2>&1 | tail -7 >> $L
timeout 400 python -m pytest tests/test_gpu_multi.py -q --timeout=200 -x \
  -k "(group_context and ${G}-push) or (sharded and ${G}) or cli_on_several" > gpurun_out/r2_f_pytest${G}.log 2>&1
tail -4 gpurun_out/r2_f_pytest${G}.log
tail -16 $L
# the driver's launch
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/r2_f_bench${G}.json 2> gpurun_out/r2_f_bench${G}.err
tail -c 1500 gpurun_out/r2_f_bench${G}.json
