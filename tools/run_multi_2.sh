mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_d_topo.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py -q --timeout=300 -x > gpurun_out/r2_d_pytest_multi.log 2>&1
tail -15 gpurun_out/r2_d_pytest_multi.log
for mode in nccl push pushk ce; do
  case $mode in
    nccl) env="BLK_EXCHANGE=nccl";;
    push) env="BLK_EXCHANGE=push";;
    pushk) env="BLK_EXCHANGE=push BLK_PUSH_AV=kernel";;
    ce) env="BLK_EXCHANGE=ce";;
  esac
  echo "== $mode" >> gpurun_out/r2_d_mgtime2.log
  env $env timeout 200 tools/mg_check time 2 20000000 20000000 600000000 16 2147483647 10 >> gpurun_out/r2_d_mgtime2.log 2>&1
done
timeout 200 tools/mg_check check 2 >> gpurun_out/r2_d_mgtime2.log 2>&1
cat gpurun_out/r2_d_mgtime2.log
