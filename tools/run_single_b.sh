#!/bin/bash
# 1-GPU call: persistent cooperative loop kernel (loop_coop.cu) -- parity suite, then configs 1-3 coop vs graph, then cfg4 unchanged
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 -k "persistent or loop_state or runtime_correctness or full_run or cli" > gpurun_out/r2_k_pytest.log 2>&1; tail -5 gpurun_out/r2_k_pytest.log
python tools/small_configs_ab.py > gpurun_out/r2_k_small_coop.json 2>&1
BLK_LOOP=graph python tools/small_configs_ab.py > gpurun_out/r2_k_small_graph.json 2>&1
cut -c1-1200 gpurun_out/r2_k_small_coop.json gpurun_out/r2_k_small_graph.json
python tools/small_phase_prof.py > gpurun_out/r2_k_small_phases.json 2>&1; cut -c1-2500 gpurun_out/r2_k_small_phases.json
