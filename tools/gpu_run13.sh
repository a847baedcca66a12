#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
TAIL=15 run t5_tests python -m pytest tests/test_t5_gpu.py -x -q
TAIL=6 run t5_bench python tools/t5_bench.py
TAIL=3 run gpu_tests python -m pytest tests -x -q -m gpu --deselect tests/test_t5_gpu.py
TAIL=7 run sweep python tools/sweep.py
TMO=500 TAIL=2 run ncu_attn720 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -c 1 -f -o gpurun_out/r01_attn_final_720p python tools/microbench.py attn_720p
TMO=300 TAIL=2 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm -c 1 -f -o gpurun_out/r01_gemm_final python tools/microbench.py gemm_one
TMO=300 TAIL=2 run ncu_t5 ncu --set full --clock-control none --import-source on -k regex:t5_attention -c 1 -f -o gpurun_out/r01_t5_attention python -m pytest tests/test_t5_gpu.py -x -q -k "umt5_width"
TMO=700 TAIL=2 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-vae
