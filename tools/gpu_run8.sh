#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
TAIL=8 run dbg_o1 python tools/attn_debug.py
MV_ATTN_ORDER=0 TAIL=8 run dbg_o0 python tools/attn_debug.py
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu
MV_ATTN_ORDER=0 run t_attn_o0 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
run t_full python -m pytest tests/test_fullsize_gpu.py -q -m gpu
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run bench_attn_o1 python tools/microbench.py attn
MV_ATTN_ORDER=0 run bench_attn_o0 python tools/microbench.py attn
TAIL=3 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -o gpurun_out/r01_attn_v4 python tools/microbench.py attn_one
TAIL=3 run bench python bench.py --steps 2 --warmup 3
TAIL=3 run ncu_vae ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_vae720.csv python tools/vae_bench.py 720p 1
