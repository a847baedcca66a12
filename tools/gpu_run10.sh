#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
TAIL=8 run dbg python tools/attn_debug.py
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu
MV_ATTN_EMU=0 run t_attn_e0 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
run t_full python -m pytest tests/test_fullsize_gpu.py -q -m gpu
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run t_vae python -m pytest tests/test_vae_gpu.py -q -m gpu
run bench_attn python tools/microbench.py attn
MV_ATTN_EMU=0 TAIL=2 run bench_attn_e0 python tools/microbench.py attn_one
MV_ATTN_EMU=2 TAIL=2 run bench_attn_e2 python tools/microbench.py attn_one
TAIL=2 run vae_720 python tools/vae_bench.py 720p 2
TAIL=2 run vae_1080 python tools/vae_bench.py 1080p 2
TAIL=3 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -o gpurun_out/r01_attn_v5 python tools/microbench.py attn_one
TAIL=3 run bench python bench.py --steps 2 --warmup 3
