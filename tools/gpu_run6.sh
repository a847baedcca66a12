#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
TAIL=8 run dbg_o0 python tools/attn_debug.py
MV_ATTN_ORDER=1 TAIL=8 run dbg_o1 python tools/attn_debug.py
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu
MV_ATTN_ORDER=1 run t_attn_o1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
MV_ATTN_ORDER=1 MV_ATTN_EMU=1 run t_attn_o1e1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
MV_ATTN_EMU=2 run t_attn_emu2 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run bench_attn_o0e0 python tools/microbench.py attn
MV_ATTN_ORDER=1 run bench_attn_o1e0 python tools/microbench.py attn
MV_ATTN_EMU=1 run bench_attn_o0e1 python tools/microbench.py attn
MV_ATTN_ORDER=1 MV_ATTN_EMU=1 run bench_attn_o1e1 python tools/microbench.py attn
MV_ATTN_ORDER=1 TAIL=3 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -o gpurun_out/r01_attn_v3 python tools/microbench.py attn_one
TAIL=3 run bench python bench.py --steps 2 --warmup 3
