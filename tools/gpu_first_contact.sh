#!/bin/bash
# First-contact run on the GPU box: each kernel family in its own process under a timeout so that a trapped
# kernel (sticky CUDA error) cannot hide the results of the others.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/$name.log; }
run t_rowops python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "ln_modulate or rmsnorm or patchify or head_unpatchify or time_embedding or errors"
run t_gemm python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm"
run t_attn python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention"
run bench_gemm python tools/microbench.py gemm
run bench_attn python tools/microbench.py attn
run bench_rowops python tools/microbench.py rowops
