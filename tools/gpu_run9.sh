#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
run t_vae python -m pytest tests/test_vae_gpu.py -q -m gpu
TAIL=2 run vae_720 python tools/vae_bench.py 720p 2
TAIL=2 run vae_1080 python tools/vae_bench.py 1080p 2
for sk in 0 300 600; do for em in 0 1 2; do
  echo "--- skew $sk emu $em"; MV_ATTN_SKEW=$sk MV_ATTN_EMU=$em timeout 300 python tools/microbench.py attn_one 2>&1 | tail -1
done; done
MV_ATTN_SKEW=300 run t_attn_skew python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
