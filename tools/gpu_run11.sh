#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
MV_ATTN_PINGPONG=1 TAIL=8 run dbg_pp python tools/attn_debug.py
MV_ATTN_PINGPONG=1 run t_attn_pp python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention
MV_ATTN_PINGPONG=1 run t_full_pp python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k attention
for pp in 0 1; do for em in 0 1 2; do
  echo "--- pingpong $pp emu $em"; MV_ATTN_PINGPONG=$pp MV_ATTN_EMU=$em timeout 300 python tools/microbench.py attn_one 2>&1 | tail -1
done; done
MV_ATTN_PINGPONG=1 TAIL=6 run bench_attn_pp python tools/microbench.py attn
MV_ATTN_PINGPONG=1 TAIL=3 run ncu_attn_pp ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -o gpurun_out/r01_attn_v6_pp python tools/microbench.py attn_one
MV_ATTN_PINGPONG=1 TAIL=3 run bench_pp python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vae
