#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
export MV_ATTN_KSTEP=128 MV_ATTN_EMU=0
TAIL=3 run tests_k128b python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_t5_gpu.py -x -q
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_PINGPONG=0
mb MV_ATTN_PINGPONG=1
mb MV_ATTN_KSTEP=64 MV_ATTN_EMU=1
TAIL=20 run trace_pp0 python tools/attn_trace.py
MV_ATTN_PINGPONG=1 TAIL=20 run trace_pp1 python tools/attn_trace.py
TAIL=1 run bench_k128 python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
MV_ATTN_PINGPONG=1 TAIL=1 run bench_k128pp python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
