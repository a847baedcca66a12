#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
TAIL=3 run tests_default_attn python -m pytest tests/test_kernels_gpu.py -x -q -k "attention"
export MV_ATTN_KSTEP=128 MV_ATTN_EMU=0
MV_ATTN_STALE=0 TAIL=3 run tests_k128_classic python -m pytest tests/test_kernels_gpu.py -x -q -k "attention"
MV_ATTN_STALE=1 TAIL=3 run tests_k128_stale python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_STALE=0
mb MV_ATTN_STALE=1
MV_ATTN_STALE=1 TAIL=20 run trace_stale python tools/attn_trace.py
MV_ATTN_STALE=1 TAIL=1 run bench_k128stale python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
MV_ATTN_STALE=0 TAIL=1 run bench_k128classic python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
