#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 TAIL=4 run tests_k128 python -m pytest tests -x -q -m gpu
TAIL=3 run tests_k64_attn python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -x -q -k "attention"
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_SPLITP=1
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_SPLITP=0
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=1 MV_ATTN_SPLITP=1
mb MV_ATTN_KSTEP=64 MV_ATTN_EMU=1
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 TAIL=1 run bench_k128e0 python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
MV_ATTN_KSTEP=64 MV_ATTN_EMU=1 TAIL=1 run bench_k64e1 python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
TAIL=4 run t5_bench2 python tools/t5_bench.py
echo "--- gemm stream=1"; MV_GEMM_STREAM=1 timeout -s KILL 300 python tools/microbench.py gemm 2>&1 | tail -7
echo "--- gemm stream=0"; MV_GEMM_STREAM=0 timeout -s KILL 300 python tools/microbench.py gemm 2>&1 | tail -7
