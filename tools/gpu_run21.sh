#!/bin/bash
# call 5: fixed-reference softmax (no per-step row max) in the 128-key kernel and in the row-split kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
MV_ATTN_SPLIT=1 MV_ATTN_STALE=1 TAIL=6 run tests_split_fix python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
MV_ATTN_SPLIT=0 MV_ATTN_STALE=1 TAIL=3 run tests_fix python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -x -q -k "attention or full"
MV_ATTN_SPLIT=1 TAIL=3 run tests_split_scatter python -m pytest tests/test_kernels_gpu.py -x -q -k "attention"
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_SPLIT=0 MV_ATTN_STALE=0
mb MV_ATTN_SPLIT=0 MV_ATTN_STALE=1
mb MV_ATTN_SPLIT=1 MV_ATTN_STALE=0
mb MV_ATTN_SPLIT=1 MV_ATTN_STALE=1
mb MV_ATTN_SPLIT=1 MV_ATTN_STALE=1 MV_ATTN_EMU=1
MV_ATTN_SPLIT=1 MV_ATTN_STALE=1 TAIL=20 run trace_split_fix python tools/attn_trace.py
TMO=420 TAIL=8 run ab_step4 python tools/ab_step.py 720p 128:0:0:0:0 128:0:1:0:0 128:0:1:0:1 128:0:0:0:0 128:0:1:0:0 128:0:1:0:1
