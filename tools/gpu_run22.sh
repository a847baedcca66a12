#!/bin/bash
# call 6: paired-warpgroup attention kernel (MV_ATTN_SPLIT=2) vs the fixed-reference default
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
MV_ATTN_SPLIT=2 TAIL=6 run tests_pair python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_SPLIT=0
mb MV_ATTN_SPLIT=2
mb MV_ATTN_SPLIT=2 MV_ATTN_EMU=1
MV_ATTN_SPLIT=2 TAIL=20 run trace_pair python tools/attn_trace.py
MV_ATTN_SPLIT=0 TAIL=20 run trace_fix python tools/attn_trace.py
TMO=420 TAIL=8 run ab_step5 python tools/ab_step.py 720p 128:0:1:0:0 128:0:1:0:2 128:1:1:0:2 128:0:1:0:0 128:0:1:0:2
