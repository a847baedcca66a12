#!/bin/bash
# call 2: new default (128-key classic, o_done probed first) vs the shadow-max stale variant
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
TAIL=3 run tests_default python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
MV_ATTN_STALE=1 TAIL=3 run tests_stale2 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -x -q -k "attention or fullsize or full"
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_STALE=0
mb MV_ATTN_STALE=1
MV_ATTN_STALE=1 TAIL=20 run trace_stale2 python tools/attn_trace.py
MV_ATTN_STALE=0 TAIL=20 run trace_classic2 python tools/attn_trace.py
TMO=420 TAIL=8 run ab_step2 python tools/ab_step.py 720p 128:0:0:0 128:0:1:0 128:0:0:0 128:0:1:0
