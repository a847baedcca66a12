#!/bin/bash
# call 7: one-time phase skew between the two Q tiles of the 128-key fixed-reference kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-100; }
mb MV_ATTN_SKEW=0
mb MV_ATTN_SKEW=400
mb MV_ATTN_SKEW=800
mb MV_ATTN_SKEW=1200
mb MV_ATTN_SKEW=1600
mb MV_ATTN_SKEW=2200
MV_ATTN_SKEW=800 TAIL=20 run trace_skew800 python tools/attn_trace.py
MV_ATTN_SKEW=1600 TAIL=20 run trace_skew1600 python tools/attn_trace.py
