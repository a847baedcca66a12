#!/bin/bash
# call 1 of the resumed session: validate HEAD (default attention), the 128-key kernels (classic / stale / stale+emu),
# micro-benchmarks, per-step traces, the in-step A/B on the real 14B 720P forward, one ncu capture of the stale kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv,noheader
TAIL=3 run tests_default_attn python -m pytest tests/test_kernels_gpu.py -x -q -k "attention"
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=1 TAIL=3 run tests_k128_stale python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
MV_ATTN_KSTEP=128 MV_ATTN_EMU=1 MV_ATTN_STALE=1 TAIL=3 run tests_k128_stale_emu python -m pytest tests/test_kernels_gpu.py -x -q -k "attention"
mb() { echo "--- $*"; env "$@" timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1 | cut -c1-110; }
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=1
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=1 MV_ATTN_STALE=1
mb MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=0
mb MV_ATTN_KSTEP=64 MV_ATTN_EMU=1
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=1 TAIL=20 run trace_stale python tools/attn_trace.py
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=0 TAIL=20 run trace_classic python tools/attn_trace.py
TMO=420 TAIL=8 run ab_step python tools/ab_step.py 720p
MV_ATTN_KSTEP=128 MV_ATTN_EMU=0 MV_ATTN_STALE=1 TMO=300 TAIL=2 run ncu_stale ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -f -o gpurun_out/attn_k128_stale python tools/microbench.py attn_one
