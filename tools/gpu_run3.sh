#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-15} gpurun_out/$name.log; }
run t_vae python -m pytest tests/test_vae_gpu.py -q -m gpu -x
TAIL=3 run vae_small python tools/vae_bench.py small 2
TAIL=3 run vae_720 python tools/vae_bench.py 720p 2
TAIL=3 run vae_1080 python tools/vae_bench.py 1080p 2
# ncu: per-launch list of the bench command (shares of the step), then full sets for the two dominant kernels
TMO=1500 TAIL=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline
TAIL=3 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -o gpurun_out/r01_attn python tools/microbench.py attn_one
TAIL=3 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 1 -o gpurun_out/r01_gemm python tools/microbench.py gemm_one
