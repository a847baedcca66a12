#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
export MV_ATTN_KSTEP=128
TAIL=6 run k128_tests python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -x -q
for emu in 0 1 2; do
  for ks in 64 128; do
    echo "--- kstep $ks emu $emu"
    MV_ATTN_KSTEP=$ks MV_ATTN_EMU=$emu timeout -s KILL 200 python tools/microbench.py attn_one 2>&1 | tail -1
  done
done
MV_ATTN_KSTEP=128 TAIL=1 run bench_k128 python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
MV_ATTN_KSTEP=64 TAIL=1 run bench_k64 python bench.py --steps 2 --warmup 3 --no-vae --no-cpu-baseline
MV_ATTN_KSTEP=128 TMO=300 TAIL=2 run ncu_k128 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -f -o gpurun_out/r01_attn_v7_k128 python tools/microbench.py attn_one
