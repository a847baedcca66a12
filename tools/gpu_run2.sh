#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-15} gpurun_out/$name.log; }
run smoke python __graft_entry__.py --smoke
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu -x
TAIL=5 run bench_dbg python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline
TAIL=5 run bench python bench.py --steps 2 --warmup 3
