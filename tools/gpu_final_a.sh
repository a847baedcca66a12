#!/bin/bash
# final validation on one B200: everything the driver runs (pytest -m gpu, smoke, bench) + the 1080P data points
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-1200} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
run pytest_gpu python -m pytest tests -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=2 run bench_attn python tools/microbench.py attn
TAIL=3 run bench python bench.py --steps 2 --warmup 3
TAIL=3 run bench_1080 python bench.py --steps 1 --warmup 3 --workload 1080p --no-cpu-baseline
