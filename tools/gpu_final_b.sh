#!/bin/bash
# final validation of the resumed session on one B200: everything the driver runs (pytest -m gpu, smoke, bench, reference
# arm) + the ncu launch list of the bench command + one full ncu capture of the self-attention launch at the bench shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
run pytest_gpu python -m pytest tests -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=2 run bench python bench.py --steps 2 --warmup 3
TAIL=1 TMO=300 run bench_ref python bench.py --impl reference --steps 1 --warmup 1
TAIL=6 run mb_attn python tools/microbench.py attn
MV_ATTN_EMU=1 TAIL=1 run mb_attn_emu1 python tools/microbench.py attn_one
TMO=500 TAIL=2 run ncu_attn720 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -f -o gpurun_out/r01_attn_fixref_720p python tools/microbench.py attn_720p
TMO=700 TAIL=2 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches_bench_fixref.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-vae
