#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-4} gpurun_out/$name.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run sp_check4 $TR --master-port 29521 tools/sp_check.py
TAIL=2 run bench4_720 $TR --master-port 29522 bench.py --gpus 4 --steps 2 --warmup 3 --no-vae
MOVII_SP_MODE=nccl TAIL=2 run bench4_720_nccl $TR --master-port 29523 bench.py --gpus 4 --steps 2 --warmup 3 --no-vae
TAIL=2 run bench4_1080 $TR --master-port 29524 bench.py --gpus 4 --steps 1 --warmup 3 --workload 1080p --no-vae
