#!/bin/bash
# 2-GPU box: Ulysses correctness with the fixed-reference attention kernel (NCCL + fused P2P) and the N=2 bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-100} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-8} gpurun_out/$name.log; }
TMO=90 run sp_check2b python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py
TMO=120 TAIL=2 run bench2_fixref python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 3 --no-vae
