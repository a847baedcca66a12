#!/bin/bash
# 2-GPU box: Ulysses correctness (NCCL and fused P2P modes) + a short SP bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -s KILL ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAIL:-8} gpurun_out/$name.log; }
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run sp_check2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py
MOVII_SP_MODE=nccl TAIL=3 run bench2_nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --layers 4 --no-vae
MOVII_SP_MODE=p2p TAIL=3 run bench2_p2p python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --layers 4 --no-vae
TAIL=3 run bench2_full python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --no-vae
