"""one full 14B DiT forward (40 layers, random-init) at L = 4k .. 128k tokens, SP = 1 (plain
python) or SP = WORLD_SIZE under torchrun (Ulysses over all ranks).  Grids (F, H', W') = (1|2|4|8|16|32, 64, 64).
One JSON line per L (rank 0): ms (max over ranks), model TFLOP/s per GPU (algorithmic FLOPs of §8d)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402
from bench import fwd_flops  # noqa: E402
from wan.configs import Config, t2v_14B  # noqa: E402
from wan.modules.model import WanModel  # noqa: E402


def main():
    frames = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32]
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mv.device_check()
    if world > 1:
        import types
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from wan.distributed.xdit_context_parallel import usp_dit_forward
        from xfuser.core.distributed import init_distributed_environment, initialize_model_parallel
        init_distributed_environment(rank=rank, world_size=world)
        initialize_model_parallel(sequence_parallel_degree=world, ring_degree=1, ulysses_degree=world)
    cfg = Config(t2v_14B)
    torch.manual_seed(1234)
    model = WanModel(model_type="t2v", patch_size=cfg.patch_size, text_len=cfg.text_len, in_dim=16, dim=cfg.dim,
                     ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=4096, out_dim=16, num_heads=cfg.num_heads,
                     num_layers=cfg.num_layers, window_size=cfg.window_size, qk_norm=True, cross_attn_norm=True,
                     eps=cfg.eps, device=dev, dtype=torch.bfloat16)
    torch.nn.init.normal_(model.head.head.weight, std=0.02)
    model.eval().requires_grad_(False)
    if world > 1:
        model.forward = types.MethodType(usp_dit_forward, model)
    g = torch.Generator().manual_seed(0)
    ctx = [torch.randn(512, 4096, generator=g).to(torch.bfloat16).to(dev)]
    t = torch.tensor([500], device=dev)
    for F in frames:
        L = F * 64 * 64
        lat = torch.randn(16, F, 128, 128, generator=g).to(dev)
        ts = []
        for i in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = model([lat], t=t, context=ctx, seq_len=L)[0]
            e.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if i:
                ts.append(ms.item())
        ms = min(ts)
        fl = fwd_flops(L)
        if rank == 0:
            print(json.dumps(dict(kind="dit_forward", L=L, sp=world, grid=[F, 64, 64], ms=round(ms, 2),
                                  tflops_per_gpu=round(fl / ms / 1e9 / world, 1), pflop=round(fl / 1e15, 4),
                                  finite=bool(torch.isfinite(out).all()))), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
