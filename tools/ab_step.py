"""In-step A/B of the attention kernel variants on the real 14B 720P forward (one process, one model build).
For every variant: 1 untimed forward, then 2 timed forwards (= the DiT part of one denoising step) with per-launch
CUDA-event timing of the self-attention kernel; SM clock / power sampled with nvidia-smi while timing.
Usage: python tools/ab_step.py [workload] [variant ...]   variant = kstep:emu:stale:pingpong[:skew]  (default list below)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import bench  # noqa: E402
import movii_b200 as mv  # noqa: E402
from wan.configs import Config, t2v_14B  # noqa: E402
from wan.modules.model import WanModel  # noqa: E402


def main():
    args = sys.argv[1:]
    workload = args[0] if args and args[0] in bench.WORKLOADS else "720p"
    # variant = kstep:emu:stale:pingpong[:skew][,pair=0|1][,hi=0|1]   (pair: CTA-pair GEMM kernel; hi: role warps at the
    # highest warp ids in every tcgen05 kernel)
    variants = [a for a in args if ":" in a] or ["128:0:1:0,pair=0", "128:0:1:0,pair=1", "128:0:1:0,pair=0"]
    dev = torch.device("cuda", 0)
    mv.device_check()
    cfg = Config(t2v_14B)
    torch.manual_seed(1234)
    model = WanModel(model_type="t2v", patch_size=cfg.patch_size, text_len=cfg.text_len, in_dim=16, dim=cfg.dim,
                     ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=4096, out_dim=16, num_heads=cfg.num_heads,
                     num_layers=cfg.num_layers, window_size=cfg.window_size, qk_norm=True, cross_attn_norm=True,
                     eps=cfg.eps, device=dev, dtype=torch.bfloat16)
    torch.nn.init.normal_(model.head.head.weight, std=0.02)
    model.eval().requires_grad_(False)
    W, H, Fr = bench.WORKLOADS[workload]
    shape = (16, (Fr - 1) // 4 + 1, H // 8, W // 8)
    seq_len = shape[1] * (shape[2] // 2) * (shape[3] // 2)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(*shape, generator=g).to(dev)
    ctx = [torch.randn(512, 4096, generator=g).to(torch.bfloat16).to(dev)]
    t = torch.tensor([900], device=dev)
    ref = None
    for var in variants:
        attn_var, _, opts = var.partition(",")
        ks, emu, stale, pp, skew = (int(x) for x in (attn_var.split(":") + ["0"])[:5])
        mv.attention_config(ks, emu, stale, pp, skew)
        for o in [x for x in opts.split(",") if x]:
            k, _, v = o.partition("=")
            if k == "pair":
                mv.gemm_config(int(v))
            if k == "hi":
                mv.roles_config(int(v))
            if k == "spin":
                mv.attention_config(wait_spin=int(v))
            if k == "pack":
                mv.attention_config(pack=int(v))
        out = model([lat], t=t, context=ctx, seq_len=seq_len)[0]
        torch.cuda.synchronize()
        sampler = bench.ClockSampler(0)
        rec = mv.time_kernels(["mv_attention_fwd"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            out = model([lat], t=t, context=ctx, seq_len=seq_len)[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sa = [s.elapsed_time(e) for (s, e, a) in rec["mv_attention_fwd"] if a[9] > 512]
        mv.time_kernels(None)
        clocks = sampler.stop()
        fl = 4.0 * seq_len * seq_len * 40 * 128
        if ref is None:
            ref = out.float().clone()
        rel = ((out.float() - ref).norm() / ref.norm()).item()
        print(json.dumps(dict(variant=var, two_forwards_ms=round(ms, 1), attn_avg_ms=round(sum(sa) / len(sa), 3),
                              attn_tflops=round(fl / (sum(sa) / len(sa)) / 1e9, 1), attn_share=round(sum(sa) / ms, 4),
                              rest_ms=round(ms - sum(sa), 1), rel_vs_first=round(rel, 6),
                              finite=bool(torch.isfinite(out).all().item()), clocks=clocks)), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("ab_step wall %.0f s" % (time.time() - t0))
