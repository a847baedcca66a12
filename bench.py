#!/usr/bin/env python
"""Headline benchmark: denoising-steps/sec of the 14B T2V DiT (BASELINE.json metric).

One "step" = one iteration of the reference sampling loop (wan/text2video.py:233-254): two DiT forwards
(cond / uncond), classifier-free guidance, one UniPC scheduler update.  Synthetic latents and text embeddings,
random-init 14B weights (no checkpoints exist on the box).

    python bench.py --gpus 1 --steps 2 --warmup 3                 # 720P on one B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # Ulysses sequence parallel over N GPUs
    python bench.py --impl reference                              # the reference arithmetic on the host CPU cores

Prints ONE JSON line (contract in the task statement): value = steps/s with inputs resident in HBM; e2e = the same
through WanT2V.denoise_step with pinned-host inputs/outputs copied every step; roofline = the dominant kernel
(self-attention) timed live with CUDA events; cpu_baseline = the CPU oracle on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))

import torch  # noqa: E402

WORKLOADS = {
    # name: (width, height, frames) -> latent 16 x (f-1)/4+1 x h/8 x w/8, tokens = F * (h/16) * (w/16)
    "720p": (1280, 720, 81),
    "1080p": (1920, 832, 81),
    "480p": (832, 480, 81),
    "tiny": (256, 256, 17),
}
DIM, FFN, HEADS, LAYERS, TEXT_LEN, TEXT_DIM = 5120, 13824, 40, 40, 512, 4096


def fwd_flops(L, d=DIM, f=FFN, T=TEXT_LEN, layers=LAYERS):
    """algorithmic FLOPs of one WanModel.forward."""
    per_layer = 12 * L * d * d + 4 * T * d * d + 4 * L * L * d + 4 * L * T * d + 4 * L * d * f
    small = 2 * L * 64 * d + 2 * L * d * 64 + 2 * T * (TEXT_DIM * d + d * d) + 2 * (256 * d + d * d + 6 * d * d)
    return layers * per_layer + small


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm=d.get("hbm_gbs"),
                    src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_block_sample(L=1024, threads=None, layers_dim=(DIM, FFN, HEADS)):
    """Times the CPU oracle (oracle/dit_oracle.py, fp32) on ONE 14B-width WanAttentionBlock with L tokens and
    extrapolates steps/s at the workload by FLOPs.  Returns (seconds, flops, cores)."""
    from oracle import dit_oracle as O
    from oracle.fill import state_dict_like
    dim, ffn, nh = layers_dim
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = {}
    for a in ("self_attn", "cross_attn"):
        for n in ("q", "k", "v", "o"):
            shapes["b.%s.%s.weight" % (a, n)] = (dim, dim)
            shapes["b.%s.%s.bias" % (a, n)] = (dim,)
        shapes["b.%s.norm_q.weight" % a] = (dim,)
        shapes["b.%s.norm_k.weight" % a] = (dim,)
    shapes.update({"b.norm3.weight": (dim,), "b.norm3.bias": (dim,), "b.ffn.0.weight": (ffn, dim),
                   "b.ffn.0.bias": (ffn,), "b.ffn.2.weight": (dim, ffn), "b.ffn.2.bias": (dim,),
                   "b.modulation": (1, 6, dim)})
    sd = state_dict_like(shapes, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(L, dim, generator=g)
    e = torch.randn(6, dim, generator=g) * 0.1
    ctx = torch.randn(TEXT_LEN, dim, generator=g)
    side = int(round(L ** 0.5))
    grid = (1, side, L // side)
    ang = O.rope_table(grid, dim // nh, L)
    with torch.no_grad():
        O.block_forward(sd, "b.", x[:64], e, ang[:64], ctx, nh, 1e-6, O.ident, k_len=64)  # warm the allocator
        t0 = time.perf_counter()
        O.block_forward(sd, "b.", x, e, ang, ctx, nh, 1e-6, O.ident, k_len=L)
        dt = time.perf_counter() - t0
    flops = 12 * L * dim * dim + 4 * TEXT_LEN * dim * dim + 4 * L * L * dim + 4 * L * TEXT_LEN * dim + 4 * L * dim * ffn
    return dt, flops, cores


def cpu_steps_per_sec(seq_len, sample_L=1024):
    dt, fl, cores = cpu_block_sample(sample_L)
    rate = fl / dt
    return dict(value=rate / (2.0 * fwd_flops(seq_len)), unit="steps/s", cores=cores, kind="port",
                sample="oracle/dit_oracle.py fp32: one 14B-width WanAttentionBlock, L=%d tokens, %.2f s, %.1f GFLOP/s; "
                       "extrapolated by FLOPs to 2 forwards at L=%d" % (sample_L, dt, rate / 1e9, seq_len),
                sample_seconds=round(dt, 3), gflops=round(rate / 1e9, 1))


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU arithmetic (oracle port; the Python reference cannot travel to the box)."""
    if rank != 0:
        return
    W, H, Fr = WORKLOADS[args.workload]
    seq_len = ((Fr - 1) // 4 + 1) * (H // 16) * (W // 16)
    vals = []
    last = None
    for i in range(max(1, min(args.warmup, 1)) + max(1, args.steps)):
        last = cpu_steps_per_sec(seq_len, sample_L=args.cpu_sample_tokens)
        if i >= 1 or args.steps + args.warmup <= 1:
            vals.append(last["value"])
    vals = vals or [last["value"]]
    v = sorted(vals)[len(vals) // 2]
    last["value"] = v
    line = {"metric": "denoising_steps_per_sec", "value": v, "unit": "steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, seq_len, 1),
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, seq_len, world):
    W, H, Fr = WORKLOADS[name]
    return {"workload": "MoviiGen/Wan 14B T2V %s (%dx%d, %d frames): 2 DiT forwards + CFG + UniPC update per step"
                        % (name, W, H, Fr),
            "tokens": seq_len, "global_batch": 1, "parallelism": "ulysses-sp%d" % world if world > 1 else "single-gpu",
            "l2": "inputs larger than L2 (weights 28.6 GB + activations >7 GB stream through every step)"}


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MOVII_BENCH_WORKLOAD", "720p"), choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=LAYERS, help="debug only: a run with fewer layers is not a bench value")
    ap.add_argument("--cpu-sample-tokens", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="skip the secondary WanVAE-decode measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    import movii_b200 as mv
    import wan
    from wan.configs import Config, t2v_14B
    from wan.modules.model import WanModel
    from wan.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mv.device_check()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from xfuser.core.distributed import init_distributed_environment, initialize_model_parallel
        init_distributed_environment(rank=rank, world_size=world)
        initialize_model_parallel(sequence_parallel_degree=world, ring_degree=1, ulysses_degree=world)
    if HEADS % world != 0:
        raise SystemExit("num_heads %d is not divisible by %d ranks" % (HEADS, world))

    W, H, Fr = WORKLOADS[args.workload]
    cfg = Config(t2v_14B)
    torch.manual_seed(1234)  # identical weights on every rank (replicated)
    model = WanModel(model_type="t2v", patch_size=cfg.patch_size, text_len=cfg.text_len, in_dim=16, dim=cfg.dim,
                     ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=TEXT_DIM, out_dim=16,
                     num_heads=cfg.num_heads, num_layers=args.layers, window_size=cfg.window_size, qk_norm=True,
                     cross_attn_norm=True, eps=cfg.eps, device=dev, dtype=torch.bfloat16)
    torch.nn.init.normal_(model.head.head.weight, std=0.02)
    model.eval().requires_grad_(False)

    class _NoVae:
        class model:
            z_dim = 16
    t2v = wan.WanT2V(cfg, "", device_id=local_rank, rank=rank, use_usp=world > 1, model=model, vae=_NoVae())
    shape, seq_len = t2v.latent_geometry((W, H), Fr)
    g = torch.Generator().manual_seed(0)
    lat_host = torch.randn(*shape, generator=g).pin_memory()
    ctx_host = torch.randn(TEXT_LEN, TEXT_DIM, generator=g).to(torch.bfloat16).pin_memory()
    ctxn_host = torch.randn(TEXT_LEN, TEXT_DIM, generator=g).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(*shape).pin_memory()

    K, Wm = args.steps, args.warmup
    total_steps = Wm + 2 * K
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sched.set_timesteps(max(total_steps, 4), device=dev, shift=5.0)
    ts = list(sched.timesteps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, first_idx):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            fn(first_idx + i)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item()

    state = {"lat": lat_host.to(dev), "ctx": [ctx_host.to(dev)], "ctxn": [ctxn_host.to(dev)]}

    def step_resident(i):
        state["lat"] = t2v.denoise_step(sched, state["lat"], ts[i], state["ctx"], state["ctxn"], seq_len, 5.0)

    def step_e2e(i):
        lat = lat_host.to(dev, non_blocking=True)
        c = [ctx_host.to(dev, non_blocking=True)]
        cn = [ctxn_host.to(dev, non_blocking=True)]
        new = t2v.denoise_step(sched, lat, ts[i], c, cn, seq_len, 5.0)
        out_host.copy_(new, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the result

    for i in range(Wm):
        step_resident(i)
    cvd = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
    smi_index = cvd[local_rank] if local_rank < len(cvd) else local_rank
    sampler = ClockSampler(smi_index) if rank == 0 else None
    l0 = mv.LAUNCHES
    rec = mv.time_kernels(["mv_attention_fwd", "mv_attention_fwd_scatter"])
    ms_res, wall_res = timed(step_resident, K, Wm)
    launches = mv.LAUNCHES - l0
    # normalise both entry points to (Lq, Lk, H): plain = args[8:11]; scatter (fused Ulysses return) = args[11:14]
    attn_events = [(s_, e_, (a[8], a[9], a[10])) for (s_, e_, a) in rec.get("mv_attention_fwd", [])] + \
                  [(s_, e_, (a[11], a[12], a[13])) for (s_, e_, a) in rec.get("mv_attention_fwd_scatter", [])]
    mv.time_kernels(None)
    clocks = sampler.stop() if sampler else None
    ms_e2e, wall_e2e = timed(step_e2e, K, Wm + K)
    finite = bool(torch.isfinite(state["lat"]).all().item())

    # ---- dominant kernel: self-attention launches (Lk == sequence length), algorithmic FLOPs / event time
    self_attn = [(s.elapsed_time(e), a) for (s, e, a) in attn_events if a[1] > TEXT_LEN]   # a = (Lq, Lk, H)
    roof = None
    pk = peaks()
    if self_attn:
        avg_ms = sum(t for t, _ in self_attn) / len(self_attn)
        a = self_attn[0][1]
        Lq, Lk, Hh = a
        fl = 4.0 * Lq * Lk * Hh * 128
        ach = fl / (avg_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if os.path.isfile(tp):
            try:
                with open(tp) as fh:
                    traffic = json.load(fh).get("%s_sp%d" % (args.workload, world))
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": "mv_attention_fwd (self-attention, Lq=%d Lk=%d H=%d)" % (Lq, Lk, Hh),
                "achieved": round(ach, 1), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops"], 4),
                "traffic": traffic, "peak_source": pk["src"], "launches_timed": len(self_attn),
                "avg_launch_ms": round(avg_ms, 4),
                "share_of_step": round(sum(t for t, _ in self_attn) / max(ms_res, 1e-9), 4)}

    # ---- WanVAE decode of this workload's latent (rank 0 only, as in the reference: text2video.py:260-261)
    vae_rec = None
    if rank == 0 and not args.no_vae:
        try:
            from wan.modules.vae import WanVAE
            torch.manual_seed(3)
            vae = WanVAE(vae_pth=None, device=dev)
            zlat = torch.randn(*shape, device=dev)
            vae.decode([zlat])                      # warm-up (packs weights, sizes the allocator)
            torch.cuda.synchronize()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0v = mv.LAUNCHES
            v0.record()
            vid = vae.decode([zlat])[0]
            v1.record()
            torch.cuda.synchronize()
            vms = v0.elapsed_time(v1)
            vae_rec = {"metric": "vae_decode_fps", "value": round(vid.shape[1] / vms * 1e3, 2), "unit": "frames/s",
                       "ms": round(vms, 1), "frames": int(vid.shape[1]), "out": list(vid.shape),
                       "gpu_launches": mv.LAUNCHES - l0v, "finite": bool(torch.isfinite(vid).all().item())}
            del vae, vid, zlat
            torch.cuda.empty_cache()
        except Exception as ex:  # the DiT number must still be reported
            vae_rec = {"error": repr(ex)[:200]}
    if world > 1:
        dist.barrier()

    if rank == 0:
        sps = K / (ms_res * 1e-3)
        sps_e2e = K / (ms_e2e * 1e-3)
        step_fl = 2.0 * fwd_flops(seq_len, layers=args.layers)
        line = {"metric": "denoising_steps_per_sec", "value": sps, "unit": "steps/s", "n_gpus": world, "steps": K,
                "warmup": Wm, "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args.workload, seq_len, world),
                "model_tflops_per_gpu": round(step_fl * sps / world / 1e12, 1),
                "model_frac_of_peak": round(step_fl * sps / world / 1e12 / pk["tflops"], 4),
                "e2e": {"value": sps_e2e, "unit": "steps/s",
                        "h2d_bytes_per_step": lat_host.numel() * 4 + 2 * ctx_host.numel() * 2,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / K},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "finite": finite,
                "wall_ms_per_step": wall_res / K}
        if args.layers != LAYERS:
            line["INVALID"] = "debug run with %d of %d layers" % (args.layers, LAYERS)
        if vae_rec is not None:
            line["vae_decode"] = vae_rec
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_steps_per_sec(seq_len, args.cpu_sample_tokens)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
