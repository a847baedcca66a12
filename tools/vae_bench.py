"""WanVAE decode benchmark (BASELINE.json configs[3]): frames/s of WanVAE.decode on a synthetic latent, one B200.
Prints one JSON line.  VAE FLOPs/bytes per SURVEY.md §8d (1080P: 1116.5 TF, 720P: 639.2 TF).
`python tools/vae_bench.py enc:1080p` times WanVAE.encode on the matching 81-frame video instead (§8f-4)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402
from wan.modules.vae import WanVAE  # noqa: E402

SHAPES = {"1080p": (21, 104, 240, 1116.5), "720p": (21, 90, 160, 639.2), "480p": (21, 60, 104, None),
          "small": (5, 32, 32, None)}


def encode_tflop(T, H, W, dim=96):
    """Algorithmic FLOPs of Encoder3d (vae.py:265-366) on a [3, T, H, W] video, in TFLOP (2 * MACs; attention incl.)."""
    fl, t, h, w = 0.0, T, H, W
    conv = lambda ci, co, taps, vox: 2.0 * ci * co * taps * vox  # noqa: E731
    fl += conv(3, dim, 27, t * h * w)
    dims = [dim, dim, 2 * dim, 4 * dim, 4 * dim]
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(2):
            fl += conv(ci, co, 27, t * h * w) + conv(co, co, 27, t * h * w) + (conv(ci, co, 1, t * h * w) if ci != co else 0)
            ci = co
        if i < 3:
            h, w = h // 2, w // 2
            fl += conv(co, co, 9, t * h * w)
            if i > 0:
                t = 1 + (t - 1) // 2
                fl += conv(co, co, 3, (t - 1) * h * w)
    c = dims[-1]
    fl += 4 * conv(c, c, 27, t * h * w) + conv(c, 3 * c, 1, t * h * w) + conv(c, c, 1, t * h * w) + t * 4.0 * (h * w) ** 2 * c
    fl += conv(c, 32, 27, t * h * w)
    return fl / 1e12


def main_encode(which, iters):
    T, h, w, _ = SHAPES[which]
    F, H, W = 1 + 4 * (T - 1), 8 * h, 8 * w
    mv.device_check()
    torch.manual_seed(3)
    vae = WanVAE(vae_pth=None, device="cuda")
    x = torch.rand(3, F, H, W, device="cuda") * 2 - 1
    mu = vae.encode([x])[0]
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        l0 = mv.LAUNCHES
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        mu = vae.encode([x])[0]
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
        launches = mv.LAUNCHES - l0
    ms = sorted(ts)[len(ts) // 2]
    tf = encode_tflop(F, H, W)
    print(json.dumps(dict(metric="vae_encode_fps", workload=which, video=[3, F, H, W], out=list(mu.shape), ms=round(ms, 2),
                          value=round(F / ms * 1e3, 2), unit="frames/s", gpu_launches=launches, algorithmic_tflop=round(tf, 1),
                          tflops=round(tf / ms * 1e3, 1), finite=bool(torch.isfinite(mu).all()),
                          peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "1080p"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    if which.startswith("enc:"):
        return main_encode(which[4:], iters)
    T, h, w, tf = SHAPES[which]
    mv.device_check()
    torch.manual_seed(3)
    vae = WanVAE(vae_pth=None, device="cuda")
    z = torch.randn(16, T, h, w, device="cuda")
    out = vae.decode([z])[0]  # warm-up (also builds the engine)
    out = vae.decode([z])[0]  # second warm-up: with MOVII_VAE_GRAPH=1 this shape is captured in a CUDA graph here
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        l0 = mv.LAUNCHES
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = vae.decode([z])[0]
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
        launches = mv.LAUNCHES - l0
    ms = sorted(ts)[len(ts) // 2]
    frames = out.shape[1]
    rec = dict(metric="vae_decode_fps", workload=which, latent=[16, T, h, w], out=list(out.shape), ms=round(ms, 2),
               value=round(frames / ms * 1e3, 2), unit="frames/s", gpu_launches=launches,
               finite=bool(torch.isfinite(out).all()), peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))
    if tf:
        rec["tflops"] = round(tf / ms * 1e3, 1)
    print(json.dumps(rec), flush=True)
    # where the decode goes: one more decode with every launch bracketed by CUDA events (adds launch gaps; shares only)
    names = ["mv_vae_conv", "mv_vae_conv_fused", "mv_vae_rmsnorm_silu", "mv_gemm_f16", "mv_softmax_rows", "mv_vae_latent_in"]
    timed = mv.time_kernels(names)
    vae.decode([z])
    torch.cuda.synchronize()
    agg = {}
    for n in names:
        for (s_, e_, a) in timed.get(n, []):
            if n == "mv_vae_conv":
                key = "conv %dtap %d->%d @H%d" % (a[15], a[4], a[13], a[2])
            elif n == "mv_vae_conv_fused":
                key = "conv+norm %dtap %d->%d @H%d" % (a[13], a[4], a[12], a[2])
            elif n == "mv_vae_rmsnorm_silu":
                key = "rmsnorm_silu C%d" % a[4]
            else:
                key = n
            t_, c_ = agg.get(key, (0.0, 0))
            agg[key] = (t_ + s_.elapsed_time(e_), c_ + 1)
    mv.time_kernels(None)
    tot = sum(v[0] for v in agg.values())
    print(json.dumps(dict(breakdown_ms={k: [round(v[0], 1), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
                          sum_ms=round(tot, 1))), flush=True)


if __name__ == "__main__":
    main()
