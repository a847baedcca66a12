"""Mints tests/golden/*.pt from the REAL reference code (/root/reference, imported through oracle/ref_loader.py).

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The reference ships no tests or golden vectors; these files pin the oracle (and, on the GPU, the
CUDA path) to the reference's own forward code on fixed seeds.  fp32, CPU, SDPA in place of flash-attn.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.fill import fill_parameters, fill_t5  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_block_cfg1(model):
    """BASELINE.json configs[0]: single WanAttentionBlock, dim 128, 1x16x16 token grid (256 tokens)."""
    torch.manual_seed(0)
    blk = model.WanAttentionBlock("t2v_cross_attn", 128, 512, 4, (-1, -1), True, True, 1e-6).eval()
    fill_parameters(blk, 101)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 256, 128, generator=g)
    e = torch.randn(1, 6, 128, generator=g)
    ctx = torch.randn(1, 512, 128, generator=g)
    dh = 32
    freqs = torch.cat([model.rope_params(1024, dh - 4 * (dh // 6)), model.rope_params(1024, 2 * (dh // 6)),
                       model.rope_params(1024, 2 * (dh // 6))], 1)
    with torch.no_grad():
        y = blk(x, e, torch.tensor([256]), torch.tensor([[1, 16, 16]]), freqs, ctx, None)
    return dict(x=x, e=e, ctx=ctx, y=y, seed=101, grid=(1, 16, 16),
                param_shapes={k: tuple(v.shape) for k, v in blk.state_dict().items()})


TINY = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=16, in_dim=16, dim=256, ffn_dim=512, freq_dim=64,
            text_dim=64, out_dim=16, num_heads=2, num_layers=2, qk_norm=True, cross_attn_norm=True, eps=1e-6)


def golden_model_tiny(model):
    """Tiny WanModel with head_dim 128 (what the sm_100a attention kernel supports): 2 layers, 32 real tokens padded to 40."""
    m = model.WanModel(**TINY).eval()
    fill_parameters(m, 202)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(16, 2, 8, 8, generator=g)
    ctx = torch.randn(11, 64, generator=g)
    t = torch.tensor([642])
    outs = {}
    with torch.no_grad():
        for seq_len in (32, 40):
            outs[seq_len] = m([x], t, [ctx], seq_len)[0]
    return dict(cfg=TINY, x=x, ctx=ctx, t=t, y32=outs[32], y40=outs[40], seed=202,
                param_shapes={k: tuple(v.shape) for k, v in m.state_dict().items()})


def golden_unipc():
    """Trajectory of the reference FlowUniPCMultistepScheduler on a fixed pseudo-model (needs a fuller diffusers stub)."""
    du = types.ModuleType("diffusers.utils")
    du.deprecate = lambda *a, **k: None
    du.is_scipy_available = lambda: False
    sch = types.ModuleType("diffusers.schedulers")
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")

    class SchedulerMixin:
        pass

    class SchedulerOutput(dict):
        pass

    class _K:
        name = "x"

    su.KarrasDiffusionSchedulers = [_K]
    su.SchedulerMixin, su.SchedulerOutput = SchedulerMixin, SchedulerOutput

    class ConfigMixin:
        pass

    def register_to_config(init):
        def wrapped(self, *a, **k):
            import inspect
            sig = inspect.signature(init)
            ba = sig.bind(self, *a, **k)
            ba.apply_defaults()
            cfg = types.SimpleNamespace(**{n: v for n, v in ba.arguments.items() if n != "self"})
            self.config = cfg
            self.register_to_config = lambda **kw: [setattr(cfg, a_, b_) for a_, b_ in kw.items()]
            init(self, *a, **k)
        return wrapped

    cu = sys.modules["diffusers.configuration_utils"]
    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    sys.modules["diffusers.utils"] = du
    sys.modules["diffusers.schedulers"] = sch
    sys.modules["diffusers.schedulers.scheduling_utils"] = su
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location(
        "refwan_unipc", os.path.join(ref_loader.REF_ROOT, "wan", "utils", "fm_solvers_unipc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = {}
    for steps in (4, 9):
        s = mod.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(steps, device="cpu", shift=5.0)
        g = torch.Generator().manual_seed(steps)
        x = torch.randn(1, 16, 2, 4, 4, generator=g)
        traj, outs = [x.clone()], []
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints debug lines every step
            for t in s.timesteps:
                v = torch.sin(3.0 * x) * 0.5 + 0.1 * torch.randn(x.shape, generator=g)  # pseudo model output
                outs.append(v)
                x = s.step(v, t, x, return_dict=False)[0]
                traj.append(x.clone())
        res[steps] = dict(timesteps=s.timesteps.clone(), sigmas=s.sigmas.clone(), model_outputs=outs, traj=traj)
    return res


def golden_vae(vae):
    """Reference chunked decode (WanVAE_.decode with its feature cache) on small latents: T = 1, 2, 3, 5 covers the
    first-chunk 'Rep' path, the 1-frame cache concat and the steady state; h x w = 5 x 9 exercises ragged 8x16 tiles."""
    from oracle.vae_oracle import VAE_MEAN, VAE_STD
    m = vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True]).eval()
    # only the decoder half is filled (and stored): the tests rebuild exactly these parameters by name
    fill_parameters([(n, p) for n, p in m.named_parameters() if n.startswith(("decoder.", "conv2."))], 404)
    mean, std = torch.tensor(VAE_MEAN), torch.tensor(VAE_STD)
    res = {"seed": 404, "cases": {}}
    res["param_shapes"] = {k: tuple(v.shape) for k, v in m.state_dict().items()
                           if k.startswith("decoder.") or k.startswith("conv2.")}
    for T, h, w in ((1, 4, 6), (2, 5, 9), (3, 4, 6), (5, 4, 4)):
        g = torch.Generator().manual_seed(100 + T)
        z = torch.randn(16, T, h, w, generator=g)
        with torch.no_grad():
            y = m.decode(z[None], [mean, 1.0 / std]).float().clamp_(-1, 1)[0]
        res["cases"][(T, h, w)] = dict(z=z, y=y.to(torch.float16))  # fp16 storage: 5e-4 abs on [-1,1], keeps files small
    return res


def golden_vae_encode(vae):
    """Reference chunked encode (WanVAE_.encode: 1 + 4 + 4 ... frames with its feature cache) on small videos: T = 1 (no
    time_conv at all), 5 (one cached frame), 9 / 13 (steady state); H x W multiples of 8 incl. ragged 8x16 tiles."""
    from oracle.vae_oracle import VAE_MEAN, VAE_STD
    m = vae.WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
                    temperal_downsample=[False, True, True]).eval()
    fill_parameters([(n, p) for n, p in m.named_parameters() if n.startswith(("encoder.", "conv1."))], 505)
    mean, std = torch.tensor(VAE_MEAN), torch.tensor(VAE_STD)
    res = {"seed": 505, "cases": {}}
    res["param_shapes"] = {k: tuple(v.shape) for k, v in m.state_dict().items()
                           if k.startswith("encoder.") or k.startswith("conv1.")}
    for T, H, W in ((1, 16, 24), (5, 24, 40), (9, 16, 16), (13, 32, 16)):
        g = torch.Generator().manual_seed(200 + T)
        x = (torch.rand(3, T, H, W, generator=g) * 2 - 1).to(torch.float16).float()     # fp16-exact pixels in [-1, 1]
        with torch.no_grad():
            mu = m.encode(x[None], [mean, 1.0 / std]).float()[0]
        res["cases"][(T, H, W)] = dict(x=x.to(torch.float16), mu=mu)
    return res


def golden_dpmpp():
    """Trajectory of the reference FlowDPMSolverMultistepScheduler driven like text2video.py:214-223 (needs the
    diffusers stubs installed by golden_unipc)."""
    import contextlib
    import importlib.util
    import io
    du = sys.modules["diffusers.utils"]
    tu = types.ModuleType("diffusers.utils.torch_utils")
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(shape, generator=generator, dtype=dtype)
    sys.modules["diffusers.utils.torch_utils"] = tu
    du.torch_utils = tu
    spec = importlib.util.spec_from_file_location(
        "refwan_dpm", os.path.join(ref_loader.REF_ROOT, "wan", "utils", "fm_solvers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = {}
    for steps in (3, 8, 20):
        s = mod.FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sig = mod.get_sampling_sigmas(steps, 5.0)
        timesteps, _ = mod.retrieve_timesteps(s, device="cpu", sigmas=sig)
        g = torch.Generator().manual_seed(50 + steps)
        x = torch.randn(1, 16, 2, 4, 4, generator=g)
        traj, outs = [x.clone()], []
        with contextlib.redirect_stdout(io.StringIO()):
            for t in timesteps:
                v = torch.cos(2.0 * x) * 0.5 + 0.1 * torch.randn(x.shape, generator=g)
                outs.append(v)
                x = s.step(v, t, x, return_dict=False)[0]
                traj.append(x.clone())
        res[steps] = dict(timesteps=timesteps.clone(), sigmas=s.sigmas.clone(), model_outputs=outs, traj=traj)
    return res


T5_TINY = dict(vocab=97, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2, num_buckets=32,
               shared_pos=False, dropout=0.1)


def golden_t5():
    """Reference T5Encoder (t5.py:267-312), tiny umT5-style config with head_dim 64 (what the sm_100a kernel
    supports), fp32 / eval.  Cases: (L, valid) = (24, 17) one key box; (300, 263) two boxes, three query tiles and
    relative distances past max_dist=128 (every bucket); (40, 40) no padding."""
    t5 = ref_loader.load_reference_t5()
    m = t5.T5Encoder(**T5_TINY).eval()
    fill_t5(m, 505)
    res = {"cfg": T5_TINY, "seed": 505, "cases": {},
           "param_shapes": {k: tuple(v.shape) for k, v in m.state_dict().items()}}
    for L, valid in ((24, 17), (300, 263), (40, 40)):
        g = torch.Generator().manual_seed(600 + L)
        ids = torch.randint(1, T5_TINY["vocab"], (L,), generator=g)
        ids[valid:] = 0
        mask = (torch.arange(L) < valid).long()
        with torch.no_grad():
            y = m(ids[None], mask[None])[0]
        res["cases"][(L, valid)] = dict(ids=ids, mask=mask, y=y)
    return res


def main():
    """python oracle/make_golden.py [name ...]: mint all goldens, or only the named ones (e.g. vae_encode)."""
    os.makedirs(OUT, exist_ok=True)
    att, model, vae = ref_loader.load_reference()
    jobs = {"vae_decode": lambda: golden_vae(vae), "vae_encode": lambda: golden_vae_encode(vae),
            "block_cfg1": lambda: golden_block_cfg1(model), "model_tiny_hd128": lambda: golden_model_tiny(model),
            "unipc": golden_unipc, "dpmpp": golden_dpmpp, "t5_encoder": golden_t5}
    for name in (sys.argv[1:] or list(jobs)):
        torch.save(jobs[name](), os.path.join(OUT, name + ".pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
