"""WanT2V with the reference's constructor / generate() surface (wan/text2video.py:29-271), running the denoising
loop on the B200-native DiT engine and the native WanVAE decoder.

Differences that are deliberate and documented:
  * the umT5 text encoder (wan/modules/t5.py, sm_100a kernels) is used when its checkpoint and tokenizer exist
    under checkpoint_dir (text2video.py:71-78); otherwise generate() takes pre-computed text embeddings
    (`context=`, `context_null=`) or, for smoke/bench runs, deterministic synthetic embeddings;
  * random-init DiT / VAE weights are used when the checkpoints are absent (there are none on the box);
  * no FSDP (weights are replicated), no model offload to the CPU.
"""
import logging
import math
import os
import random
import sys
import types

import torch
import torch.distributed as dist

from .modules.model import WanModel
from .modules.vae import WanVAE
from .utils.fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas, retrieve_timesteps
from .utils.fm_solvers_unipc import FlowUniPCMultistepScheduler


def synthetic_text_embedding(prompt, text_dim=4096, max_len=512, device="cpu"):
    """Deterministic stand-in for umT5(prompt): [len, text_dim] bf16, seeded by the prompt text."""
    seed = int.from_bytes(prompt.encode("utf-8")[:8].ljust(8, b"\0"), "little") % (2 ** 31)
    n = max(1, min(max_len, len(prompt.split()) + 1))
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, text_dim, generator=g).to(torch.bfloat16).to(device)


class WanT2V:

    def __init__(self, config, checkpoint_dir, device_id=0, rank=0, t5_fsdp=False, dit_fsdp=False, use_usp=False,
                 t5_cpu=False, model=None, vae=None, allow_random_init=False):
        """Reference signature (text2video.py:31-41) + three extensions: `model` / `vae` inject already-built modules
        (benchmarks, tests); `allow_random_init=True` lets MISSING checkpoint files fall back to random-init weights
        and synthetic text embeddings — without it a missing file raises FileNotFoundError like the reference does."""
        self.device = torch.device("cuda:%d" % device_id)
        self.allow_random_init = allow_random_init
        if t5_cpu:
            logging.info("t5_cpu=True: the umT5 encoder still runs on the GPU here (there is no CPU path)")
        self.config = config
        self.rank = rank
        self.t5_cpu = t5_cpu
        self.num_train_timesteps = config.num_train_timesteps
        self.param_dtype = config.param_dtype
        if t5_fsdp or dit_fsdp:
            from .distributed.fsdp import shard_model
            shard_model(None, device_id)  # raises: out of scope
        self.vae_stride = config.vae_stride
        self.patch_size = config.patch_size

        ckpt = checkpoint_dir or ""
        t5_pth = os.path.join(ckpt, config.t5_checkpoint)
        t5_tok = os.path.join(ckpt, config.t5_tokenizer)
        if os.path.isfile(t5_pth) and os.path.isdir(t5_tok):   # text2video.py:71-78
            from .modules.t5 import T5EncoderModel
            self.text_encoder = T5EncoderModel(text_len=config.text_len, dtype=config.t5_dtype, device=self.device,
                                               checkpoint_path=t5_pth, tokenizer_path=t5_tok)
        else:
            # generate() also accepts pre-computed embeddings (context=...), so a missing encoder only matters when a
            # prompt has to be encoded: encode_prompt() raises then unless allow_random_init
            logging.warning("no umT5 checkpoint/tokenizer under %r", ckpt)
            self.text_encoder = None
        vae_pth = os.path.join(ckpt, config.vae_checkpoint)
        if vae is not None:
            self.vae = vae
        elif os.path.isfile(vae_pth):
            self.vae = WanVAE(vae_pth=vae_pth, device=self.device)
        elif allow_random_init:
            logging.warning("no VAE checkpoint %r: random-init WanVAE", vae_pth)
            self.vae = WanVAE(vae_pth=None, device=self.device)
        else:
            raise FileNotFoundError("WanVAE checkpoint %r not found (pass allow_random_init=True for a random-init "
                                    "smoke run)" % vae_pth)
        if model is not None:
            self.model = model
        elif os.path.isfile(os.path.join(ckpt, "config.json")):
            logging.info("Creating WanModel from %s", ckpt)
            self.model = WanModel.from_pretrained(ckpt, device=self.device, dtype=torch.bfloat16)
        elif not allow_random_init:
            raise FileNotFoundError("no DiT checkpoint (config.json + *.safetensors) under %r (pass "
                                    "allow_random_init=True for a random-init smoke run)" % ckpt)
        else:
            logging.warning("no DiT checkpoint under %r: using random-init weights of the configured architecture", ckpt)
            self.model = WanModel(model_type="t2v", patch_size=config.patch_size, text_len=config.text_len,
                                  in_dim=16, dim=config.dim, ffn_dim=config.ffn_dim, freq_dim=config.freq_dim,
                                  text_dim=4096, out_dim=16, num_heads=config.num_heads,
                                  num_layers=config.num_layers, window_size=config.window_size,
                                  qk_norm=config.qk_norm, cross_attn_norm=config.cross_attn_norm, eps=config.eps,
                                  device=self.device, dtype=torch.bfloat16)
            torch.nn.init.normal_(self.model.head.head.weight, std=0.02)  # the zero-init head would output 0
        self.model.eval().requires_grad_(False)

        if use_usp:
            from xfuser.core.distributed import get_sequence_parallel_world_size

            from .distributed.xdit_context_parallel import usp_attn_forward, usp_dit_forward
            for block in self.model.blocks:
                block.self_attn.forward = types.MethodType(usp_attn_forward, block.self_attn)
            self.model.forward = types.MethodType(usp_dit_forward, self.model)
            self.sp_size = get_sequence_parallel_world_size()
        else:
            self.sp_size = 1
        if dist.is_initialized():
            dist.barrier()
        self.model.to(self.device)
        self.sample_neg_prompt = config.sample_neg_prompt

    # ------------------------------------------------------------------------------------------------
    def encode_prompt(self, prompt):
        if self.text_encoder is not None:                      # text2video.py:174-184 (always on the GPU here)
            return self.text_encoder([prompt], self.device)
        if not self.allow_random_init:
            raise FileNotFoundError("the umT5 checkpoint / tokenizer is missing: pass context= / context_null= "
                                    "embeddings to generate(), or allow_random_init=True for synthetic ones")
        return [synthetic_text_embedding(prompt, 4096, self.config.text_len, self.device)]

    def denoise_step(self, scheduler, latent, t, context, context_null, seq_len, guide_scale):
        """One iteration of the reference's sampling loop (text2video.py:233-254): two DiT forwards, classifier-free
        guidance, scheduler step.  This is the unit of the `denoising-steps/sec` metric."""
        timestep = torch.stack([t])
        cond = self.model([latent], t=timestep, context=context, seq_len=seq_len)[0]
        uncond = self.model([latent], t=timestep, context=context_null, seq_len=seq_len)[0]
        if hasattr(scheduler, "step_cfg"):      # UniPC: CFG + scheduler update fused into one kernel
            return scheduler.step_cfg(cond, uncond, guide_scale, t, latent)[0]
        noise_pred = uncond + guide_scale * (cond - uncond)
        return scheduler.step(noise_pred.unsqueeze(0), t, latent.unsqueeze(0), return_dict=False)[0].squeeze(0)

    def latent_geometry(self, size, frame_num):
        """text2video.py:160-166: latent shape and (sp-padded) token count."""
        z = self.vae.model.z_dim
        shape = (z, (frame_num - 1) // self.vae_stride[0] + 1, size[1] // self.vae_stride[1],
                 size[0] // self.vae_stride[2])
        seq_len = math.ceil((shape[2] * shape[3]) / (self.patch_size[1] * self.patch_size[2]) * shape[1] /
                            self.sp_size) * self.sp_size
        return shape, seq_len

    def generate(self, input_prompt, size=(1280, 720), frame_num=81, shift=5.0, sample_solver="unipc",
                 sampling_steps=50, guide_scale=5.0, n_prompt="", seed=-1, offload_model=True, context=None,
                 context_null=None, decode=True):
        if offload_model:
            logging.debug("offload_model=True is ignored: the bf16 DiT (28.6 GB) and the VAE stay resident on the B200")
        target_shape, seq_len = self.latent_geometry(size, frame_num)
        if n_prompt == "":
            n_prompt = self.sample_neg_prompt
        seed = seed if seed >= 0 else random.randint(0, sys.maxsize)
        seed_g = torch.Generator(device=self.device)
        seed_g.manual_seed(seed)
        if context is None:
            context = self.encode_prompt(input_prompt)
        if context_null is None:
            context_null = self.encode_prompt(n_prompt)
        context = [c.to(self.device) for c in context]
        context_null = [c.to(self.device) for c in context_null]
        latent = torch.randn(*target_shape, dtype=torch.float32, device=self.device, generator=seed_g)

        with torch.no_grad():
            if sample_solver == "unipc":                      # text2video.py:206-213
                scheduler = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                        use_dynamic_shifting=False)
                scheduler.set_timesteps(sampling_steps, device=self.device, shift=shift)
                timesteps = scheduler.timesteps
            elif sample_solver == "dpm++":                    # text2video.py:214-223
                scheduler = FlowDPMSolverMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1,
                                                            use_dynamic_shifting=False)
                timesteps, _ = retrieve_timesteps(scheduler, device=self.device,
                                                  sigmas=get_sampling_sigmas(sampling_steps, shift))
            else:
                raise NotImplementedError("Unsupported solver.")
            for t in timesteps:
                latent = self.denoise_step(scheduler, latent, t, context, context_null, seq_len, guide_scale)
            videos = None
            if self.rank == 0 and decode:
                videos = self.vae.decode([latent])
            elif self.rank == 0:
                videos = [latent]
        if dist.is_initialized():
            dist.barrier()
        return videos[0] if self.rank == 0 else None
