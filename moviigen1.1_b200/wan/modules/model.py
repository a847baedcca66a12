"""WanModel with the reference's module tree, parameter names and forward() signatures
(wan/modules/model.py in ZulutionAI/MoviiGen1.1), computed by the B200-native engine.

The nn.Module classes below only *hold parameters* under the reference's names (so reference
state dicts load and `types.MethodType` rebinding as in wan/text2video.py:90-103 keeps working);
their forward() methods hand raw device pointers to the sm_100a kernels via `engine.py`.
There is no eager-PyTorch implementation of the arithmetic here and no CPU path.
"""
import json
import math
import os
import weakref

import torch
import torch.nn as nn

from . import engine as E
from .attention import flash_attention

__all__ = ["WanModel"]


def sinusoidal_embedding_1d(dim, position):
    """model.py:15-25 (host-side helper kept for API compatibility; the engine uses mv_sinusoid_embed)."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """model.py:28-36: unit complex128 rotations [max_seq_len, dim/2]."""
    assert dim % 2 == 0
    ang = torch.outer(torch.arange(max_seq_len, dtype=torch.float64),
                      1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64).div(dim)))
    return torch.polar(torch.ones_like(ang), ang)


class WanRMSNorm(nn.Module):
    """Parameter holder for the full-row q/k RMSNorm (model.py:70-86); applied by mv_rmsnorm_rope."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        shp = x.shape
        y = x.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
        E.mv.rmsnorm_rope(y, self.weight.detach().float().contiguous(), None, min(128, shp[-1]), self.eps)
        return y.view(shp)


class WanLayerNorm(nn.LayerNorm):
    """model.py:89-99; the engine fuses it with the adaLN modulation (mv_ln_modulate)."""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x):
        shp = x.shape
        out = torch.empty(x.numel() // shp[-1], shp[-1], dtype=torch.bfloat16, device=x.device)
        w = self.weight.detach().float().contiguous() if self.elementwise_affine else None
        b = self.bias.detach().float().contiguous() if self.elementwise_affine else None
        E.mv.ln_modulate(x.reshape(-1, shp[-1]).float().contiguous(), out, weight=w, bias=b, eps=self.eps)
        return out.view(shp).type_as(x)


class WanSelfAttention(nn.Module):

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        assert dim % num_heads == 0
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.window_size, self.qk_norm, self.eps = window_size, qk_norm, eps
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_owner_ref", None)      # weak link to the owning block: re-created by the block's __setstate__
        return state

    def forward(self, x, seq_lens, grid_sizes, freqs):
        """model.py:127-156 for a standalone call: x [B, L, C] (already normalised + modulated)."""
        blk = _OwnerRef.get(self)
        return blk._engine_self_attention(x, seq_lens, grid_sizes, freqs)


class WanT2VCrossAttention(WanSelfAttention):

    def forward(self, x, context, context_lens):
        blk = _OwnerRef.get(self)
        return blk._engine_cross_attention(x, context, context_lens)


WAN_CROSSATTENTION_CLASSES = {"t2v_cross_attn": WanT2VCrossAttention}


class _OwnerRef:
    """Maps an attention sub-module to its owning block through a weak reference stored on the child (no module
    cycle, nothing kept alive after `del model`; deepcopy / pickle re-create it in WanAttentionBlock.__setstate__)."""

    @staticmethod
    def set(child, owner):
        object.__setattr__(child, "_owner_ref", weakref.ref(owner))

    @staticmethod
    def get(child):
        ref = getattr(child, "_owner_ref", None)
        owner = ref() if ref is not None else None
        if owner is None:
            raise RuntimeError("attention module is not attached to a live WanAttentionBlock")
        return owner


class WanAttentionBlock(nn.Module):

    def __init__(self, cross_attn_type, dim, ffn_dim, num_heads, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=False, eps=1e-6):
        super().__init__()
        if cross_attn_type not in WAN_CROSSATTENTION_CLASSES:
            raise NotImplementedError("only 't2v_cross_attn' is in scope (i2v is dead code for WAN_CONFIGS)")
        self.dim, self.ffn_dim, self.num_heads = dim, ffn_dim, num_heads
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = window_size, qk_norm, cross_attn_norm, eps
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm else nn.Identity()
        self.cross_attn = WAN_CROSSATTENTION_CLASSES[cross_attn_type](dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        _OwnerRef.set(self.self_attn, self)
        _OwnerRef.set(self.cross_attn, self)
        self._packed = None

    def __setstate__(self, state):
        state = dict(state)
        state["_packed"] = None            # packed weights are rebuilt lazily for the copy
        super().__setstate__(state)
        _OwnerRef.set(self.self_attn, self)
        _OwnerRef.set(self.cross_attn, self)

    # -- engine plumbing -----------------------------------------------------------------------
    def _weights(self):
        if self._packed is None:
            self._packed = E.BlockWeights(self)
        return self._packed

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens):
        """model.py:274-313.  x [B, L, C] (fp32, or bf16 for the first block), e [B, 6, C] fp32,
        context [B, Lc, C]; returns fp32 [B, L, C]."""
        assert e.dtype == torch.float32
        if context_lens is not None:
            raise NotImplementedError("context_lens is always None on the t2v path (model.py:548)")
        bw = self._weights()
        B, L, C = x.shape
        dev = x.device
        ws = E.Workspace(L, C, self.ffn_dim, context.shape[1], dev)
        outs = []
        for i in range(B):
            grid = tuple(int(v) for v in grid_sizes[i].tolist())
            cs = E.rope_cos_sin(freqs.cpu(), grid, L, 0, L, dev)
            ws.x.copy_(x[i])
            em = (bw.mod + e[i].to(dev)).contiguous()
            ctx = context[i].to(torch.bfloat16).contiguous()
            E.block_forward(bw, ws, L, em, cs, ctx, int(seq_lens[i]), first_block=(x.dtype == torch.bfloat16))
            outs.append(ws.x.clone())
        return torch.stack(outs)

    def _engine_self_attention(self, x, seq_lens, grid_sizes, freqs):
        bw = self._weights()
        B, L, C = x.shape
        ws = E.Workspace(L, C, 8, 8, x.device)
        outs = []
        for i in range(B):
            grid = tuple(int(v) for v in grid_sizes[i].tolist())
            cs = E.rope_cos_sin(freqs.cpu(), grid, L, 0, L, x.device)
            ws.h.copy_(x[i])
            E.mv.gemm(ws.h, bw.w_qkv, bw.b_qkv, ws.qkv, E.mv.MV_EPI_BF16)
            E.mv.qkv_norm_rope(ws.qkv, bw.g_q, bw.g_k, cs, 128, bw.eps)
            E.self_attention_core(ws, L, int(seq_lens[i]), self.num_heads)
            y = torch.empty(L, C, dtype=torch.bfloat16, device=x.device)
            E.mv.gemm(ws.attn, bw.w_o, bw.b_o, y, E.mv.MV_EPI_BF16)
            outs.append(y)
        return torch.stack(outs)

    def _engine_cross_attention(self, x, context, context_lens):
        if context_lens is not None:
            raise NotImplementedError("context_lens is always None on the t2v path")
        bw = self._weights()
        B, L, C = x.shape
        nh = self.num_heads
        outs = []
        for i in range(B):
            h = x[i].to(torch.bfloat16).contiguous()
            ctx = context[i].to(torch.bfloat16).contiguous()
            q = torch.empty(L, C, dtype=torch.bfloat16, device=x.device)
            kv = torch.empty(ctx.shape[0], 2 * C, dtype=torch.bfloat16, device=x.device)
            E.mv.gemm(h, bw.w_cq, bw.b_cq, q, E.mv.MV_EPI_BF16)
            E.mv.rmsnorm_rope(q, bw.g_cq, None, 128, bw.eps)
            E.mv.gemm(ctx, bw.w_ckv, bw.b_ckv, kv, E.mv.MV_EPI_BF16)
            E.mv.rmsnorm_rope(kv[:, :C], bw.g_ck, None, 128, bw.eps)
            a = torch.empty(L, C, dtype=torch.bfloat16, device=x.device)
            kk = kv.as_strided((ctx.shape[0], nh, 128), (2 * C, 128, 1), 0)
            vv = kv.as_strided((ctx.shape[0], nh, 128), (2 * C, 128, 1), C)
            E.mv.attention(q.view(L, nh, 128), kk, vv, a.view(L, nh, 128))
            y = torch.empty(L, C, dtype=torch.bfloat16, device=x.device)
            E.mv.gemm(a, bw.w_co, bw.b_co, y, E.mv.MV_EPI_BF16)
            outs.append(y)
        return torch.stack(outs)


class Head(nn.Module):

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim, self.out_dim, self.patch_size, self.eps = dim, out_dim, patch_size, eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)


class WanModel(nn.Module):
    """Wan / MoviiGen T2V diffusion backbone (reference: wan/modules/model.py:359-633).

    `device` / `dtype` are extensions: they let the 14B model be created directly in bf16 on the GPU
    (a CPU fp32 construction would need 57 GB of host memory)."""

    ignore_for_config = ["patch_size", "cross_attn_norm", "qk_norm", "text_dim", "window_size"]
    _no_split_modules = ["WanAttentionBlock"]

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, device=None, dtype=None, init=True):
        super().__init__()
        if model_type != "t2v":
            raise NotImplementedError("only model_type='t2v' is in scope (WAN_CONFIGS has no i2v entry)")
        self.model_type = model_type
        self.patch_size, self.text_len, self.in_dim, self.dim, self.ffn_dim = patch_size, text_len, in_dim, dim, ffn_dim
        self.freq_dim, self.text_dim, self.out_dim, self.num_heads, self.num_layers = (freq_dim, text_dim, out_dim,
                                                                                       num_heads, num_layers)
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = window_size, qk_norm, cross_attn_norm, eps
        self.config = dict(model_type=model_type, text_len=text_len, in_dim=in_dim, dim=dim, ffn_dim=ffn_dim,
                           freq_dim=freq_dim, out_dim=out_dim, num_heads=num_heads, num_layers=num_layers, eps=eps)
        fk = {}
        if device is not None:
            fk["device"] = device
        if dtype is not None:
            fk["dtype"] = dtype
        prev = torch.get_default_dtype()
        ctx_dev = torch.device(device) if device is not None else None
        try:
            if dtype is not None:
                torch.set_default_dtype(dtype)
            with (ctx_dev if ctx_dev is not None else torch.device("cpu")):
                self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
                self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"),
                                                    nn.Linear(dim, dim))
                self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
                self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
                self.blocks = nn.ModuleList([
                    WanAttentionBlock("t2v_cross_attn", dim, ffn_dim, num_heads, window_size, qk_norm,
                                      cross_attn_norm, eps) for _ in range(num_layers)])
                self.head = Head(dim, out_dim, patch_size, eps)
        finally:
            torch.set_default_dtype(prev)
        assert (dim % num_heads) == 0 and (dim // num_heads) % 2 == 0
        d = dim // num_heads
        # not a registered buffer: must stay complex128 (model.py:471-479)
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1)
        self._engine = None
        self.keep_fp32_parameters()
        if init:
            self.init_weights()

    def keep_fp32_parameters(self):
        """The parameters the reference uses OUTSIDE bf16 autocast stay fp32 whatever `dtype` the model was created /
        loaded with: time_embedding, time_projection and the head run inside `amp.autocast(dtype=float32)` regions
        (model.py:340-343,541-545), the adaLN `modulation` tables are added in fp32 (:292-295), RMSNorm / norm3 gains
        multiply fp32 values (:83-86,97-99).  Only the operands of the bf16-autocast Linears / the patch Conv3d may be
        stored in bf16 (autocast rounds them to bf16 at every call anyway, biases included)."""
        for m in (self.time_embedding, self.time_projection, self.head):
            m.float()
        for b in self.blocks:
            b.modulation.data = b.modulation.data.float()
            for n in (b.norm3, b.self_attn.norm_q, b.self_attn.norm_k, b.cross_attn.norm_q, b.cross_attn.norm_k):
                n.float()
        return self

    # -- engine ------------------------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            self._engine = E.DitEngine(self)
        return self._engine

    def invalidate_engine(self):
        """Call after changing parameters (load_state_dict does it automatically)."""
        self._engine = None
        for b in self.blocks:
            b._packed = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_engine()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate_engine()
        return r

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None):
        """model.py:486-579.  x: list of [C_in, F, H, W] fp32 latents; t: [B] timesteps; context: list of
        [L_txt, text_dim] text embeddings; returns a list of fp32 [C_out, F, H, W]."""
        if clip_fea is not None or y is not None:
            raise NotImplementedError("i2v inputs (clip_fea / y) are out of scope")
        eng = self.engine()
        outs = []
        for i, u in enumerate(x):
            ti = t[i] if t.dim() > 0 and t.numel() > 1 else t
            outs.append(eng.forward_single(u, ti, context[i], seq_len, self.freqs))
        return outs

    def unpatchify(self, x, grid_sizes):
        """model.py:581-609 (host-side helper; the engine fuses this into mv_head_unpatchify)."""
        c = self.out_dim
        out = []
        for u, v in zip(x, grid_sizes.tolist()):
            u = u[:math.prod(v)].view(*v, *self.patch_size, c)
            u = torch.einsum("fhwpqrc->cfphqwr", u)
            out.append(u.reshape(c, *[i * j for i, j in zip(v, self.patch_size)]))
        return out

    def init_weights(self):
        """Same distributions as model.py:611-633 (Xavier-uniform Linears, N(0,.02) embeddings, zero head)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for seq in (self.text_embedding, self.time_embedding):
            for m in seq.modules():
                if isinstance(m, nn.Linear):
                    nn.init.normal_(m.weight, std=.02)
        nn.init.zeros_(self.head.head.weight)
        self.invalidate_engine()

    # -- loading (diffusers-free replacement for ModelMixin.from_pretrained, text2video.py:87) ----------
    @classmethod
    def from_pretrained(cls, checkpoint_dir, device=None, dtype=None):
        with open(os.path.join(checkpoint_dir, "config.json")) as fh:
            cfg = json.load(fh)
        keys = ("model_type", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "out_dim", "num_heads",
                "num_layers", "eps")
        model = cls(**{k: cfg[k] for k in keys if k in cfg}, device=device, dtype=dtype, init=False)
        from safetensors.torch import load_file
        files = sorted(f for f in os.listdir(checkpoint_dir) if f.endswith(".safetensors"))
        if not files:
            raise FileNotFoundError("no *.safetensors in %s" % checkpoint_dir)
        sd = {}
        for f in files:
            sd.update(load_file(os.path.join(checkpoint_dir, f)))
        model.load_state_dict(sd, strict=True)
        return model
