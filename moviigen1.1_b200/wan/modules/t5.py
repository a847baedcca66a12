"""umT5-XXL text encoder on the sm_100a kernels (reference: wan/modules/t5.py).

Same module tree and parameter names as the reference's T5Encoder (`token_embedding`, `blocks.{i}.norm1`,
`blocks.{i}.attn.{q,k,v,o}`, `blocks.{i}.norm2`, `blocks.{i}.ffn.{gate.0,fc1,fc2}`,
`blocks.{i}.pos_embedding.embedding`, `norm`), so `models_t5_umt5-xxl-enc-bf16.pth` loads with
`load_state_dict`; the arithmetic runs through the C ABI:

    x  = mv_embed_gather(token_embedding, ids)                         fp32 residual stream [L, dim]
    per block:
      n  = mv_t5_rmsnorm(x, norm1)                                     bf16
      q|k|v = mv_gemm_bf16(n, W{q,k,v})  -> one [L, 3*dim_attn] buffer bf16 (three launches, no weight copies)
      a  = mv_t5_attention(q, k, v, rel-pos bias table, kv_len)        bf16, tcgen05 QK^T / PV, exact softmax
      x += mv_gemm_bf16(a, Wo, RESID_F32)
      n  = mv_t5_rmsnorm(x, norm2)
      h  = mv_mul_bf16(mv_gemm_bf16(n, Wfc1), mv_gemm_bf16(n, Wgate, GELU))
      x += mv_gemm_bf16(h, Wfc2, RESID_F32)
    out = mv_t5_rmsnorm(x, norm)                                       bf16 (what the DiT's text_embedding consumes)

Differences from the reference that do not change results: dropout is identity (eval); the residual stream is kept
in fp32 instead of bf16 (closer to the fp32 semantics); `T5EncoderModel.__call__` encodes only the valid prefix of
each prompt — padded keys are masked and padded query rows are discarded by the reference's `u[:v]` (t5.py:517), so
the kept rows are identical.  There is no CPU path: `t5_cpu=True` callers still get the GPU encoder.
The T5 decoder / seq2seq model is never instantiated by the reference pipeline (`encoder_only=True`, t5.py:490) and
is not built.
"""
import os
import logging
import math

import torch
import torch.nn as nn

import movii_b200 as mv

from .tokenizers import HuggingfaceTokenizer

__all__ = ["T5Model", "T5Encoder", "T5Decoder", "T5EncoderModel", "umt5_xxl"]

HEAD_DIM = 64  # the sm_100a T5 attention kernel is specialised for umT5's 64-wide heads


class T5LayerNorm(nn.Module):
    """Parameter holder (t5.py:53-66); applied by mv_t5_rmsnorm."""

    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))


class T5Attention(nn.Module):
    """Parameter holder (t5.py:69-84): bias-free q/k/v/o projections."""

    def __init__(self, dim, dim_attn, num_heads, dropout=0.1):
        assert dim_attn % num_heads == 0
        super().__init__()
        self.dim, self.dim_attn, self.num_heads, self.head_dim = dim, dim_attn, num_heads, dim_attn // num_heads
        self.q = nn.Linear(dim, dim_attn, bias=False)
        self.k = nn.Linear(dim, dim_attn, bias=False)
        self.v = nn.Linear(dim, dim_attn, bias=False)
        self.o = nn.Linear(dim_attn, dim, bias=False)


class T5FeedForward(nn.Module):
    """Parameter holder (t5.py:123-134): gate is a Sequential so the key is `gate.0.weight`."""

    def __init__(self, dim, dim_ffn, dropout=0.1):
        super().__init__()
        self.dim, self.dim_ffn = dim, dim_ffn
        self.gate = nn.Sequential(nn.Linear(dim, dim_ffn, bias=False), nn.Identity())
        self.fc1 = nn.Linear(dim, dim_ffn, bias=False)
        self.fc2 = nn.Linear(dim_ffn, dim, bias=False)


class T5RelativeEmbedding(nn.Module):
    """t5.py:221-264.  `table(lq, lk)` returns the bias as a per-head lookup over the key-query distance instead
    of the reference's materialised [1, N, lq, lk] tensor: fp32 [N, lq + lk - 1], entry (j - i) + (lq - 1)."""

    def __init__(self, num_buckets, num_heads, bidirectional, max_dist=128):
        super().__init__()
        self.num_buckets, self.num_heads, self.bidirectional, self.max_dist = num_buckets, num_heads, bidirectional, max_dist
        self.embedding = nn.Embedding(num_buckets, num_heads)

    def buckets(self, rel_pos):
        """Bucket index of each key-minus-query distance (host int64 tensor), t5.py:245-264."""
        nb = self.num_buckets
        if self.bidirectional:
            nb //= 2
            base = (rel_pos > 0).long() * nb
            dist = rel_pos.abs()
        else:
            base = torch.zeros_like(rel_pos)
            dist = (-rel_pos).clamp(min=0)
        exact = nb // 2
        far = exact + (torch.log(dist.float() / exact) / math.log(self.max_dist / exact) * (nb - exact)).long()
        far = far.clamp(max=nb - 1)
        return base + torch.where(dist < exact, dist, far)

    def table(self, lq, lk):
        idx = self.buckets(torch.arange(-(lq - 1), lk, dtype=torch.int64)).to(self.embedding.weight.device)
        return self.embedding.weight.detach().float()[idx].t().contiguous()


class T5SelfAttention(nn.Module):
    """Parameter holder for one encoder block (t5.py:144-175)."""

    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos=True, dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_buckets, self.shared_pos = num_heads, num_buckets, shared_pos
        self.norm1 = T5LayerNorm(dim)
        self.attn = T5Attention(dim, dim_attn, num_heads, dropout)
        self.norm2 = T5LayerNorm(dim)
        self.ffn = T5FeedForward(dim, dim_ffn, dropout)
        self.pos_embedding = None if shared_pos else T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True)


def _bf16_weight(lin):
    w = lin.weight.detach()
    if w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise RuntimeError("T5 encoder weights must be contiguous bf16 on the GPU (call .to(torch.bfloat16))")
    return w


class T5Engine:
    """Launch sequence of the encoder over a T5Encoder's parameters (no copies of the big matrices)."""

    def __init__(self, enc):
        self.enc = enc
        self.device = enc.token_embedding.weight.device
        if self.device.type != "cuda":
            raise RuntimeError("the umT5 encoder runs on the B200 only (no CPU path); move the module to cuda")
        if enc.dim_attn // enc.num_heads != HEAD_DIM:
            raise RuntimeError("mv_t5_attention supports head_dim 64 (umT5); got %d" % (enc.dim_attn // enc.num_heads))
        self.norm_w = [(b.norm1.weight.detach().float().contiguous(), b.norm2.weight.detach().float().contiguous())
                       for b in enc.blocks]
        self.final_w = enc.norm.weight.detach().float().contiguous()
        self._bias = {}
        self._ws = {}
        # One encode = 266 launches of ~20 us kernels: the Python / ctypes launch path (~20 us per call) is as long as the
        # kernels, so the second encode of a (rows, valid keys) shape is captured in a CUDA graph and replayed afterwards
        # (MOVII_T5_GRAPH=0: direct launches every time).
        self.use_graph = os.environ.get("MOVII_T5_GRAPH", "1") != "0"
        self._graphs = {}

    def bias_tables(self, L):
        if L not in self._bias:
            enc = self.enc
            if enc.shared_pos:
                t = enc.pos_embedding.table(L, L)
                self._bias[L] = [t] * len(enc.blocks)
            else:
                self._bias[L] = [b.pos_embedding.table(L, L) for b in enc.blocks]
        return self._bias[L]

    def workspace(self, L):
        if L not in self._ws:
            e, dev, bf = self.enc, self.device, torch.bfloat16
            self._ws[L] = dict(
                x=torch.empty(L, e.dim, dtype=torch.float32, device=dev),
                n=torch.empty(L, e.dim, dtype=bf, device=dev),
                qkv=torch.empty(L, 3 * e.dim_attn, dtype=bf, device=dev),
                a=torch.empty(L, e.dim_attn, dtype=bf, device=dev),
                f=torch.empty(L, e.dim_ffn, dtype=bf, device=dev),
                g=torch.empty(L, e.dim_ffn, dtype=bf, device=dev),
                out=torch.empty(L, e.dim, dtype=bf, device=dev))
        return self._ws[L]

    def forward(self, ids, kv_len=None):
        """ids: int64 [L] (L <= 512) on the device; keys >= kv_len are masked.  Returns bf16 [L, dim] (a workspace
        view: clone to keep it across calls)."""
        L = ids.numel()
        kv_len = L if kv_len is None else int(kv_len)
        if not self.use_graph or mv.timing_enabled():
            return self._forward(ids, kv_len)
        key = (L, kv_len, mv.CONFIG_EPOCH)
        ent = self._graphs.get(key)
        if ent is None:                                    # first encode of this shape: direct launches (also warms up)
            self._graphs[key] = "seen"
            return self._forward(ids, kv_len)
        if ent == "seen":
            try:
                static_ids = ids.contiguous().clone()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0 = mv.LAUNCHES
                with torch.cuda.graph(graph):
                    self._forward(static_ids, kv_len)
                ent = (graph, static_ids, mv.LAUNCHES - l0)
                self._graphs[key] = ent
            except Exception as ex:                           # capture refused: keep launching directly (same kernels)
                logging.warning("T5Engine: CUDA graph capture failed (%r); using direct launches", ex)
                self.use_graph = False
                return self._forward(ids, kv_len)
        graph, static_ids, n_launch = ent
        static_ids.copy_(ids)
        graph.replay()
        mv.count_replayed_launches(n_launch)
        return self.workspace(L)["out"]

    def _forward(self, ids, kv_len):
        enc = self.enc
        L = ids.numel()
        ws = self.workspace(L)
        x, n, qkv, a, f, g = ws["x"], ws["n"], ws["qkv"], ws["a"], ws["f"], ws["g"]
        A, H = enc.dim_attn, enc.num_heads
        q3, k3, v3 = (qkv[:, i * A:(i + 1) * A].unflatten(1, (H, HEAD_DIM)) for i in range(3))
        a3 = a.view(L, H, HEAD_DIM)
        bias = self.bias_tables(L)
        mv.embed_gather(_bf16_weight(enc.token_embedding), ids.contiguous(), x)
        for i, blk in enumerate(enc.blocks):
            w1, w2 = self.norm_w[i]
            mv.t5_rmsnorm(x, w1, n, blk.norm1.eps)
            for j, lin in enumerate((blk.attn.q, blk.attn.k, blk.attn.v)):
                mv.gemm(n, _bf16_weight(lin), None, qkv[:, j * A:(j + 1) * A], mv.MV_EPI_BF16)
            mv.t5_attention(q3, k3, v3, a3, bias[i], L - 1, kv_len)
            mv.gemm(a, _bf16_weight(blk.attn.o), None, x, mv.MV_EPI_RESID_F32)
            mv.t5_rmsnorm(x, w2, n, blk.norm2.eps)
            mv.gemm(n, _bf16_weight(blk.ffn.gate[0]), None, g, mv.MV_EPI_BF16_GELU)
            mv.gemm(n, _bf16_weight(blk.ffn.fc1), None, f, mv.MV_EPI_BF16)
            mv.mul_bf16(f, g, f)
            mv.gemm(f, _bf16_weight(blk.ffn.fc2), None, x, mv.MV_EPI_RESID_F32)
        mv.t5_rmsnorm(x, self.final_w, ws["out"], enc.norm.eps)
        return ws["out"]


class T5Encoder(nn.Module):
    """t5.py:267-312 with the reference's constructor; forward(ids, mask) -> [B, L, dim] bf16."""

    def __init__(self, vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=True,
                 dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = num_heads, num_layers, num_buckets, shared_pos
        self.token_embedding = vocab if isinstance(vocab, nn.Embedding) else nn.Embedding(vocab, dim)
        self.pos_embedding = T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True) if shared_pos else None
        self.blocks = nn.ModuleList([
            T5SelfAttention(dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos, dropout)
            for _ in range(num_layers)])
        self.norm = T5LayerNorm(dim)
        self._engine = None

    def engine(self):
        if self._engine is None or self._engine.device != self.token_embedding.weight.device:
            self._engine = T5Engine(self)
        return self._engine

    # the engine snapshots fp32 norm gains and the relative-position tables: any parameter change invalidates it
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._engine = None
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine = None
        return r

    @staticmethod
    def _prefix_len(mask_row, L):
        if mask_row is None:
            return L
        n = int(mask_row.gt(0).sum())
        if n < 1 or not bool(mask_row[:n].gt(0).all()):
            raise RuntimeError("T5Encoder: the key mask must be a non-empty prefix mask (tokenizer right-padding)")
        return n

    def forward(self, ids, mask=None):
        eng = self.engine()
        outs = []
        for b in range(ids.size(0)):
            n = self._prefix_len(None if mask is None else mask[b], ids.size(1))
            outs.append(eng.forward(ids[b].to(eng.device), n).clone())
        return torch.stack(outs)

    def encode_prefix(self, ids_row, n):
        """Only the n valid tokens of one prompt: identical to forward(ids, mask)[0, :n] (see module docstring)."""
        eng = self.engine()
        return eng.forward(ids_row[:n].to(eng.device), n).clone()


class T5Decoder(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("T5Decoder is never instantiated by the MoviiGen pipeline (encoder_only=True, "
                                  "wan/modules/t5.py:490) and is not built")


class T5Model(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("the seq2seq T5Model is never instantiated by the MoviiGen pipeline "
                                  "(encoder_only=True, wan/modules/t5.py:490) and is not built")


def _t5(name, encoder_only=False, decoder_only=False, return_tokenizer=False, tokenizer_kwargs={},
        dtype=torch.float32, device="cpu", **kwargs):
    """t5.py:415-453 for the encoder-only case."""
    if not encoder_only or decoder_only:
        raise NotImplementedError("only encoder_only=True is built (see T5Model / T5Decoder)")
    kwargs["vocab"] = kwargs.pop("vocab_size")
    kwargs["num_layers"] = kwargs.pop("encoder_layers")
    kwargs.pop("decoder_layers")
    with torch.device(device):
        model = T5Encoder(**kwargs)
    model = model.to(dtype=dtype, device=device)
    if return_tokenizer:
        return model, HuggingfaceTokenizer("google/%s" % name, **tokenizer_kwargs)
    return model


def umt5_xxl(**kwargs):
    """t5.py:456-469."""
    cfg = dict(vocab_size=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, encoder_layers=24,
               decoder_layers=24, num_buckets=32, shared_pos=False, dropout=0.1)
    cfg.update(**kwargs)
    return _t5("umt5-xxl", **cfg)


class T5EncoderModel:
    """t5.py:472-517: tokenizer + encoder; __call__(texts, device) -> list of [len_i, 4096] bf16."""

    def __init__(self, text_len, dtype=torch.bfloat16, device=None, checkpoint_path=None, tokenizer_path=None,
                 shard_fn=None):
        if dtype != torch.bfloat16:
            raise RuntimeError("the B200 umT5 encoder computes in bf16 (the reference default); got %s" % dtype)
        if shard_fn is not None:
            raise RuntimeError("t5_fsdp sharding is not supported: the bf16 encoder (11.4 GB) fits one B200")
        self.text_len, self.dtype = text_len, dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            # the reference's t5_cpu mode exists to save VRAM on 24-80 GB cards; a B200 has 180 GB
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.checkpoint_path, self.tokenizer_path = checkpoint_path, tokenizer_path
        model = umt5_xxl(encoder_only=True, return_tokenizer=False, dtype=dtype, device=self.device)
        model = model.eval().requires_grad_(False)
        if checkpoint_path is not None:
            logging.info("loading %s", checkpoint_path)
            model.load_state_dict(torch.load(checkpoint_path, map_location="cpu", weights_only=True))
        self.model = model
        self.tokenizer = HuggingfaceTokenizer(name=tokenizer_path, seq_len=text_len, clean="whitespace")

    def __call__(self, texts, device=None):
        ids, mask = self.tokenizer(texts, return_mask=True, add_special_tokens=True)
        seq_lens = mask.gt(0).sum(dim=1).tolist()
        ids = ids.to(self.device)
        return [self.model.encode_prefix(ids[i], int(n)) for i, n in enumerate(seq_lens)]
