"""Host-side orchestration of the DiT forward over the C-ABI kernels (movii_b200).

Python/PyTorch here only owns memory (packed weights, workspaces) and sequences the launches; all
arithmetic of a WanAttentionBlock (wan/modules/model.py:274-313 in the reference) runs in the sm_100a
kernels.  The launch order per block is documented in DESIGN.md §3.

Sequence parallelism (Ulysses, wan/distributed/xdit_context_parallel.py in the reference) plugs in through
`SeqParallel`: token rows are sharded across ranks, the self-attention core is computed per head group after
an all-to-all (see wan/distributed/ulysses.py).
"""

import torch

import movii_b200 as mv

BF16 = torch.bfloat16
F32 = torch.float32


def _bf16(t):
    return t.detach().to(BF16).contiguous()


def _f32(t):
    return t.detach().to(F32).contiguous()


def _bias(t):
    """Bias of a bf16-autocast Linear: autocast casts every tensor argument of F.linear to bf16, so the value added to
    the fp32 accumulator is the bf16-rounded bias (kept as fp32 storage for the kernels' epilogues)."""
    return t.detach().to(BF16).to(F32).contiguous()


class BlockWeights:
    """Packed, kernel-ready weights of one WanAttentionBlock (bf16 GEMM operands, fp32 everything else)."""

    def __init__(self, blk):
        sa, ca = blk.self_attn, blk.cross_attn
        self.w_qkv = torch.cat([_bf16(sa.q.weight), _bf16(sa.k.weight), _bf16(sa.v.weight)], 0).contiguous()
        self.b_qkv = torch.cat([_bias(sa.q.bias), _bias(sa.k.bias), _bias(sa.v.bias)], 0).contiguous()
        self.w_o, self.b_o = _bf16(sa.o.weight), _bias(sa.o.bias)
        self.g_q, self.g_k = _f32(sa.norm_q.weight), _f32(sa.norm_k.weight)
        self.w_cq, self.b_cq = _bf16(ca.q.weight), _bias(ca.q.bias)
        self.w_ckv = torch.cat([_bf16(ca.k.weight), _bf16(ca.v.weight)], 0).contiguous()
        self.b_ckv = torch.cat([_bias(ca.k.bias), _bias(ca.v.bias)], 0).contiguous()
        self.w_co, self.b_co = _bf16(ca.o.weight), _bias(ca.o.bias)
        self.g_cq, self.g_ck = _f32(ca.norm_q.weight), _f32(ca.norm_k.weight)
        if getattr(blk, "cross_attn_norm", True) and hasattr(blk.norm3, "weight") and blk.norm3.weight is not None:
            self.n3_w, self.n3_b = _f32(blk.norm3.weight), _f32(blk.norm3.bias)
        else:
            self.n3_w = self.n3_b = None
        self.w_1, self.b_1 = _bf16(blk.ffn[0].weight), _bias(blk.ffn[0].bias)
        self.w_2, self.b_2 = _bf16(blk.ffn[2].weight), _bias(blk.ffn[2].bias)
        self.mod = _f32(blk.modulation).view(6, -1)
        self.dim = self.w_o.shape[0]
        self.ffn_dim = self.w_1.shape[0]
        self.num_heads = blk.num_heads
        self.eps = blk.eps
        self.qk_norm = getattr(blk, "qk_norm", True)

    def rebind(self, blk):
        """Point the module's bf16 Linear weights at the packed storage (no duplicate copy of the 14B weights)."""
        C = self.dim
        sa, ca = blk.self_attn, blk.cross_attn
        pairs = [(sa.q, self.w_qkv[0:C]), (sa.k, self.w_qkv[C:2 * C]), (sa.v, self.w_qkv[2 * C:]), (sa.o, self.w_o),
                 (ca.q, self.w_cq), (ca.k, self.w_ckv[0:C]), (ca.v, self.w_ckv[C:]), (ca.o, self.w_co),
                 (blk.ffn[0], self.w_1), (blk.ffn[2], self.w_2)]
        for lin, view in pairs:
            if lin.weight.dtype == BF16 and lin.weight.device == view.device:
                lin.weight.data = view


class Workspace:
    """Activation buffers for `rows` local tokens (allocated once per shape, reused by every block)."""

    def __init__(self, rows, dim, ffn_dim, ctx_len, device):
        self.rows = rows
        self.x = torch.empty(rows, dim, dtype=F32, device=device)        # fp32 residual stream
        self.h = torch.empty(rows, dim, dtype=BF16, device=device)       # normalised / modulated GEMM input
        self.qkv = torch.empty(rows, 3 * dim, dtype=BF16, device=device)
        self.attn = torch.empty(rows, dim, dtype=BF16, device=device)
        self.ffn = torch.empty(rows, ffn_dim, dtype=BF16, device=device)
        self.ckv = torch.empty(ctx_len, 2 * dim, dtype=BF16, device=device)


def rope_cos_sin(freqs, grid, seq_len, start=0, rows=None, device=None):
    """[rows, d/2, 2] fp32 (cos, sin) table for tokens [start, start+rows) of a (F,H',W') grid padded to seq_len.

    `freqs` is the reference's complex128 table [1024, d/2] (model.py:474-479); the per-token multiplier is
    freq_f[f] | freq_h[h] | freq_w[w] (model.py:44-62), 1+0j for padding tokens (xdit_context_parallel.py:11-21)."""
    f, h, w = grid
    c = freqs.shape[1]
    parts = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    tab = torch.cat([
        parts[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        parts[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        parts[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, c)
    if seq_len > tab.shape[0]:
        pad = torch.ones(seq_len - tab.shape[0], c, dtype=tab.dtype, device=tab.device)
        tab = torch.cat([tab, pad])
    rows = seq_len - start if rows is None else rows
    tab = tab[start:start + rows]
    cs = torch.view_as_real(tab).to(F32).contiguous()
    return cs.to(device) if device is not None else cs


def self_attention_core(ws, rows, kv_rows, num_heads):
    """Local (P = 1) attention over the fused qkv buffer: q/k/v are column slabs of ws.qkv."""
    C = ws.attn.shape[1]
    ld = ws.qkv.stride(0)
    q = ws.qkv.as_strided((rows, num_heads, 128), (ld, 128, 1), ws.qkv.storage_offset())
    k = ws.qkv.as_strided((kv_rows, num_heads, 128), (ld, 128, 1), ws.qkv.storage_offset() + C)
    v = ws.qkv.as_strided((kv_rows, num_heads, 128), (ld, 128, 1), ws.qkv.storage_offset() + 2 * C)
    mv.attention(q, k, v, ws.attn[:rows].view(rows, num_heads, 128))


def block_forward(bw, ws, rows, e, cs, ctx, kv_rows, first_block=False, sp=None):
    """One WanAttentionBlock on ws.x[:rows] in place.  e: [6, C] fp32 = modulation + e0 (model.py:292-295);
    cs: RoPE table for these rows; ctx: [Lc, C] bf16 embedded text; kv_rows: keys of real tokens
    (flash_attention(k_lens=seq_lens), model.py:146-151).  sp: UlyssesGroup -> rows are this rank's token shard and
    the attention core runs head-parallel after an all-to-all (xdit_context_parallel.py:155-198)."""
    C, nh, eps = bw.dim, bw.num_heads, bw.eps
    x, h, qkv, attn = ws.x[:rows], ws.h[:rows], ws.qkv[:rows], ws.attn[:rows]
    if not bw.qk_norm:
        raise NotImplementedError("qk_norm=False is not a configuration the 14B model uses")
    # --- self-attention: x += o(attn(rope(rms(q)), rope(rms(k)), v)) * e2          model.py:298-302
    mv.ln_modulate(x, h, shift=e[0], scale=e[1], eps=eps, round_ln=first_block)
    if sp is None or sp.world == 1:
        mv.gemm(h, bw.w_qkv, bw.b_qkv, qkv, mv.MV_EPI_BF16)
        mv.qkv_norm_rope(qkv, bw.g_q, bw.g_k, cs, 128, eps)      # q and k in one launch
        self_attention_core(ws, rows, kv_rows, nh)
        mv.gemm(attn, bw.w_o, bw.b_o, x, mv.MV_EPI_RESID_F32, gate=e[2])
    else:
        from ..distributed.ulysses import sp_self_attention

        def qkv_gemm(i):      # column slab i of the fused QKV projection (same tiles, same bits as the single launch)
            if i is None:
                mv.gemm(h, bw.w_qkv, bw.b_qkv, qkv, mv.MV_EPI_BF16)
            else:
                b = None if bw.b_qkv is None else bw.b_qkv[i * C:(i + 1) * C]
                mv.gemm(h, bw.w_qkv[i * C:(i + 1) * C], b, qkv[:, i * C:(i + 1) * C], mv.MV_EPI_BF16)
        o_slabs = sp_self_attention(mv, sp, ws, rows, bw, cs, kv_rows, qkv_gemm=qkv_gemm)
        mv.gemm_ksplit(o_slabs, bw.w_o, bw.b_o, x, mv.MV_EPI_RESID_F32, gate=e[2])
    # --- cross-attention: x += o(attn(rms(q(norm3 x)), rms(k(ctx)), v(ctx)))        model.py:306,159-181
    if bw.n3_w is not None:
        mv.ln_modulate(x, h, weight=bw.n3_w, bias=bw.n3_b, eps=eps)
    else:
        raise NotImplementedError("cross_attn_norm=False (norm3 = Identity) is not a configuration the 14B model uses")
    cq = qkv[:, 0:C]
    mv.gemm(h, bw.w_cq, bw.b_cq, cq, mv.MV_EPI_BF16)
    mv.rmsnorm_rope(cq, bw.g_cq, None, 128, eps)
    Lc = ctx.shape[0]
    ckv = ws.ckv[:Lc]
    mv.gemm(ctx, bw.w_ckv, bw.b_ckv, ckv, mv.MV_EPI_BF16)
    mv.rmsnorm_rope(ckv[:, 0:C], bw.g_ck, None, 128, eps)
    ld, ldc = qkv.stride(0), ckv.stride(0)
    qv = qkv.as_strided((rows, nh, 128), (ld, 128, 1), qkv.storage_offset())
    kv = ckv.as_strided((Lc, nh, 128), (ldc, 128, 1), ckv.storage_offset())
    vv = ckv.as_strided((Lc, nh, 128), (ldc, 128, 1), ckv.storage_offset() + C)
    mv.attention(qv, kv, vv, attn.view(rows, nh, 128))
    mv.gemm(attn, bw.w_co, bw.b_co, x, mv.MV_EPI_RESID_F32, gate=None)
    # --- FFN: x += W2 gelu(W1 (LN(x)(1+e4)+e3)) * e5                                model.py:307-309
    mv.ln_modulate(x, h, shift=e[3], scale=e[4], eps=eps)
    f = ws.ffn[:rows]
    mv.gemm(h, bw.w_1, bw.b_1, f, mv.MV_EPI_BF16_GELU)
    mv.gemm(f, bw.w_2, bw.b_2, x, mv.MV_EPI_RESID_F32, gate=e[5])


class DitEngine:
    """Packed weights + workspaces of one WanModel; runs WanModel.forward for one sample at a time."""

    def __init__(self, model):
        self.model = model
        dev = model.patch_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("WanModel must live on a CUDA (B200) device: movii_b200 has no CPU path")
        self.device = dev
        self.dim, self.ffn_dim = model.dim, model.ffn_dim
        self.num_heads, self.eps = model.num_heads, model.eps
        if self.dim // self.num_heads != 128:
            raise NotImplementedError("the sm_100a attention kernel is specialised for head_dim 128 "
                                      "(got %d)" % (self.dim // self.num_heads))
        self.patch = tuple(model.patch_size)
        if self.patch[0] != 1:
            raise NotImplementedError("temporal patch size must be 1")
        self.blocks = []
        for blk in model.blocks:
            bw = BlockWeights(blk)
            bw.rebind(blk)
            self.blocks.append(bw)
        self.mods = torch.stack([bw.mod for bw in self.blocks]).contiguous()          # [n_layers, 6, C]
        pe = model.patch_embedding
        self.w_patch = _bf16(pe.weight).flatten(1).contiguous()
        self.b_patch = _bias(pe.bias)
        te = model.text_embedding
        self.w_t0, self.b_t0 = _bf16(te[0].weight), _bias(te[0].bias)
        self.w_t2, self.b_t2 = _bf16(te[2].weight), _bias(te[2].bias)
        tm, tp = model.time_embedding, model.time_projection
        self.w_e0, self.b_e0 = _f32(tm[0].weight), _f32(tm[0].bias)
        self.w_e2, self.b_e2 = _f32(tm[2].weight), _f32(tm[2].bias)
        self.w_p, self.b_p = _f32(tp[1].weight), _f32(tp[1].bias)
        self.w_head, self.b_head = _f32(model.head.head.weight), _f32(model.head.head.bias)
        self.head_mod = _f32(model.head.modulation).view(2, -1)
        self.text_len, self.text_dim = model.text_len, model.text_dim
        self.freq_dim, self.out_dim = model.freq_dim, model.out_dim
        self._ws = {}
        self._cs = {}
        dev = self.device
        self.sin = torch.empty(self.freq_dim, dtype=F32, device=dev)
        self.e_hidden = torch.empty(self.dim, dtype=F32, device=dev)
        self.e = torch.empty(self.dim, dtype=F32, device=dev)
        self.e0 = torch.empty(6 * self.dim, dtype=F32, device=dev)
        self.E = torch.empty_like(self.mods)                    # per-layer (modulation + e0) table of one forward
        self.em = torch.empty(2, self.dim, dtype=F32, device=dev)
        self.ctx_in = torch.zeros(self.text_len, self.text_dim, dtype=BF16, device=dev)
        self.ctx_h = torch.empty(self.text_len, self.dim, dtype=BF16, device=dev)
        self.ctx = torch.empty(self.text_len, self.dim, dtype=BF16, device=dev)

    # -- buffers ------------------------------------------------------------------------------
    def workspace(self, rows):
        ws = self._ws.get(rows)
        if ws is None:
            self._ws.clear()  # one live shape at a time: 720P/1080P workspaces are GBs
            ws = Workspace(rows, self.dim, self.ffn_dim, self.text_len, self.device)
            self._ws[rows] = ws
        return ws

    def rope_table(self, freqs, grid, seq_len, start, rows):
        key = (tuple(grid), seq_len, start, rows)
        cs = self._cs.get(key)
        if cs is None:
            self._cs.clear()
            cs = rope_cos_sin(freqs.cpu(), grid, seq_len, start, rows, self.device)
            self._cs[key] = cs
        return cs

    # -- pieces of WanModel.forward ----------------------------------------------------------------
    def embed_time(self, t):
        """model.py:541-545: e [C] and e0 [6, C], fp32."""
        t = t.reshape(-1)[:1].to(self.device)
        if t.dtype not in (torch.int64, torch.float32):
            t = t.to(torch.float32)
        mv.sinusoid_embed(t, self.sin)
        mv.linear_f32_vec(self.sin, self.w_e0, self.b_e0, self.e_hidden, act_in=0)
        mv.linear_f32_vec(self.e_hidden, self.w_e2, self.b_e2, self.e, act_in=1)
        mv.linear_f32_vec(self.e, self.w_p, self.b_p, self.e0, act_in=1)
        return self.e, self.e0.view(6, self.dim)

    def embed_text(self, context):
        """model.py:549-554: zero-pad to text_len THEN Linear-GELU-Linear (padding rows become bias rows)."""
        n = context.shape[0]
        if n > self.text_len:
            raise ValueError("context longer than text_len")
        self.ctx_in.zero_()
        self.ctx_in[:n].copy_(context)
        mv.gemm(self.ctx_in, self.w_t0, self.b_t0, self.ctx_h, mv.MV_EPI_BF16_GELU)
        mv.gemm(self.ctx_h, self.w_t2, self.b_t2, self.ctx, mv.MV_EPI_BF16)
        return self.ctx

    def embed_patches(self, latent, ws, start=0, rows=None):
        """model.py:529-538: patchify + Linear into ws.x rows [0, rows) for tokens [start, start+rows); zero rows
        beyond the real tokens.  Returns (grid, n_real_tokens)."""
        C_in, Fr, H, W = latent.shape
        ph, pw = self.patch[1], self.patch[2]
        grid = (Fr, H // ph, W // pw)
        L = grid[0] * grid[1] * grid[2]
        a = torch.empty(L, C_in * ph * pw, dtype=BF16, device=self.device)
        mv.patchify(latent.to(F32).contiguous(), a, (ph, pw))
        rows = ws.rows if rows is None else rows
        real = max(0, min(L - start, rows))
        if real > 0:
            mv.gemm(a[start:start + real], self.w_patch, self.b_patch, ws.x[:real], mv.MV_EPI_F32_ROUND)
        if real < rows:
            ws.x[real:rows].zero_()
        return grid, L

    def head(self, ws, rows, e, grid, out):
        em = mv.modulation_table(self.head_mod, e, self.em)                    # model.py:341
        mv.head_unpatchify(ws.x[:rows], em[0], em[1], self.w_head, self.b_head, out, grid,
                           (self.patch[1], self.patch[2]), self.eps)

    # -- WanModel.forward, one sample, P = 1 ----------------------------------------------------------
    def forward_single(self, latent, t, context, seq_len, freqs):
        ws = self.workspace(seq_len)
        grid, L = self.embed_patches(latent, ws)
        if L > seq_len:
            raise AssertionError("seq_len smaller than the token count")
        e, e0 = self.embed_time(t)
        ctx = self.embed_text(context)
        cs = self.rope_table(freqs, grid, seq_len, 0, seq_len)
        E = mv.modulation_table(self.mods, self.e0, self.E).view(len(self.blocks), 6, self.dim)   # model.py:292-295
        for i, bw in enumerate(self.blocks):
            block_forward(bw, ws, seq_len, E[i], cs, ctx, L, first_block=(i == 0))
        C_out = self.out_dim
        out = torch.empty(C_out, grid[0], grid[1] * self.patch[1], grid[2] * self.patch[2], dtype=F32,
                          device=self.device)
        self.head(ws, seq_len, e, grid, out)
        return out

    # -- WanModel.forward under Ulysses sequence parallelism (xdit_context_parallel.py:65-152) ---------
    def forward_single_sp(self, latent, t, context, seq_len, freqs, grp):
        from ..distributed.ulysses import token_range
        start, rows = token_range(seq_len, grp.world, grp.rank)
        ws = self.workspace(rows)
        # only this rank's tokens are embedded (the reference embeds the full sequence on every rank and then
        # chunks, :137-139 — same values, P times less work)
        grid, L = self.embed_patches(latent, ws, start, rows)
        if L > seq_len:
            raise AssertionError("seq_len smaller than the token count")
        e, e0 = self.embed_time(t)
        ctx = self.embed_text(context)
        cs = self.rope_table(freqs, grid, seq_len, start, rows)
        E = mv.modulation_table(self.mods, self.e0, self.E).view(len(self.blocks), 6, self.dim)
        for i, bw in enumerate(self.blocks):
            # the reference USP path does not un-pad: all seq_len keys are attended (:178-183 TODO)
            block_forward(bw, ws, rows, E[i], cs, ctx, seq_len, first_block=(i == 0), sp=grp)
        em = mv.modulation_table(self.head_mod, e, self.em)
        nout = self.w_head.shape[0]
        tok = torch.empty(rows, nout, dtype=F32, device=self.device)
        mv.head_tokens(ws.x[:rows], em[0], em[1], self.w_head, self.b_head, tok, self.eps)
        full = grp.all_gather_rows(tok)
        out = torch.empty(self.out_dim, grid[0], grid[1] * self.patch[1], grid[2] * self.patch[2], dtype=F32,
                          device=self.device)
        mv.unpatchify(full, out, grid, (self.patch[1], self.patch[2]))
        return out
