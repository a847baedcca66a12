"""The two seams the reference rebinds for sequence parallelism (wan/text2video.py:90-103):
`usp_attn_forward` (per-block self-attention) and `usp_dit_forward` (whole model).  Here they route to the
B200-native Ulysses path (wan/distributed/ulysses.py + wan/modules/engine.py)."""
import torch

from xfuser.core.distributed import get_sp_group


def usp_dit_forward(self, x, t, context, seq_len, clip_fea=None, y=None, guidance=None):
    """WanModel.forward under Ulysses SP (reference: xdit_context_parallel.py:65-152).  Every rank returns the full
    output (the final all_gather, :148)."""
    if clip_fea is not None or y is not None:
        raise NotImplementedError("i2v inputs are out of scope")
    eng = self.engine()
    grp = get_sp_group().ulysses
    outs = []
    for i, u in enumerate(x):
        ti = t[i] if t.dim() > 0 and t.numel() > 1 else t
        outs.append(eng.forward_single_sp(u, ti, context[i], seq_len, self.freqs, grp))
    return outs


def usp_attn_forward(self, x, seq_lens, grid_sizes, freqs, dtype=torch.bfloat16):
    """WanSelfAttention.forward under Ulysses SP as a standalone call (reference: xdit_context_parallel.py:155-198):
    x [B, L/P, C] is this rank's token shard (already normalised + modulated); q/k/v Linear, full-row RMSNorm, RoPE
    at the shard's GLOBAL positions (:51-57), head-scatter / sequence-gather exchange, attention over all L keys for
    40/P heads (the reference does not un-pad here, :178-183), return exchange, o Linear.  Returns [B, L/P, C] bf16.
    DitEngine.forward_single_sp runs the same launches inside its block loop (engine.block_forward(sp=...))."""
    from ..modules import engine as E
    from ..modules.model import _OwnerRef
    from .ulysses import sp_self_attention
    mv = E.mv
    blk = _OwnerRef.get(self)
    bw = blk._weights()
    grp = get_sp_group().ulysses
    B, rows, C = x.shape
    total = rows * grp.world
    ws = E.Workspace(rows, C, 8, 8, x.device)
    outs = []
    for i in range(B):
        grid = tuple(int(v) for v in grid_sizes[i].tolist())
        cs = E.rope_cos_sin(freqs.cpu(), grid, total, grp.rank * rows, rows, x.device)
        ws.h.copy_(x[i])
        mv.gemm(ws.h, bw.w_qkv, bw.b_qkv, ws.qkv, mv.MV_EPI_BF16)
        if grp.world == 1:
            mv.qkv_norm_rope(ws.qkv, bw.g_q, bw.g_k, cs, 128, bw.eps)
            E.self_attention_core(ws, rows, total, self.num_heads)
            y = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device)
            mv.gemm(ws.attn, bw.w_o, bw.b_o, y, mv.MV_EPI_BF16)
        else:
            o_slabs = sp_self_attention(mv, grp, ws, rows, bw, cs, total)
            y = torch.empty(rows, C, dtype=torch.bfloat16, device=x.device)
            mv.gemm_ksplit(o_slabs, bw.w_o, bw.b_o, y, mv.MV_EPI_BF16)
        outs.append(y)
    return torch.stack(outs)
