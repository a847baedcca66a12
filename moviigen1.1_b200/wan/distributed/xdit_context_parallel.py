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
    """WanSelfAttention.forward under SP is fused into the engine's block loop (engine.block_forward with
    attn_core=sp_self_attention); a standalone per-module call is not a code path of the product."""
    raise NotImplementedError("usp_attn_forward is executed inside DitEngine.forward_single_sp; "
                              "call the model's forward (usp_dit_forward) instead")
