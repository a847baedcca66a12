"""flash_attention() with the reference signature (wan/modules/attention.py:24-130), served by the sm_100a
tcgen05 kernel (csrc/attention_sm100.cu) through the C ABI.  No FA2/FA3/SDPA dispatch, no CPU path."""
import torch

import movii_b200 as mv

__all__ = ["flash_attention", "attention"]


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """q [B, Lq, N, 128], k/v [B, Lk, N, 128] -> [B, Lq, N, 128] in q's dtype.

    Same semantics as the reference for the cases the DiT uses: non-causal, no dropout, no window; per-sample
    key lengths honoured (keys beyond k_lens[i] are not attended), every query row computed (q_lens only trims
    in the reference's packed layout; padded query rows are produced here as they are there)."""
    if causal or dropout_p != 0. or tuple(window_size) != (-1, -1):
        raise NotImplementedError("movii_b200 attention: only non-causal, dropout-free, global attention (the DiT path)")
    if q.size(-1) != 128 or k.size(2) != q.size(2):
        raise NotImplementedError("movii_b200 attention: head_dim must be 128 and Nq == Nk")
    assert dtype in (torch.bfloat16,), "the sm_100a kernel computes in bf16"
    out_dtype = q.dtype
    b, lq, n, d = q.shape
    if q_scale is not None:
        q = q * q_scale
    qh, kh, vh = (t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16) for t in (q, k, v))
    out = torch.empty(b, lq, n, d, dtype=torch.bfloat16, device=q.device)
    for i in range(b):
        kl = k.size(1) if k_lens is None else int(k_lens[i])
        ql = lq if q_lens is None else int(q_lens[i])
        qi, ki, vi = qh[i].contiguous(), kh[i, :kl].contiguous(), vh[i, :kl].contiguous()
        mv.attention(qi[:ql], ki, vi, out[i, :ql], softmax_scale)
        if ql < lq:
            out[i, ql:].zero_()
    return out.type(out_dtype)


def attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
              window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, fa_version=None):
    """wan/modules/attention.py:133-179 dispatches to flash-attn when present; here it is the same kernel."""
    return flash_attention(q, k, v, q_lens, k_lens, dropout_p, softmax_scale, q_scale, causal, window_size,
                           deterministic, dtype, fa_version)
