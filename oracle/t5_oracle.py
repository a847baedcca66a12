"""CPU oracle for the umT5 text encoder — TEST INFRASTRUCTURE ONLY.

A plain-torch fp32 restatement of the reference's T5Encoder forward
(/root/reference/wan/modules/t5.py) written functionally over a reference-named state dict
(`token_embedding.weight`, `blocks.{i}.{norm1,norm2}.weight`, `blocks.{i}.attn.{q,k,v,o}.weight`,
`blocks.{i}.ffn.{gate.0,fc1,fc2}.weight`, `blocks.{i}.pos_embedding.embedding.weight` (or the shared
`pos_embedding.embedding.weight`), `norm.weight`).  Only tests/ and bench/smoke checker legs may
import it; the product path never does.

Pinning: the reference ships no golden vectors; oracle/make_golden.py runs the reference's own
T5Encoder (imported from /root/reference in the build container) on a seeded tiny configuration and
stores ids/mask/output under tests/golden/t5_encoder.pt (tests/test_oracle_golden.py).

`rb` emulates the bf16 cast points of the reference's bf16 model (`T5EncoderModel(dtype=bfloat16)`,
t5.py:472-497): identity -> fp32 semantics, `bf16_rt` -> the arithmetic contract of the CUDA path
(bf16 GEMM operands / outputs, fp32 accumulation, fp32 softmax, fp32 residual stream).
"""
import math

import torch
import torch.nn.functional as F


def ident(x):
    return x


def bf16_rt(x):
    return x.to(torch.bfloat16).to(torch.float32)


def gelu_tanh(x):
    """t5.py:46-50."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def t5_layer_norm(x, weight, eps=1e-6, rb=ident):
    """t5.py:61-66: x * rsqrt(mean(x^2) + eps), cast to the weight dtype, times weight."""
    y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    return rb(rb(y) * weight)


def relative_position_bucket(rel_pos, num_buckets, bidirectional=True, max_dist=128):
    """t5.py:245-264; rel_pos = key index - query index (int64)."""
    if bidirectional:
        num_buckets = num_buckets // 2
        rel_buckets = (rel_pos > 0).long() * num_buckets
        rel_pos = torch.abs(rel_pos)
    else:
        rel_buckets = 0
        rel_pos = -torch.min(rel_pos, torch.zeros_like(rel_pos))
    max_exact = num_buckets // 2
    rel_pos_large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact) *
                                 (num_buckets - max_exact)).long()
    rel_pos_large = torch.min(rel_pos_large, torch.full_like(rel_pos_large, num_buckets - 1))
    rel_buckets = rel_buckets + torch.where(rel_pos < max_exact, rel_pos, rel_pos_large)
    return rel_buckets


def relative_bias(embedding, lq, lk, num_buckets):
    """t5.py:233-243: [H, lq, lk] additive bias from the bucket embedding [num_buckets, H]."""
    rel_pos = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    buckets = relative_position_bucket(rel_pos, num_buckets)
    return embedding[buckets].permute(2, 0, 1).contiguous()


def t5_attention(x, sd, prefix, num_heads, mask, pos_bias, rb=ident):
    """t5.py:86-120 for self-attention (context = x); x [L, C], mask [L] (1 = token) or None."""
    L = x.shape[0]
    q = rb(rb(x) @ rb(sd[prefix + "q.weight"].float()).t()).view(L, num_heads, -1)
    k = rb(rb(x) @ rb(sd[prefix + "k.weight"].float()).t()).view(L, num_heads, -1)
    v = rb(rb(x) @ rb(sd[prefix + "v.weight"].float()).t()).view(L, num_heads, -1)
    attn = torch.einsum("inc,jnc->nij", q, k)                 # no 1/sqrt(d) (t5.py:111)
    if pos_bias is not None:
        attn = attn + pos_bias
    if mask is not None:
        attn = attn.masked_fill(mask.view(1, 1, -1) == 0, torch.finfo(torch.float32).min)
    attn = F.softmax(attn.float(), dim=-1)
    o = torch.einsum("nij,jnc->inc", attn, v).reshape(L, -1)
    return rb(rb(o) @ rb(sd[prefix + "o.weight"].float()).t())


def t5_ffn(x, sd, prefix, rb=ident):
    """t5.py:136-141: fc2(fc1(x) * gelu(gate(x)))."""
    g = rb(gelu_tanh(rb(x) @ rb(sd[prefix + "gate.0.weight"].float()).t()))
    f = rb(rb(x) @ rb(sd[prefix + "fc1.weight"].float()).t())
    return rb(rb(f * g) @ rb(sd[prefix + "fc2.weight"].float()).t())


def t5_encoder_forward(sd, ids, mask, num_heads, num_buckets, shared_pos=False, rb=ident):
    """t5.py:303-312 (dropout is identity in eval): ids [L] int64, mask [L] -> [L, dim] fp32."""
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    x = sd["token_embedding.weight"].float()[ids]
    L = x.shape[0]
    shared = relative_bias(sd["pos_embedding.embedding.weight"].float(), L, L, num_buckets) if shared_pos else None
    for i in range(n_layers):
        p = "blocks.%d." % i
        e = shared if shared_pos else relative_bias(sd[p + "pos_embedding.embedding.weight"].float(), L, L,
                                                    num_buckets)
        x = x + t5_attention(t5_layer_norm(x, sd[p + "norm1.weight"].float(), rb=rb), sd, p + "attn.", num_heads,
                             mask, e, rb=rb)
        x = x + t5_ffn(t5_layer_norm(x, sd[p + "norm2.weight"].float(), rb=rb), sd, p + "ffn.", rb=rb)
    return t5_layer_norm(x, sd["norm.weight"].float(), rb=rb)
