"""CPU oracle for the DiT hot path — TEST INFRASTRUCTURE ONLY.

A plain-torch (fp32 / fp64, CPU) restatement of the reference's WanModel forward
(/root/reference/wan/modules/model.py, attention.py) written functionally over a reference-named
state dict.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (moviigen1.1_b200/) never does.

Pinning: the reference ships no tests or golden vectors.  This oracle is pinned
against outputs of the reference's own code, imported in the build container by
oracle/make_golden.py, and stored under tests/golden/ (tests/test_oracle_golden.py).

`rb` ("round to bf16") emulates the cast points torch.autocast('cuda', bf16) introduces in the
reference: identity -> exact fp32 reference semantics on CPU; `bf16_rt` ->
bf16-autocast semantics, the arithmetic contract of the CUDA kernels.
"""

import torch
import torch.nn.functional as F


def ident(x):
    return x


def bf16_rt(x):
    """bf16 round trip (what an autocast Linear does to its inputs / output)."""
    return x.to(torch.bfloat16).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# embeddings                                                             model.py:15-36
# ------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim, position):
    """model.py:15-25: cos|sin of position * 10000^(-i/half), computed in float64."""
    half = dim // 2
    position = position.to(torch.float64)
    freqs = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                      -torch.arange(half, dtype=torch.float64) / half)
    ang = torch.outer(position, freqs)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_angles(max_seq_len, dim, theta=10000.0):
    """model.py:28-36 (rope_params) as angles instead of unit complex numbers; float64 [max_seq_len, dim/2]."""
    inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64),
                          torch.arange(0, dim, 2, dtype=torch.float64) / dim)
    return torch.outer(torch.arange(max_seq_len, dtype=torch.float64), inv)


def rope_axis_dims(head_dim):
    """model.py:474-479: per-axis rotary widths (frame, height, width) in real channels."""
    d = head_dim
    return d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)


def rope_table(grid, head_dim, seq_len=None, max_pos=1024):
    """Per-token rotation angles [seq_len, head_dim/2] (float64) for a (F, H', W') token grid.

    model.py:39-67: pair j of token n=(f,h,w) rotates by angle_f[f] | angle_h[h] | angle_w[w]
    (concatenated along the pair axis, 22|21|21 pairs for head_dim 128); tokens >= F*H'*W' are left
    unrotated (angle 0).
    """
    f, h, w = grid
    df, dh, dw = rope_axis_dims(head_dim)
    af, ah, aw = rope_angles(max_pos, df), rope_angles(max_pos, dh), rope_angles(max_pos, dw)
    tab = torch.cat([
        af[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        ah[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        aw[:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f * h * w, head_dim // 2)
    if seq_len is not None and seq_len > tab.shape[0]:
        tab = torch.cat([tab, tab.new_zeros(seq_len - tab.shape[0], tab.shape[1])])
    return tab


def rope_apply(x, angles):
    """x [L, n, d] (any float dtype), angles [L, d/2] float64 -> float32 [L, n, d].

    model.py:54-67: interleaved pairs (x[2j], x[2j+1]) treated as complex, multiplied in float64."""
    L, n, d = x.shape
    xd = x.to(torch.float64).reshape(L, n, d // 2, 2)
    c = torch.cos(angles).view(L, 1, d // 2)
    s = torch.sin(angles).view(L, 1, d // 2)
    out = torch.stack([xd[..., 0] * c - xd[..., 1] * s, xd[..., 0] * s + xd[..., 1] * c], dim=-1)
    return out.reshape(L, n, d).float()


# ------------------------------------------------------------------------------------------------
# norms / linear / attention                                     model.py:70-99, attention.py:24-130
# ------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps, rb):
    """model.py:70-86: (x.float() * rsqrt(mean(x^2)+eps)).type_as(x) * weight.  Under autocast x is the bf16
    Linear output, so the normalised value is rounded to bf16 before the fp32 weight multiply."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return rb(y) * weight.float()


def layer_norm(x, eps, weight=None, bias=None):
    """model.py:89-99: fp32 LayerNorm over the last dim (affine only for norm3)."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps)


def linear(x, w, b, rb):
    """nn.Linear; with rb = bf16_rt: autocast semantics (bf16 inputs/weights, fp32 accumulate, bf16 out)."""
    y = rb(x.float()) @ rb(w.float()).t()
    if b is not None:
        y = y + rb(b.float())      # autocast casts EVERY tensor argument of linear to bf16, the bias included
    return rb(y)


def attention(q, k, v, rb, scale=None):
    """attention.py:24-130 with one un-padded sequence: softmax(q k^T * scale) v; q [Lq,n,d], k,v [Lk,n,d].
    Inputs are cast to bf16 (rb), accumulation fp32, output bf16 (rb)."""
    d = q.shape[-1]
    scale = d ** -0.5 if scale is None else scale
    qh, kh, vh = (rb(t.float()).transpose(0, 1) for t in (q, k, v))  # [n, L, d]
    if qh.shape[1] * kh.shape[1] * qh.shape[0] <= (1 << 28):
        s = torch.matmul(qh, kh.transpose(1, 2)) * scale
        p = torch.softmax(s, dim=-1)
        return rb(torch.matmul(p, vh).transpose(0, 1).contiguous())
    # long sequences (full-size parity runs of this oracle on a GPU): same arithmetic, one head and <= 4096 query rows
    # at a time so that the score block stays small; every row's softmax is still exact over all keys
    out = torch.empty(qh.shape[0], qh.shape[1], vh.shape[2], dtype=torch.float32, device=q.device)
    for h in range(qh.shape[0]):
        kt = kh[h].t().contiguous()
        for r0 in range(0, qh.shape[1], 4096):
            p = torch.softmax(torch.matmul(qh[h, r0:r0 + 4096], kt) * scale, dim=-1)
            out[h, r0:r0 + 4096] = torch.matmul(p, vh[h])
    return rb(out.transpose(0, 1).contiguous())


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


# ------------------------------------------------------------------------------------------------
# block                                                                   model.py:274-313
# ------------------------------------------------------------------------------------------------
def self_attention(sd, pre, h, angles, num_heads, eps, rb, k_len=None):
    """model.py:127-156. h [L, C] fp32 (already normalised + modulated)."""
    L, C = h.shape
    d = C // num_heads
    q = rms_norm(linear(h, sd[pre + "q.weight"], sd[pre + "q.bias"], rb), sd[pre + "norm_q.weight"], eps, rb)
    k = rms_norm(linear(h, sd[pre + "k.weight"], sd[pre + "k.bias"], rb), sd[pre + "norm_k.weight"], eps, rb)
    v = linear(h, sd[pre + "v.weight"], sd[pre + "v.bias"], rb)
    q = rope_apply(q.view(L, num_heads, d), angles)
    k = rope_apply(k.view(L, num_heads, d), angles)
    v = v.view(L, num_heads, d)
    if k_len is not None:  # flash_attention(k_lens=seq_lens): keys restricted to the real tokens
        k, v = k[:k_len], v[:k_len]
    a = attention(q, k, v, rb).reshape(L, C)
    return linear(a, sd[pre + "o.weight"], sd[pre + "o.bias"], rb)


def cross_attention(sd, pre, h, ctx, num_heads, eps, rb):
    """model.py:159-181 (t2v): all context rows are attended (context_lens=None, model.py:548)."""
    L, C = h.shape
    d = C // num_heads
    q = rms_norm(linear(h, sd[pre + "q.weight"], sd[pre + "q.bias"], rb), sd[pre + "norm_q.weight"], eps, rb)
    k = rms_norm(linear(ctx, sd[pre + "k.weight"], sd[pre + "k.bias"], rb), sd[pre + "norm_k.weight"], eps, rb)
    v = linear(ctx, sd[pre + "v.weight"], sd[pre + "v.bias"], rb)
    a = attention(q.view(L, num_heads, d), k.view(-1, num_heads, d), v.view(-1, num_heads, d), rb).reshape(L, C)
    return linear(a, sd[pre + "o.weight"], sd[pre + "o.bias"], rb)


def block_forward(sd, pre, x, e0, angles, ctx, num_heads, eps, rb, k_len=None, x_is_bf16=False):
    """WanAttentionBlock.forward (model.py:274-313) for one sample.

    x [L, C] fp32 residual stream; e0 [6, C] fp32; ctx [Lc, C]; `x_is_bf16` reproduces block 0 where x
    is still the bf16 patch embedding so norm1(x).type_as(x) rounds to bf16 (SURVEY.md §8a notes)."""
    e = (sd[pre + "modulation"].float().view(6, -1) + e0.float())
    h = layer_norm(x, eps)
    if x_is_bf16:
        h = rb(h)
    h = h * (1 + e[1]) + e[0]
    y = self_attention(sd, pre + "self_attn.", h, angles, num_heads, eps, rb, k_len)
    x = x.float() + y * e[2]
    if (pre + "norm3.weight") in sd:
        h = layer_norm(x, eps, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    else:
        h = x
    x = x + cross_attention(sd, pre + "cross_attn.", h, ctx, num_heads, eps, rb)
    h = layer_norm(x, eps) * (1 + e[4]) + e[3]
    y = linear(rb(gelu_tanh(linear(h, sd[pre + "ffn.0.weight"], sd[pre + "ffn.0.bias"], rb))),
               sd[pre + "ffn.2.weight"], sd[pre + "ffn.2.bias"], rb)
    return x + y * e[5]


# ------------------------------------------------------------------------------------------------
# whole model                                                             model.py:486-609
# ------------------------------------------------------------------------------------------------
def patchify(u, patch):
    """[C,F,H,W] -> [L, C*pt*ph*pw] rows in (f,h,w) order, columns in patch_embedding.weight.flatten(1) order."""
    C, Fr, H, W = u.shape
    pt, ph, pw = patch
    x = u.view(C, Fr // pt, pt, H // ph, ph, W // pw, pw)
    x = x.permute(1, 3, 5, 0, 2, 4, 6).reshape((Fr // pt) * (H // ph) * (W // pw), C * pt * ph * pw)
    return x, (Fr // pt, H // ph, W // pw)


def unpatchify(y, grid, patch, out_dim):
    """model.py:581-609: [L, pt*ph*pw*c] -> [c, F*pt, H*ph, W*pw]."""
    f, h, w = grid
    u = y[: f * h * w].view(f, h, w, *patch, out_dim)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(out_dim, f * patch[0], h * patch[1], w * patch[2])


def model_forward(sd, cfg, x, t, context, seq_len, rb=ident):
    """WanModel.forward (model.py:486-579) for one sample (t2v).

    sd: reference-named state dict; cfg: dict(dim, num_heads, num_layers, freq_dim, text_len, patch_size, out_dim,
    eps); x [C_in,F,H,W] fp32; t: scalar tensor; context [Ltxt, text_dim]."""
    dim, nh, eps = cfg["dim"], cfg["num_heads"], cfg["eps"]
    patch = tuple(cfg["patch_size"])
    a, grid = patchify(x.float(), patch)
    w = sd["patch_embedding.weight"].float().flatten(1)
    tok = linear(a, w, sd["patch_embedding.bias"], rb)                     # model.py:529 (bf16 under autocast)
    L = tok.shape[0]
    assert L <= seq_len
    tok = torch.cat([tok, tok.new_zeros(seq_len - L, dim)])                # :535-538

    # time embedding, fp32 (autocast disabled)                             # :541-545
    sin = sinusoidal_embedding_1d(cfg["freq_dim"], t.reshape(1)).float()
    e = F.silu(sin @ sd["time_embedding.0.weight"].float().t() + sd["time_embedding.0.bias"].float())
    e = e @ sd["time_embedding.2.weight"].float().t() + sd["time_embedding.2.bias"].float()
    e0 = (F.silu(e) @ sd["time_projection.1.weight"].float().t() + sd["time_projection.1.bias"].float()).view(6, dim)

    # text embedding on the zero-padded context                            # :549-554
    ctx = torch.cat([context.float(), context.new_zeros(cfg["text_len"] - context.shape[0], context.shape[1]).float()])
    ctx = linear(rb(gelu_tanh(linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"], rb))),
                 sd["text_embedding.2.weight"], sd["text_embedding.2.bias"], rb)

    angles = rope_table(grid, dim // nh, seq_len)
    xs = tok
    for i in range(cfg["num_layers"]):
        xs = block_forward(sd, "blocks.%d." % i, xs, e0, angles, ctx, nh, eps, rb, k_len=L,
                           x_is_bf16=(i == 0 and rb is not ident))

    # head, fp32                                                            # :333-343
    em = sd["head.modulation"].float().view(2, dim) + e.view(1, dim)
    h = layer_norm(xs, eps) * (1 + em[1]) + em[0]
    y = h @ sd["head.head.weight"].float().t() + sd["head.head.bias"].float()
    return unpatchify(y, grid, patch, cfg["out_dim"])
