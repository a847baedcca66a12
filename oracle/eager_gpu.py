"""GPU eager baseline — TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product path).

A restatement of how the REFERENCE executes one WanAttentionBlock on a GPU: eager PyTorch, the
arithmetic `torch.autocast('cuda', bfloat16)` produces (bf16 cuBLAS Linears, fp32 islands where the reference
disables autocast), flash-attn-2 for the attention core exactly as wan/modules/attention.py:113-127 calls it, and
RoPE in complex128 as wan/modules/model.py:39-67 does.  /root/reference cannot travel to the GPU box, so the block is
restated here op for op (each step cites the reference line); bench.py times it as the `gpu_eager_baseline` record
("the number to beat on the same box"), it is never the thing shipped or the thing parity is claimed for.
"""
import math
import time

import torch
import torch.nn.functional as F


def _attention_fa2(q, k, v):
    """attention.py:52-130: one sequence, bf16, softmax_scale None (= d^-1/2), non-causal, window (-1, -1)."""
    L, n, d = q.shape
    try:
        import flash_attn
        lens = torch.tensor([0, L], dtype=torch.int32, device=q.device)
        lk = torch.tensor([0, k.shape[0]], dtype=torch.int32, device=q.device)
        o = flash_attn.flash_attn_varlen_func(q=q, k=k, v=v, cu_seqlens_q=lens, cu_seqlens_k=lk, max_seqlen_q=L,
                                              max_seqlen_k=k.shape[0], dropout_p=0.0, softmax_scale=None, causal=False,
                                              window_size=(-1, -1), deterministic=False)
        return o, "flash_attn %s (flash_attn_varlen_func)" % getattr(flash_attn, "__version__", "?")
    except Exception as ex:                                   # noqa: BLE001 — documented fallback of the BASELINE
        o = F.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])
        return o[0].transpose(0, 1).contiguous(), "torch SDPA (flash_attn unavailable: %s)" % repr(ex)[:80]


def _rope_apply(x, grid, freqs):
    """model.py:39-67 verbatim in behaviour: complex128 multiply, float32 result.  x [L, n, d]."""
    L, n, d = x.shape
    c = d // 2
    f, h, w = grid
    parts = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    seq = f * h * w
    xi = torch.view_as_complex(x[:seq].to(torch.float64).reshape(seq, n, -1, 2))
    fr = torch.cat([parts[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
                    parts[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                    parts[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, -1)
    xi = torch.view_as_real(xi * fr).flatten(2)
    return torch.cat([xi, x[seq:]]).float()


def _rms(x, w, eps):
    """model.py:70-86 under autocast: x is the bf16 Linear output."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * w


def block_forward(P, x, e0, grid, freqs, ctx, nh, eps):
    """model.py:274-313 for one sample.  P: dict of parameters (bf16 Linear weights/biases, fp32 norms/modulation);
    x [L, C] fp32; e0 [6, C] fp32; ctx [Lc, C] bf16.  Returns (x_out fp32, attention backend string)."""
    L, C = x.shape
    d = C // nh
    lin = lambda t, n: F.linear(t.to(torch.bfloat16), P[n + ".weight"], P[n + ".bias"])   # noqa: E731  (autocast Linear)
    e = (P["modulation"].view(6, C) + e0).chunk(6, dim=0)                                  # :292-295 fp32
    h = F.layer_norm(x.float(), (C,), None, None, eps) * (1 + e[1]) + e[0]                 # :298-299
    q = _rms(lin(h, "self_attn.q"), P["self_attn.norm_q.weight"], eps).view(L, nh, d)      # :139-141
    k = _rms(lin(h, "self_attn.k"), P["self_attn.norm_k.weight"], eps).view(L, nh, d)
    v = lin(h, "self_attn.v").view(L, nh, d)
    q, k = _rope_apply(q, grid, freqs), _rope_apply(k, grid, freqs)                        # :146-148 (fp32 out)
    a, backend = _attention_fa2(q.to(torch.bfloat16), k.to(torch.bfloat16), v)             # attention.py:59-83,113-127
    a = a.type(q.dtype)                                                                     # :130 back to fp32
    y = lin(a.flatten(1), "self_attn.o")                                                    # :154-155
    x = x + y * e[2]                                                                        # :302 fp32
    hc = F.layer_norm(x.float(), (C,), P["norm3.weight"], P["norm3.bias"], eps)            # :306
    Lc = ctx.shape[0]
    q = _rms(lin(hc, "cross_attn.q"), P["cross_attn.norm_q.weight"], eps).view(L, nh, d)   # :171-173
    k = _rms(lin(ctx, "cross_attn.k"), P["cross_attn.norm_k.weight"], eps).view(Lc, nh, d)
    v = lin(ctx, "cross_attn.v").view(Lc, nh, d)
    a, _ = _attention_fa2(q.to(torch.bfloat16), k.to(torch.bfloat16), v)                   # :176
    x = x + lin(a.type(q.dtype).flatten(1), "cross_attn.o")                                # :180, :306
    h = F.layer_norm(x.float(), (C,), None, None, eps) * (1 + e[4]) + e[3]                 # :307
    y = lin(F.gelu(lin(h, "ffn.0"), approximate="tanh"), "ffn.2")                          # :267-269
    return x + y * e[5], backend                                                            # :309


def make_block_params(dev, dim, ffn, seed=1):
    g = torch.Generator(device=dev).manual_seed(seed)
    P = {}

    def w(n, shape, std):
        P[n] = (torch.randn(*shape, generator=g, device=dev) * std)

    for a in ("self_attn", "cross_attn"):
        for n in ("q", "k", "v", "o"):
            w("%s.%s.weight" % (a, n), (dim, dim), 1 / math.sqrt(dim))
            w("%s.%s.bias" % (a, n), (dim,), 0.1)
        P["%s.norm_q.weight" % a] = 1 + 0.1 * torch.randn(dim, generator=g, device=dev)
        P["%s.norm_k.weight" % a] = 1 + 0.1 * torch.randn(dim, generator=g, device=dev)
    w("ffn.0.weight", (ffn, dim), 1 / math.sqrt(dim))
    w("ffn.0.bias", (ffn,), 0.1)
    w("ffn.2.weight", (dim, ffn), 1 / math.sqrt(ffn))
    w("ffn.2.bias", (dim,), 0.1)
    P["norm3.weight"] = 1 + 0.1 * torch.randn(dim, generator=g, device=dev)
    P["norm3.bias"] = 0.1 * torch.randn(dim, generator=g, device=dev)
    P["modulation"] = torch.randn(1, 6, dim, generator=g, device=dev) / math.sqrt(dim)
    for n in list(P):                      # autocast rounds Linear operands to bf16 at every call: store them rounded
        if n.endswith(".weight") and P[n].dim() == 2 or (n.endswith(".bias") and "norm" not in n):
            P[n] = P[n].to(torch.bfloat16)
    return P


def time_block(dev, seq_len, sample_tokens, dim, ffn, nh, text_len, layers, reps=2):
    """Times block_forward at `sample_tokens` tokens (CUDA events, 1 warm-up + reps) and extrapolates steps/s at seq_len:
    the attention core scales with L^2, everything else with L."""
    from wan.modules.model import rope_params          # the reference's own table construction (model.py:28-36)
    d = dim // nh
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1).to(dev)
    L = int(sample_tokens)
    side = max(1, int(round((L / 21) ** 0.5)))
    grid = (21, side, max(1, L // (21 * side))) if L >= 21 * 4 else (1, 1, L)
    L = grid[0] * grid[1] * grid[2]
    P = make_block_params(dev, dim, ffn)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(L, dim, generator=g, device=dev)
    e0 = 0.1 * torch.randn(6, dim, generator=g, device=dev)
    ctx = torch.randn(text_len, dim, generator=g, device=dev).to(torch.bfloat16)
    # attention core alone (the kernel BASELINE.md names: flash-attn-2 mma.sync kernels recompiled for sm_100)
    q = torch.randn(L, nh, d, generator=g, device=dev).to(torch.bfloat16)
    k = torch.randn(L, nh, d, generator=g, device=dev).to(torch.bfloat16)
    v = torch.randn(L, nh, d, generator=g, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        _, backend = _attention_fa2(q, k, v)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            _attention_fa2(q, k, v)
        a1.record()
        torch.cuda.synchronize()
        attn_ms = a0.elapsed_time(a1) / reps
        del q, k, v
        y, _ = block_forward(P, x, e0, grid, freqs, ctx, nh, 1e-6)      # warm-up
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        b0.record()
        for _ in range(reps):
            y, _ = block_forward(P, x, e0, grid, freqs, ctx, nh, 1e-6)
        b1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3 / reps
        block_ms = b0.elapsed_time(b1) / reps
    finite = bool(torch.isfinite(y).all().item())
    attn_fl = 4.0 * L * L * dim
    rest_ms = max(block_ms - attn_ms, 0.0)
    scale = seq_len / L
    block_at = attn_ms * scale * scale + rest_ms * scale
    step_s = 2 * layers * block_at * 1e-3
    return {"steps_per_sec_extrapolated": 1.0 / step_s, "unit": "steps/s", "kind": "port of the reference's eager GPU path "
            "(oracle/eager_gpu.py: bf16 cuBLAS Linears, %s, complex128 RoPE, eager elementwise)" % backend,
            "block_ms": round(block_ms, 2), "attention_core_ms": round(attn_ms, 2), "wall_ms": round(wall_ms, 2),
            "attention_tflops": round(attn_fl / (attn_ms * 1e-3) / 1e12, 1), "sample_tokens": L, "tokens": seq_len,
            "extrapolation": "2 forwards x %d blocks; attention core x (L/Ls)^2, rest x (L/Ls)" % layers,
            "finite": finite}
