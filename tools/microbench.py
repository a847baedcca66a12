"""Kernel micro-benchmarks on one B200 (CUDA events, L2 flushed between iterations).
Prints one JSON line per case; used to fill profiles/ and DESIGN.md.  Not the headline bench (bench.py)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402

DEV = "cuda"
FLUSH = None


def timeit(fn, iters=5, warmup=2):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm(M, N, K, epi=0):
    """Both tilings of mv_gemm_bf16 (single CTA 128x256, CTA pair 256x256) and cuBLAS (F.linear, bf16 out — it does
    LESS epilogue work than epi 1 / 2) on the same operands."""
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV)
    f32 = epi in (2, 3)
    out = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=DEV)
    gate = torch.randn(N, device=DEV) if epi == 2 else None
    res = {}
    for pair in (0, 1):
        mv.gemm_config(pair)
        res[pair] = timeit(lambda: mv.gemm(a, w, bias, out, epi, gate=gate))
    mv.gemm_config(-2)
    ms_ref = timeit(lambda: torch.nn.functional.linear(a, w, bias.bfloat16()))
    fl = 2.0 * M * N * K
    print(json.dumps(dict(kind="gemm", M=M, N=N, K=K, epi=epi, ms=round(res[0], 4), tflops=round(fl / res[0] / 1e9, 1),
                          pair_ms=round(res[1], 4), pair_tflops=round(fl / res[1] / 1e9, 1),
                          cublas_ms=round(ms_ref, 4), cublas_tflops=round(fl / ms_ref / 1e9, 1),
                          best_vs_cublas=round(ms_ref / min(res.values()), 3))), flush=True)


def bench_attn(L, H, Lk=None):
    Lk = Lk or L
    q = torch.randn(L, H, 128, device=DEV).bfloat16()
    k = torch.randn(Lk, H, 128, device=DEV).bfloat16()
    v = torch.randn(Lk, H, 128, device=DEV).bfloat16()
    o = torch.empty_like(q)
    ms = timeit(lambda: mv.attention(q, k, v, o), iters=3, warmup=1)
    fl = 4.0 * L * Lk * H * 128
    rec = dict(kind="attn", Lq=L, Lk=Lk, H=H, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))
    try:
        from flash_attn import flash_attn_func
        ms_fa = timeit(lambda: flash_attn_func(q[None], k[None], v[None]), iters=3, warmup=1)
        rec.update(fa2_ms=round(ms_fa, 4), fa2_tflops=round(fl / ms_fa / 1e9, 1))
        ref = flash_attn_func(q[None], k[None], v[None])[0]
        rec["max_abs_vs_fa2"] = round((ref.float() - o.float()).abs().max().item(), 5)
    except Exception as ex:  # flash-attn absent or not runnable on this box
        rec["fa2_error"] = repr(ex)[:120]
    print(json.dumps(rec), flush=True)


def bench_rowops(M=75600, C=5120):
    x = torch.randn(M, C, device=DEV)
    out = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    sh, sc = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    ms = timeit(lambda: mv.ln_modulate(x, out, sh, sc))
    gb = M * C * 6 / 1e9
    print(json.dumps(dict(kind="ln_modulate", M=M, C=C, ms=round(ms, 4), gbs=round(gb / ms * 1e3, 1))), flush=True)
    qk = torch.randn(M, C, device=DEV).bfloat16()
    wt = torch.randn(C, device=DEV)
    cs = torch.randn(M, 64, 2, device=DEV)
    ms = timeit(lambda: mv.rmsnorm_rope(qk, wt, cs))
    gb = (M * C * 4 + M * 512) / 1e9
    print(json.dumps(dict(kind="rmsnorm_rope", M=M, C=C, ms=round(ms, 4), gbs=round(gb / ms * 1e3, 1))), flush=True)
    qkv = torch.randn(M, 3 * C, device=DEV).bfloat16()
    wk = torch.randn(C, device=DEV)
    ms = timeit(lambda: mv.qkv_norm_rope(qkv, wt, wk, cs))
    gb = (2 * M * C * 4 + M * 512) / 1e9           # q and k read + written once (bf16), one cos/sin row per token
    print(json.dumps(dict(kind="qkv_norm_rope (q+k, one launch)", M=M, C=C, ms=round(ms, 4),
                          gbs=round(gb / ms * 1e3, 1))), flush=True)
    P = 8
    bufs = [torch.empty(P, M, C // P, dtype=torch.bfloat16, device=DEV) for _ in range(3)]
    tabs = tuple(mv.ptr_table([b[d].data_ptr() for d in range(P)]) for b in bufs)
    ms = timeit(lambda: mv.qkv_norm_rope(qkv, wt, wk, cs, dst=tabs, n_dst=P, src_slot=0))
    gb = (3 * M * C * 4 + M * 512) / 1e9
    print(json.dumps(dict(kind="qkv_norm_rope (q+k+v scatter, local)", M=M, C=C, ms=round(ms, 4),
                          gbs=round(gb / ms * 1e3, 1))), flush=True)
    lat = [torch.randn(16 * 21 * 90 * 160, device=DEV) for _ in range(9)]
    coef = [5.0, 0.9, 1.0, 2.0, 0.9, 0.1, 0.2, 0.5, 1.1, 0.0, 0.5, 0.0, 2.0, 0.9, 0.1, 0.2, 1.1, 0.0, 0.5, 0.0]
    ms = timeit(lambda: mv.unipc_cfg_step(lat[0], lat[1], lat[2], lat[3], [lat[4], lat[5], None], coef, lat[6], lat[7],
                                          lat[8]))
    gb = lat[0].numel() * 4 * 9 / 1e9
    print(json.dumps(dict(kind="unipc_cfg_step (720P latent)", ms=round(ms, 4), gbs=round(gb / ms * 1e3, 1))), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "rowops"]
    mv.device_check()
    if "gemm" in which:
        for (M, N, K, epi) in [(8192, 8192, 8192, 0), (75600, 15360, 5120, 0), (75600, 5120, 5120, 2),
                               (75600, 13824, 5120, 1), (75600, 5120, 13824, 2), (16380, 15360, 5120, 0),
                               (512, 5120, 4096, 1)]:
            bench_gemm(M, N, K, epi)
    if "attn" in which:
        for (L, H, Lk) in [(4096, 40, None), (16384, 40, None), (32768, 10, None), (75600, 5, None),
                           (75600, 40, 512)]:
            bench_attn(L, H, Lk)
    if "rowops" in which:
        bench_rowops()
    if "attn_one" in which:      # single launch for an ncu --set full capture
        bench_attn(75600, 5, None)
    if "attn_720p" in which:     # the bench.py self-attention shape (one launch; for the ncu traffic capture)
        bench_attn(75600, 40, None)
    for w in which:              # attn_shape:Lq:Lk:H -> one launch shape (ncu traffic captures for profiles/attn_traffic.json)
        if w.startswith("attn_shape:"):
            _, lq, lk, h = w.split(":")
            bench_attn(int(lq), int(h), int(lk))
        if w.startswith("gemm_shape:"):
            _, m, n, k, epi = w.split(":")
            bench_gemm(int(m), int(n), int(k), int(epi))
    if "gemm_one" in which:
        bench_gemm(75600, 5120, 5120, 2)
