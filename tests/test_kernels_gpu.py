"""-m gpu parity tests of the individual sm_100a kernels, called through the C ABI, against the CPU oracle
(oracle/dit_oracle.py with bf16 cast points).  Tolerances are stated per test."""
import math

import pytest
import torch

from oracle import dit_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (300, 256, 192), (64, 128, 64),
                                   (1000, 5120, 5120), (512, 1280, 4096), (4096, 768, 1536),
                                   # skinny problems take the 64-wide tile (8-stage ring): umT5 shapes, ragged N / M
                                   (32, 4096, 4096), (512, 10240, 4096), (100, 200, 64),
                                   # M > 512 stays on the 256-wide tile, ragged in every dimension
                                   (777, 1000, 320)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm(mv, M, N, K, epi):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g)
    y = O.bf16_rt(a.float() @ w.float().t() + bias)
    if epi == 0:
        ref = y
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi)
    elif epi == 1:
        ref = O.bf16_rt(O.gelu_tanh(y))
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi)
    elif epi == 2:
        x0 = torch.randn(M, N, generator=g)
        gate = torch.randn(N, generator=g)
        ref = x0 + y * gate
        out = x0.to(DEV)
        mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi, gate=gate.to(DEV))
    else:
        ref = y
        out = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
        mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi)
    torch.cuda.synchronize()
    # bf16 output rounding (2^-8 relative) is the only legitimate difference: fp32 accumulation order can flip
    # a rounding, so compare in rel-L2 (<= 3e-3) and bound the worst element by 2 bf16 ulps.
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) <= 3e-3
    err = (out.float().cpu() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * (1 if epi != 2 else 1 + gate.abs().max().item())
    assert (err <= tol).all(), err.max().item()


def test_gemm_no_bias_no_gate_strided(mv):
    g = torch.Generator().manual_seed(5)
    M, N, K = 200, 384, 256
    big_a = torch.randn(M, K + 64, generator=g).bfloat16().to(DEV)
    a = big_a[:, 32:32 + K]  # row stride K+64, offset 32 elements (64 B aligned)
    w = (torch.randn(N, K, generator=g) / 16).bfloat16().to(DEV)
    big_out = torch.zeros(M, N + 128, dtype=torch.float32, device=DEV)
    out = big_out[:, 64:64 + N]
    x0 = torch.randn(M, N, generator=g).to(DEV)
    out.copy_(x0)
    mv.gemm(a, w, None, out, 2)
    ref = x0.cpu() + O.bf16_rt(a.float().cpu() @ w.float().cpu().t())
    assert rel_l2(out, ref) <= 3e-3
    assert (big_out[:, :64] == 0).all() and (big_out[:, 64 + N:] == 0).all()


# ------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("Lq,Lk,H", [(256, 128, 1), (256, 256, 2), (300, 333, 2), (128, 512, 3), (1000, 512, 2),
                                     (2048, 2048, 4), (4100, 4100, 1)])
def test_attention(mv, Lq, Lk, H):
    g = torch.Generator().manual_seed(Lq + 3 * Lk + H)
    q = torch.randn(Lq, H, 128, generator=g).bfloat16()
    k = torch.randn(Lk, H, 128, generator=g).bfloat16()
    v = torch.randn(Lk, H, 128, generator=g).bfloat16()
    ref = O.attention(q, k, v, O.bf16_rt)
    out = torch.full((Lq, H, 128), float("nan"), dtype=torch.bfloat16, device=DEV)
    mv.attention(q.to(DEV), k.to(DEV), v.to(DEV), out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    # P is rounded to bf16 before P.V (as in flash-attn): rel-L2 <= 3e-3 of the fp32-softmax oracle... measured
    # against bf16 outputs, so allow 5e-3; max-abs bounded by 2e-2 for unit-variance V.
    assert rel_l2(out.float(), ref) <= 5e-3
    assert (out.float().cpu() - ref).abs().max().item() <= 2e-2


def test_attention_nonfinite_input_terminates(mv):
    """A NaN key must propagate into the outputs of its head (as in the reference softmax) and nowhere else — and the
    overflow-guarded redo of the fixed-reference softmax must not spin on it (it is bounded to one exact redo)."""
    g = torch.Generator().manual_seed(5)
    Lq, Lk, H = 256, 300, 2
    q = torch.randn(Lq, H, 128, generator=g).bfloat16()
    k = torch.randn(Lk, H, 128, generator=g).bfloat16()
    v = torch.randn(Lk, H, 128, generator=g).bfloat16()
    k[200, 0, 7] = float("nan")
    ref = O.attention(q[:, 1:], k[:, 1:], v[:, 1:], O.bf16_rt)
    out = torch.zeros(Lq, H, 128, dtype=torch.bfloat16, device=DEV)
    mv.attention(q.to(DEV), k.to(DEV), v.to(DEV), out)
    torch.cuda.synchronize()
    assert torch.isnan(out[:, 0].float()).all()
    assert rel_l2(out[:, 1:].float(), ref) <= 5e-3


@pytest.mark.parametrize("Lq,Lk,H,n_dst", [(301, 333, 2, 2), (1024, 512, 3, 4)])
def test_attention_scatter_epilogue(mv, Lq, Lk, H, n_dst):
    """mv_attention_fwd_scatter (the fused Ulysses return path): query row r is stored into destination
    r // rows_per_rank at [src_rank][r % rows_per_rank][H*128].  The destinations are local buffers here (on a
    multi-GPU box they are the peers' IPC-mapped slabs, tests/test_sp_multigpu.py)."""
    g = torch.Generator().manual_seed(Lq + Lk + H)
    q = torch.randn(Lq, H, 128, generator=g).bfloat16()
    k = torch.randn(Lk, H, 128, generator=g).bfloat16()
    v = torch.randn(Lk, H, 128, generator=g).bfloat16()
    ref = O.attention(q, k, v, O.bf16_rt)
    rows = (Lq + n_dst - 1) // n_dst
    src_rank = n_dst - 1
    dsts = [torch.full((n_dst, rows, H * 128), float("nan"), dtype=torch.bfloat16, device=DEV) for _ in range(n_dst)]
    tab = mv.ptr_table([d.data_ptr() for d in dsts])
    mv.attention_scatter(q.to(DEV), k.to(DEV), v.to(DEV), tab, n_dst, src_rank, rows, H * 128)
    torch.cuda.synchronize()
    got = torch.cat([d[src_rank] for d in dsts], 0)[:Lq].view(Lq, H, 128).float().cpu()
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) <= 5e-3
    for d in dsts:   # nothing but the src_rank slab (and only its valid rows) was written
        other = torch.cat([d[:src_rank], d[src_rank + 1:]], 0)
        assert torch.isnan(other.float()).all()
    tail = dsts[-1][src_rank][Lq - (n_dst - 1) * rows:]
    assert torch.isnan(tail.float()).all()


def test_attention_strided_views_and_scale(mv):
    """q/k/v as column slices of one fused [L, 3*H*128] QKV buffer (how the DiT block calls it), sharp softmax."""
    g = torch.Generator().manual_seed(11)
    L, H = 640, 2
    qkv = (torch.randn(L, 3 * H * 128, generator=g) * 2.0).bfloat16().to(DEV)
    q = qkv[:, : H * 128].view(L, H, 128)[:, :, :]
    q = qkv.as_strided((L, H, 128), (3 * H * 128, 128, 1), 0)
    k = qkv.as_strided((L, H, 128), (3 * H * 128, 128, 1), H * 128)
    v = qkv.as_strided((L, H, 128), (3 * H * 128, 128, 1), 2 * H * 128)
    out = torch.empty(L, H, 128, dtype=torch.bfloat16, device=DEV)
    mv.attention(q, k, v, out, softmax_scale=0.5)
    ref = O.attention(q.cpu(), k.cpu(), v.cpu(), O.bf16_rt, scale=0.5)
    assert rel_l2(out.float(), ref) <= 5e-3


@pytest.mark.parametrize("jumps", [(0.0, 0.3, 0.6, 0.9, 1.2, 1.5, 1.8, 2.1), (0.0, 0.0, 5.0, 5.0, 0.5, 12.0, 12.0, 1.0),
                                   (3.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)])
@pytest.mark.parametrize("offset", [5, 70])
def test_attention_running_max_growth(mv, jumps, offset):
    """Online-softmax bookkeeping under a row max that keeps growing along the key axis: per 128-key block b every key
    gets an extra score of 16.3 * jumps[b] in log2 units (q = ones, k += jumps[b] * ones).  Small steps exercise the
    lazy (deferred) rescale of the accumulator, jumps of more than 2^64 the exact redo path of the stale-reference
    kernel, a decreasing profile the no-rescale path.  Exact fp32 softmax reference."""
    g = torch.Generator().manual_seed(19)
    Lq, H = 256, 2
    Lk = 128 * len(jumps) - 37                   # ragged last block
    q = torch.ones(Lq, H, 128) + 0.05 * torch.randn(Lq, H, 128, generator=g)
    k = 0.2 * torch.randn(Lk, H, 128, generator=g)
    for b, c in enumerate(jumps):
        k[128 * b + offset:128 * b + offset + 4] += c        # a few keys per block carry the jump (first / second half)
    v = torch.randn(Lk, H, 128, generator=g)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    ref = O.attention(q, k, v, O.bf16_rt)
    out = torch.full((Lq, H, 128), float("nan"), dtype=torch.bfloat16, device=DEV)
    mv.attention(q.to(DEV), k.to(DEV), v.to(DEV), out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) <= 5e-3


@pytest.mark.parametrize("M,N,K", [(1024, 512, 64), (1300, 768, 192), (2048, 1000, 320), (4096, 5120, 1024),
                                   (1025, 520, 72)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_cta_pair(mv, M, N, K, epi):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tile over two SMs, each CTA staging half of the W tile)
    against the oracle: full tiles, ragged M (a pair whose second CTA is partly / entirely out of range), ragged N."""
    g = torch.Generator().manual_seed(M + N * 3 + K + epi)
    a = (torch.randn(M, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g)
    y = O.bf16_rt(a.float() @ w.float().t() + bias)
    mv.gemm_config(1)
    try:
        if epi in (0, 1):
            ref = y if epi == 0 else O.bf16_rt(O.gelu_tanh(y))
            out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi)
            gate = None
        elif epi == 2:
            x0 = torch.randn(M, N, generator=g)
            gate = torch.randn(N, generator=g)
            ref = x0 + y * gate
            out = x0.to(DEV)
            mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi, gate=gate.to(DEV))
        else:
            ref = y
            gate = None
            out = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
            mv.gemm(a.to(DEV), w.to(DEV), bias.to(DEV), out, epi)
        torch.cuda.synchronize()
    finally:
        mv.gemm_config(-2)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) <= 3e-3
    err = (out.float().cpu() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * (1 if epi != 2 else 1 + gate.abs().max().item())
    assert (err <= tol).all(), err.max().item()


def test_gemm_cta_pair_ksplit_equals_single_cta(mv):
    """K-split A operand (the Ulysses return layout) through the pair kernel == the single-CTA kernel, bit for bit
    (same K order, same fp32 accumulation)."""
    g = torch.Generator().manual_seed(11)
    P, M, Kb, N = 4, 1500, 128, 768
    a = torch.randn(P, M, Kb, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, P * Kb, generator=g) / math.sqrt(P * Kb)).bfloat16().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    outs = []
    for pair in (0, 1):
        mv.gemm_config(pair)
        try:
            o = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            mv.gemm_ksplit(a, w, bias, o, 0)
            torch.cuda.synchronize()
        finally:
            mv.gemm_config(-2)
        outs.append(o)
    ref = O.bf16_rt(a.permute(1, 0, 2).reshape(M, P * Kb).float().cpu() @ w.float().cpu().t() + bias.cpu())
    assert rel_l2(outs[1].float(), ref) <= 3e-3
    assert torch.equal(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------- rowops
@pytest.mark.parametrize("M,C", [(7, 128), (33, 5120), (5, 1536)])
@pytest.mark.parametrize("variant", ["plain", "mod", "affine", "round_mod"])
def test_ln_modulate(mv, M, C, variant):
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C, generator=g) * 3 + 0.5
    shift = scale = w = b = None
    if "mod" in variant:
        shift, scale = torch.randn(C, generator=g), torch.randn(C, generator=g) * 0.3
    if variant == "affine":
        w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    h = O.layer_norm(x, 1e-6, w, b)
    if variant == "round_mod":
        h = O.bf16_rt(h)
    if shift is not None:
        h = h * (1 + scale) + shift
    ref = O.bf16_rt(h)
    out = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    mv.ln_modulate(x.to(DEV), out, None if shift is None else shift.to(DEV), None if scale is None else scale.to(DEV),
                   None if w is None else w.to(DEV), None if b is None else b.to(DEV), 1e-6,
                   variant == "round_mod")
    assert rel_l2(out.float(), ref) <= 3e-3
    assert (out.float().cpu() - ref).abs().max() <= 2.0 ** -6 * ref.abs().max()


@pytest.mark.parametrize("M,C,hd,rope", [(9, 256, 128, True), (40, 5120, 128, True), (17, 5120, 128, False),
                                         (6, 128, 32, True)])
def test_rmsnorm_rope(mv, M, C, hd, rope):
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C + 64, generator=g) * 2).bfloat16()
    wt = torch.randn(C, generator=g)
    grid = (1, 1, M)
    ang = O.rope_table((M, 1, 1), hd, M + 2)[:M] if rope else None
    y = O.rms_norm(x[:, 8:8 + C], wt, 1e-6, O.bf16_rt)
    if rope:
        y = O.rope_apply(y.view(M, C // hd, hd), ang).reshape(M, C)
    ref = O.bf16_rt(y)
    xd = x.to(DEV)
    cs = None
    if rope:
        cs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(DEV)
    mv.rmsnorm_rope(xd[:, 8:8 + C], wt.to(DEV), cs, hd, 1e-6)
    assert rel_l2(xd[:, 8:8 + C].float(), ref) <= 3e-3
    assert torch.equal(xd[:, :8].cpu(), x[:, :8]) and torch.equal(xd[:, 8 + C:].cpu(), x[:, 8 + C:])


@pytest.mark.parametrize("M,C,hd,P", [(9, 256, 128, 0), (40, 5120, 128, 0), (33, 5120, 128, 8), (21, 1024, 128, 4),
                                      (6, 128, 32, 2), (5, 7680, 128, 0)])
def test_qkv_norm_rope_fused(mv, M, C, hd, P):
    """q and k RMSNorm + RoPE (and, with P > 0, the Ulysses head scatter of q, k, v) of a fused QKV row in ONE launch ==
    the per-tensor oracle; V is copied bit-exactly; padding columns of the buffer are untouched."""
    g = torch.Generator().manual_seed(M + C + P)
    ld = 3 * C + 16
    x = (torch.randn(M, ld, generator=g) * 2).bfloat16()
    gq, gk = 1 + 0.2 * torch.randn(C, generator=g), 1 + 0.2 * torch.randn(C, generator=g)
    ang = O.rope_table((M, 1, 1), hd, M)
    ref = {}
    for name, c0, wt in (("q", 0, gq), ("k", C, gk)):
        y = O.rms_norm(x[:, c0:c0 + C], wt, 1e-6, O.bf16_rt)
        ref[name] = O.bf16_rt(O.rope_apply(y.view(M, C // hd, hd), ang).reshape(M, C))
    ref["v"] = x[:, 2 * C:3 * C].float()
    cs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(DEV)
    xd = x.to(DEV)
    qkv = xd[:, :3 * C]
    if P == 0:
        mv.qkv_norm_rope(qkv, gq.to(DEV), gk.to(DEV), cs, hd, 1e-6)
        assert rel_l2(xd[:, 0:C].float(), ref["q"]) <= 3e-3 and rel_l2(xd[:, C:2 * C].float(), ref["k"]) <= 3e-3
        assert torch.equal(xd[:, 2 * C:].cpu(), x[:, 2 * C:])              # v and the padding are untouched
    else:
        slot = 1                                                        # write into slab `slot` of [P slots, M, C/P] buffers
        bufs = {n: torch.full((P, P, M, C // P), float("nan"), dtype=torch.bfloat16, device=DEV) for n in "qkv"}
        tabs = tuple(mv.ptr_table([bufs[n][d].data_ptr() for d in range(P)]) for n in "qkv")
        mv.qkv_norm_rope(qkv, gq.to(DEV), gk.to(DEV), cs, hd, 1e-6, dst=tabs, n_dst=P, src_slot=slot)
        assert torch.equal(xd.cpu(), x)                                  # the source rows are not modified
        for n in "qkv":
            got = bufs[n][:, slot].permute(1, 0, 2).reshape(M, C).float()   # [dst][row][C/P] -> [row][C]
            if n == "v":
                assert torch.equal(got.cpu(), ref["v"])
            else:
                assert rel_l2(got, ref[n]) <= 3e-3, n
            other = torch.cat([bufs[n][:, :slot], bufs[n][:, slot + 1:]], dim=1)
            assert torch.isnan(other.float()).all()                          # nothing outside the addressed slab


def test_qkv_prepare_p2p_per_slab_equals_fused_launch(mv):
    """The pipelined Ulysses exchange scatters q, k and v with one mv_qkv_prepare_p2p launch per column slab (so that each
    can run next to the following slab's GEMM); the three launches must write the same bits as the single fused
    mv_qkv_norm_rope launch."""
    M, C, hd, P, slot = 333, 1024, 128, 4, 2
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(M, 3 * C, generator=g) * 2).bfloat16().to(DEV)
    gq, gk = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV), (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    ang = O.rope_table((M, 1, 1), hd, M)
    cs = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(DEV)
    mk = lambda: {n: torch.full((P, P, M, C // P), float("nan"), dtype=torch.bfloat16, device=DEV) for n in "qkv"}  # noqa: E731
    fused, slabs = mk(), mk()
    tabs = {n: mv.ptr_table([fused[n][d].data_ptr() for d in range(P)]) for n in "qkv"}
    mv.qkv_norm_rope(x, gq, gk, cs, hd, 1e-6, dst=(tabs["q"], tabs["k"], tabs["v"]), n_dst=P, src_slot=slot)
    for i, (n, gain, c) in enumerate((("q", gq, cs), ("k", gk, cs), ("v", None, None))):
        tab = mv.ptr_table([slabs[n][d].data_ptr() for d in range(P)])
        mv.qkv_prepare_p2p(x[:, i * C:(i + 1) * C], gain, c, tab, slot, P, hd, 1e-6)
    torch.cuda.synchronize()
    for n in "qkv":
        a, b = fused[n].view(torch.int16), slabs[n].view(torch.int16)
        assert torch.equal(a, b), n


def test_modulation_table(mv):
    g = torch.Generator().manual_seed(4)
    mods, e0 = torch.randn(5, 6, 384, generator=g), torch.randn(6 * 384, generator=g)
    out = torch.empty(5, 6, 384, device=DEV)
    mv.modulation_table(mods.to(DEV), e0.to(DEV), out)
    assert torch.equal(out.cpu(), mods + e0.view(1, 6, 384))


def test_patchify(mv):
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(16, 3, 8, 12, generator=g)
    ref, grid = O.patchify(lat, (1, 2, 2))
    out = torch.empty(ref.shape, dtype=torch.bfloat16, device=DEV)
    mv.patchify(lat.to(DEV), out)
    assert torch.equal(out.float().cpu(), O.bf16_rt(ref))


@pytest.mark.parametrize("grid,C", [((2, 3, 5), 256), ((3, 8, 8), 5120)])
def test_head_unpatchify(mv, grid, C):
    g = torch.Generator().manual_seed(C)
    L = grid[0] * grid[1] * grid[2]
    x = torch.randn(L + 5, C, generator=g) * 2
    shift, scale = torch.randn(C, generator=g), torch.randn(C, generator=g) * 0.2
    w, b = torch.randn(64, C, generator=g) / math.sqrt(C), torch.randn(64, generator=g)
    h = O.layer_norm(x, 1e-6) * (1 + scale) + shift
    ref = O.unpatchify(h.double() @ w.double().t() + b.double(), grid, (1, 2, 2), 16).float()
    out = torch.empty(ref.shape, dtype=torch.float32, device=DEV)
    mv.head_unpatchify(x.to(DEV), shift.to(DEV), scale.to(DEV), w.to(DEV), b.to(DEV), out, grid)
    # fp32 end to end: 1e-5 relative
    assert rel_l2(out, ref) <= 1e-5


def test_time_embedding_path(mv):
    g = torch.Generator().manual_seed(9)
    dim, fd = 512, 256
    t = torch.tensor([937], dtype=torch.int64)
    sin_ref = O.sinusoidal_embedding_1d(fd, t).float()[0]
    sin = torch.empty(fd, dtype=torch.float32, device=DEV)
    mv.sinusoid_embed(t.to(DEV), sin)
    assert (sin.cpu() - sin_ref).abs().max() <= 1e-6
    w0, b0 = torch.randn(dim, fd, generator=g) * 0.05, torch.randn(dim, generator=g)
    w1, b1 = torch.randn(dim * 6, dim, generator=g) * 0.05, torch.randn(dim * 6, generator=g)
    e = torch.empty(dim, dtype=torch.float32, device=DEV)
    e0 = torch.empty(dim * 6, dtype=torch.float32, device=DEV)
    mv.linear_f32_vec(sin, w0.to(DEV), b0.to(DEV), e, act_in=0)
    mv.linear_f32_vec(e, w1.to(DEV), b1.to(DEV), e0, act_in=1)
    e_ref = sin_ref.double() @ w0.double().t() + b0.double()
    e0_ref = torch.nn.functional.silu(e_ref) @ w1.double().t() + b1.double()
    assert rel_l2(e, e_ref) <= 1e-5 and rel_l2(e0, e0_ref) <= 1e-5


def test_errors_are_reported(mv):
    a = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)   # K=12 not a multiple of 8
    w = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)
    out = torch.zeros(8, 8, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        mv.gemm(a, w, None, out, 0)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        mv.gemm(a.cpu(), w, None, out, 0)
