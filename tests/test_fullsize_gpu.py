"""-m gpu: BASELINE.json full sizes (720P: 75 600 tokens, 14B width), checked through size-independent properties
where a CPU oracle run would take hours: row-stochasticity and key-permutation invariance of attention, GEMM against
an independent tensor-core implementation, run-to-run determinism of a 14B-width block."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
L720 = 75600


def test_attention_full_length_row_stochastic_and_permutation_invariant(mv):
    g = torch.Generator(device=DEV).manual_seed(0)
    H = 2
    q = torch.randn(L720, H, 128, device=DEV, generator=g).bfloat16()
    k = torch.randn(L720, H, 128, device=DEV, generator=g).bfloat16()
    v = torch.randn(L720, H, 128, device=DEV, generator=g).bfloat16()
    # (1) softmax rows sum to one: with V = 1 the output is exactly 1 (up to bf16 rounding of P and of the output)
    ones = torch.ones_like(v)
    o1 = torch.empty_like(q)
    mv.attention(q, k, ones, o1)
    assert (o1.float() - 1).abs().max().item() <= 8e-3
    # (2) attention is invariant under a permutation of the keys/values (online softmax, tile order, masking of the
    # ragged last tile: 75 600 = 1181 * 64 + 16)
    o = torch.empty_like(q)
    mv.attention(q, k, v, o)
    perm = torch.randperm(L720, device=DEV, generator=g)
    op = torch.empty_like(q)
    mv.attention(q, k[perm].contiguous(), v[perm].contiguous(), op)
    assert torch.isfinite(o.float()).all()
    # outputs are averages of ~75k unit-variance values: |o| ~ 4e-3; compare with an absolute bound a few bf16 ulps wide
    assert (o.float() - op.float()).abs().max().item() <= 2e-3
    # (3) against an fp32 reference on a slice of the query rows (incl. the last, ragged, 256-row CTA)
    rows = torch.cat([torch.arange(0, 128), torch.arange(L720 - 200, L720)]).to(DEV)
    ref = torch.softmax((q[rows].float().transpose(0, 1) @ k.float().permute(1, 2, 0)) / math.sqrt(128), dim=-1) @ \
        v.float().transpose(0, 1)
    err = (o[rows].float() - ref.transpose(0, 1)).abs().max().item()
    assert err <= 1e-3, err


@pytest.mark.parametrize("N,K,epi", [(15360, 5120, 0), (5120, 13824, 2), (13824, 5120, 1)])
def test_gemm_full_size_against_cublas(mv, N, K, epi):
    g = torch.Generator(device=DEV).manual_seed(1)
    M = L720
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    ref = (a @ w.t()).float() + bias          # cuBLAS bf16 GEMM, fp32 accumulate, bf16 output rounding
    if epi == 0:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        mv.gemm(a, w, bias, out, 0)
        got = out.float()
        ref = ((a @ w.t()).float() + bias).bfloat16().float()
    elif epi == 1:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        mv.gemm(a, w, bias, out, 1)
        got = out.float()
        ref = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
    else:
        x0 = torch.randn(M, N, device=DEV, generator=g)
        gate = torch.randn(N, device=DEV, generator=g)
        out = x0.clone()
        mv.gemm(a, w, bias, out, 2, gate=gate)
        got = out
        ref = x0 + ref.bfloat16().float() * gate
    rel = ((got - ref).double().norm() / ref.double().norm()).item()
    assert rel <= 3e-3, rel


def test_block_14b_width_is_deterministic():
    """Two runs of a 14B-width block on 4 096 tokens give bit-identical results (no atomics, fixed tile order)."""
    from oracle.fill import fill_parameters
    from wan.modules.model import WanAttentionBlock, rope_params
    dim, ffn, nh, L = 5120, 13824, 40, 4096
    blk = WanAttentionBlock("t2v_cross_attn", dim, ffn, nh, (-1, -1), True, True, 1e-6).eval().requires_grad_(False)
    fill_parameters(blk, 5)
    blk.to(DEV)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, L, dim, generator=g).to(DEV)
    e = (torch.randn(1, 6, dim, generator=g) * 0.5).to(DEV)
    ctx = torch.randn(1, 512, dim, generator=g).to(DEV)
    d = dim // nh
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1)
    args = (x, e, torch.tensor([L]), torch.tensor([[4, 32, 32]]), freqs, ctx, None)
    y1 = blk(*args)
    y2 = blk(*args)
    assert torch.isfinite(y1).all() and torch.equal(y1, y2)


def test_block_14b_full_length_matches_oracle_on_gpu():
    """One WanAttentionBlock at the 14B width over the FULL 720P sequence (75 600 tokens, grid 21 x 45 x 80) against
    the oracle (oracle/dit_oracle.py, plain torch) executed on the GPU in fp32 (TF32 off, attention chunked by query
    rows): (a) the fp32 reference semantics, (b) the bf16 cast-point emulation.  SURVEY.md §8c block tolerance:
    rel-L2 <= 1e-2 vs fp32 and no worse than 1.5x the emulation."""
    from oracle import dit_oracle as O
    from oracle.fill import fill_parameters
    from wan.modules.model import WanAttentionBlock, rope_params
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dim, ffn, nh, L = 5120, 13824, 40, L720
    grid = (21, 45, 80)
    blk = WanAttentionBlock("t2v_cross_attn", dim, ffn, nh, (-1, -1), True, True, 1e-6).eval().requires_grad_(False)
    fill_parameters(blk, 11)
    sd = {"b." + k: v.clone().to(DEV) for k, v in blk.state_dict().items()}
    blk.to(DEV)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(1, L, dim, generator=g).to(DEV)
    e = (torch.randn(1, 6, dim, generator=g) * 0.5).to(DEV)
    ctx = torch.randn(1, 512, dim, generator=g).to(DEV)
    d = dim // nh
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1)
    y = blk(x, e, torch.tensor([L]), torch.tensor([grid]), freqs, ctx, None)[0]
    torch.cuda.synchronize()
    ang = O.rope_table(grid, d, L).to(DEV)
    with torch.no_grad():
        ref32 = O.block_forward(sd, "b.", x[0], e[0], ang, ctx[0], nh, 1e-6, O.ident, k_len=L)
        emu = O.block_forward(sd, "b.", x[0], e[0], ang, ctx[0], nh, 1e-6, O.bf16_rt, k_len=L)

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()
    err_ours, err_emu = rel(y, ref32), rel(emu, ref32)
    assert torch.isfinite(y).all()
    assert err_ours <= 1e-2 and err_ours <= 1.5 * err_emu + 1e-3, (err_ours, err_emu)
    assert rel(y, emu) <= 5e-3
    # the update of the residual stream (what the block adds) is the sensitive quantity: x itself dominates the norm
    upd, upd_ref, upd_emu = y - x[0], ref32 - x[0], emu - x[0]
    assert rel(upd, upd_ref) <= 1.5 * rel(upd_emu, upd_ref) + 2e-3, (rel(upd, upd_ref), rel(upd_emu, upd_ref))
