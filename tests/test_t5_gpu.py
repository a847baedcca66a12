"""GPU: the umT5 text-encoder path through the C ABI against the CPU oracle (oracle/t5_oracle.py,
pinned to the reference's T5Encoder by tests/golden/t5_encoder.pt) and against the golden outputs themselves."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("L,kv_len,H", [(24, 17, 2), (128, 128, 3), (300, 263, 2), (512, 512, 4), (512, 77, 4),
                                        (257, 1, 2), (333, 256, 1)])
def test_t5_attention_vs_fp32(mv, L, kv_len, H):
    """softmax(q k^T + bias + mask) v; tolerance 3e-3 rel-L2 (bf16 P and output rounding), SURVEY §8c per-kernel bar."""
    g = torch.Generator().manual_seed(L * 7 + kv_len)
    qkv = (torch.randn(L, 3 * H * 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
    q, k, v = (qkv[:, i * H * 64:(i + 1) * H * 64].unflatten(1, (H, 64)) for i in range(3))
    table = torch.randn(H, 2 * L - 1, generator=g).cuda()
    out = torch.full((L, H, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    mv.t5_attention(q, k, v, out, table, L - 1, kv_len)
    i = torch.arange(L, device="cuda").view(-1, 1)
    j = torch.arange(L, device="cuda").view(1, -1)
    s = torch.einsum("inc,jnc->nij", q.float(), k.float()) + table[:, (j - i) + (L - 1)]
    s = s.masked_fill(j.view(1, 1, -1) >= kv_len, float("-inf"))
    ref = torch.einsum("nij,jnc->inc", torch.softmax(s, -1), v.float())
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 3e-3


def test_t5_attention_without_bias_and_errors(mv):
    L, H = 64, 2
    q = torch.randn(L, H, 64).to(torch.bfloat16).cuda()
    out = torch.empty_like(q)
    mv.t5_attention(q, q, q, out)
    ref = torch.einsum("nij,jnc->inc", torch.softmax(torch.einsum("inc,jnc->nij", q.float(), q.float()), -1), q.float())
    assert rel_l2(out, ref) < 3e-3
    big = torch.zeros(513, H, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="at most 512 keys"):
        mv.t5_attention(big, big, big, torch.empty_like(big))
    with pytest.raises(RuntimeError, match="kv_len"):
        mv.t5_attention(q, q, q, out, kv_len=0)


def test_t5_row_kernels(mv):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 4096, generator=g).cuda()
    w = (1 + 0.1 * torch.randn(4096, generator=g)).cuda()
    out = torch.empty(37, 4096, dtype=torch.bfloat16, device="cuda")
    mv.t5_rmsnorm(x, w, out, 1e-6)
    y = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16).float() * w
    assert torch.equal(out, y.to(torch.bfloat16))            # same rounding points -> bit-exact
    table = torch.randn(97, 256, generator=g).to(torch.bfloat16).cuda()
    ids = torch.randint(0, 97, (50,), generator=g).cuda()
    emb = torch.empty(50, 256, device="cuda")
    mv.embed_gather(table, ids, emb)
    assert torch.equal(emb, table[ids].float())
    a = torch.randn(33, 264, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(33, 264, generator=g).to(torch.bfloat16).cuda()
    o = torch.empty_like(a)
    mv.mul_bf16(a, b, o)
    assert torch.equal(o, (a.float() * b.float()).to(torch.bfloat16))


def _tiny_encoder(g):
    from oracle.fill import fill_t5
    from wan.modules.t5 import T5Encoder
    m = T5Encoder(**g["cfg"]).eval().requires_grad_(False)
    fill_t5(m, g["seed"])
    return m.to(device="cuda", dtype=torch.bfloat16)


@pytest.mark.parametrize("case", [(24, 17), (300, 263), (40, 40)])
def test_t5_encoder_vs_golden_and_oracle(mv, case):
    """Tiny umT5-style encoder: CUDA path vs the reference's own fp32 output (golden) and vs the oracle with the
    kernels' bf16 cast points.  Bars (§8c): rel-L2 <= 1.5e-2 vs fp32, and no worse than 1.5x the bf16-emulating
    oracle's own error."""
    from oracle import t5_oracle as T
    g = torch.load(os.path.join(GOLD, "t5_encoder.pt"), weights_only=False)
    c = g["cases"][case]
    m = _tiny_encoder(g)
    n = case[1]
    y = m(c["ids"][None].cuda(), c["mask"][None].cuda())[0].float().cpu()
    sd = {k: v.float().cpu() for k, v in m.state_dict().items()}    # bf16-rounded weights, as the kernels see them
    cfg = g["cfg"]
    emu = T.t5_encoder_forward(sd, c["ids"], c["mask"], cfg["num_heads"], cfg["num_buckets"], cfg["shared_pos"],
                               rb=T.bf16_rt)
    err, err_emu = rel_l2(y[:n], c["y"][:n]), rel_l2(emu[:n], c["y"][:n])
    assert err < 1.5e-2, err
    assert err < 1.5 * err_emu + 1e-3, (err, err_emu)
    assert rel_l2(y[:n], emu[:n]) < 1e-2
    # prefix-only encoding (what T5EncoderModel.__call__ does) gives the same valid rows
    yp = m.encode_prefix(c["ids"].cuda(), n).float().cpu()
    assert rel_l2(yp, y[:n]) < 2e-3


def test_t5_umt5_width_layer_vs_oracle(mv):
    """One encoder block at umT5-XXL width (dim 4096, 64 heads, ffn 10240) on a 512-token prompt with 200 valid
    tokens, against the fp32 oracle on the host."""
    from oracle import t5_oracle as T
    from oracle.fill import fill_t5
    from wan.modules.t5 import T5Encoder
    cfg = dict(vocab=512, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=1, num_buckets=32,
               shared_pos=False)
    m = T5Encoder(**cfg).eval().requires_grad_(False)
    fill_t5(m, 77)
    m = m.to(device="cuda", dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(1, 512, (512,), generator=g)
    ids[200:] = 0
    mask = (torch.arange(512) < 200).long()
    y = m(ids[None].cuda(), mask[None].cuda())[0].float().cpu()
    sd = {k: v.float().cpu() for k, v in m.state_dict().items()}
    ref = T.t5_encoder_forward(sd, ids, mask, 64, 32, False)
    assert torch.isfinite(y).all()
    assert rel_l2(y[:200], ref[:200]) < 1.5e-2
    assert rel_l2(y, ref) < 1.5e-2          # padded query rows follow the reference too (forward(), not the prefix path)
