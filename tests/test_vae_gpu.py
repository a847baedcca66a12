"""-m gpu parity of the native WanVAE decoder (implicit-GEMM tcgen05 convs) against golden outputs of the REFERENCE's
chunked decode and against the CPU oracle.  Tolerance: PSNR >= 45 dB and max-abs <= 2e-2 on the
[-1, 1] output.  The decoder runs with FP16 operands / activations (fp32 accumulation) where the reference uses fp32
storage and TF32 convolutions; the CPU emulation of that contract measures max-abs 2e-3..4.3e-3 and PSNR 67-72 dB
against the reference goldens (the bf16 contract of round 1 measured 1.4e-2..3.6e-2), so the bounds asserted here are
the §8c ones with margin: max-abs <= 1e-2, PSNR >= 60 dB."""
import math
import os

import pytest
import torch

from oracle import vae_oracle as V
from oracle.fill import fill_parameters, state_dict_like

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return 10 * math.log10(4.0 / max(mse, 1e-20))


@pytest.fixture(scope="module")
def vae():
    from wan.modules.vae import WanVAE_
    g = torch.load(os.path.join(GOLD, "vae_decode.pt"), weights_only=False)
    m = WanVAE_().eval().requires_grad_(False)
    fill_parameters(m, g["seed"])
    return m.to(DEV), g


@pytest.fixture(params=[0, 1], ids=["conv1cta", "convpair"])
def conv_kernel(request):
    """Both convolution kernels: the single-CTA implicit GEMM and the CTA-pair one (cta_group::2, stacked tiles, one A box
    per (dt, dw)); the product default is whichever measured faster (csrc/vae_conv_sm100.cu::kDefaultConvPair)."""
    import movii_b200 as mv
    mv.vae_conv_config(request.param)
    yield request.param
    mv.vae_conv_config(-2, -2)


@pytest.mark.parametrize("case", [(1, 4, 6), (2, 5, 9), (3, 4, 6), (5, 4, 4)])
def test_vae_decode_matches_reference_golden(vae, case, conv_kernel):
    m, g = vae
    rec = g["cases"][case]
    y = m.decode(rec["z"][None].to(DEV))[0].cpu()
    ref = rec["y"].float()
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert torch.isfinite(y).all() and y.abs().max() <= 1.0
    sd = state_dict_like(g["param_shapes"], g["seed"])
    emu = V.decode(sd, rec["z"], V.f16_rt)
    assert psnr(y, ref) >= 60.0, (psnr(y, ref), psnr(emu, ref))
    assert (y - ref).abs().max().item() <= 1e-2, ((y - ref).abs().max().item(), (emu - ref).abs().max().item())
    # against the CPU emulation of the same storage contract (it differs in summation order and in the pre-summed
    # sub-pixel upsample weights): same error class as either one against the fp32 reference
    assert psnr(y, emu) >= 60.0


@pytest.mark.parametrize("case", [(5, 4, 4), (3, 4, 6)])
def test_vae_temporal_chunking_is_exact(vae, case, conv_kernel):
    """The decode runs in temporal chunks with a two-frame feature cache per temporal conv (the reference's scheme,
    vae.py:28-36,205-217, with larger chunks).  Any chunk size must give the SAME BITS as the whole-sequence pass:
    same taps, same accumulation order, only the provenance of the two history frames differs — including chunk 1
    (the reference's own chunking: frame 0 alone through the 'Rep' branch) and chunks that do not divide T."""
    m, g = vae
    rec = g["cases"][case]
    eng = m.engine()
    z = rec["z"].to(DEV)
    keep = eng.chunk
    try:
        eng.chunk = 64
        whole = eng.decode(z).clone()
        for c in (1, 2, 3, 4):
            eng.chunk = c
            assert torch.equal(eng.decode(z), whole), c
    finally:
        eng.chunk = keep
    assert (whole.cpu() - rec["y"].float()).abs().max().item() <= 1e-2


def test_vae_conv_primitives(vae, conv_kernel):
    """One 3x3x3 causal conv and one sub-pixel upsample conv against torch, ragged grid."""
    import torch.nn.functional as F
    import movii_b200 as mv
    from wan.modules.vae import _Conv, _parity_convs, _taps
    g = torch.Generator().manual_seed(1)
    T, H, W, Ci, Co = 3, 11, 21, 96, 192
    x = torch.randn(T, H, W, Ci, generator=g).half()
    wt = (torch.randn(Co, Ci, 3, 3, 3, generator=g) / math.sqrt(27 * Ci))
    b = torch.randn(Co, generator=g)
    c = _Conv(wt, b, _taps(3, 3, 3), DEV)
    out = torch.empty(T, H, W, Co, dtype=torch.float16, device=DEV)
    mv.vae_conv(x.to(DEV), c, out, o_base=0, os_t=H * W * Co, os_h=W * Co, os_w=Co)
    ref = V.causal_conv3d(x.float().permute(3, 0, 1, 2)[None], wt, b, V.f16_rt)[0].permute(1, 2, 3, 0)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 4e-3, err      # fp16 output rounding of O(1) values: 2^-11 relative
    # sub-pixel upsample: nearest-exact x2 + Conv2d 3x3
    w2 = torch.randn(Co // 2, Co, 3, 3, generator=g) / math.sqrt(9 * Co)
    b2 = torch.randn(Co // 2, generator=g)
    xin = torch.randn(T, H, W, Co, generator=g).half()
    up = torch.empty(T, 2 * H, 2 * W, Co // 2, dtype=torch.float16, device=DEV)
    for (a, bb), cc in _parity_convs(w2, b2, DEV).items():
        mv.vae_conv(xin.to(DEV), cc, up, o_base=(a * 2 * W + bb) * (Co // 2), os_t=4 * H * W * (Co // 2),
                    os_h=4 * W * (Co // 2), os_w=2 * (Co // 2))
    xf = F.interpolate(xin.float().permute(0, 3, 1, 2), scale_factor=(2.0, 2.0), mode="nearest-exact")
    ref2 = F.conv2d(xf, V.f16_rt(w2), b2, padding=1).permute(0, 2, 3, 1)
    err2 = (up.float().cpu() - ref2).abs().max().item()
    assert err2 <= 6e-3, err2    # four pre-summed fp16 sub-pixel kernels vs one 3x3 kernel on the upsampled image


@pytest.mark.parametrize("shape", [(3, 70, 40, 96, 96), (2, 33, 50, 192, 192), (2, 40, 24, 384, 384), (3, 19, 17, 96, 16),
                                   (2, 64, 32, 384, 768)])
@pytest.mark.parametrize("mode", ["plain", "res", "fused"])
@pytest.mark.parametrize("nt", [0, 1, 2, 4])
def test_vae_conv_pair_matches_single_cta(shape, mode, nt):
    """The CTA-pair kernel vs the single-CTA kernel on the decoder's channel plans (96, 192, 384 -> 2 N blocks, the
    16-wide padded head, the 768-wide time_conv with its channel-block split), ragged grids (H, W not multiples of the
    super-tile), two cached frames in front (t_off = 2), with residual / with the fused RMS_norm+SiLU epilogue.
    Same operands, same fp32 accumulation, different summation order: differences are fp16 output rounding flips."""
    import movii_b200 as mv
    from wan.modules.vae import _Conv, _taps
    T, H, W, Ci, Co = shape
    if mode == "fused" and Co > 256:
        pytest.skip("the fused norm epilogue needs the whole channel row in one tile")
    g = torch.Generator().manual_seed(H * W + Ci + Co)
    x = torch.randn(T + 2, H, W, Ci, generator=g).half().to(DEV)
    time_conv = Co == 768
    taps = _taps(3, 1, 1) if time_conv else _taps(3, 3, 3)
    wt = torch.randn(Co, Ci, 3, 1 if time_conv else 3, 1 if time_conv else 3, generator=g) / math.sqrt(len(taps) * Ci)
    c = _Conv(wt, 0.1 * torch.randn(Co, generator=g), taps, DEV)
    res = torch.randn(T, H, W, c.cout, generator=g).half().to(DEV) if mode == "res" else None
    gamma = (1 + 0.1 * torch.randn(c.cout, generator=g)).to(DEV)
    outs = []
    for pair in (0, 1):
        mv.vae_conv_config(pair, nt)          # nt: tiles per CTA of the pair kernel (0 = automatic choice)
        try:
            if time_conv:      # frame interleave store: channel block >= 384 goes to the next frame
                Ch = Co // 2
                out = torch.zeros(2 * T, H, W, Ch, dtype=torch.float16, device=DEV)
                fe = H * W * Ch
                mv.vae_conv(x, c, out, o_base=0, os_t=2 * fe, os_h=W * Ch, os_w=Ch, nsplit=Ch, nsplit_off=fe, t_off=2)
            elif mode == "fused":
                out = torch.zeros(T, H, W, c.cout, dtype=torch.float16, device=DEV)
                raw = torch.zeros_like(out)
                mv.vae_conv_fused(x, c, raw, gamma, out, o_base=0, os_t=H * W * c.cout, os_h=W * c.cout, os_w=c.cout, t_off=2)
                out = torch.cat([out, raw])
            else:
                out = torch.zeros(T, H, W, c.cout, dtype=torch.float16, device=DEV)
                mv.vae_conv(x, c, out, res=res, o_base=0, os_t=H * W * c.cout, os_h=W * c.cout, os_w=c.cout, t_off=2)
            torch.cuda.synchronize()
        finally:
            mv.vae_conv_config(-2, -2)
        outs.append(out.float())
    a, b = outs
    assert torch.isfinite(b).all()
    assert (a - b).abs().max().item() <= 4e-3 * max(1.0, a.abs().max().item()), (a - b).abs().max().item()
    assert ((a - b).norm() / a.norm()).item() <= 5e-4
