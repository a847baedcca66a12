"""CPU oracle for the WanVAE decoder and encoder — TEST INFRASTRUCTURE ONLY (see oracle/dit_oracle.py for the rules).

Restates WanVAE.decode (/root/reference/wan/modules/vae.py:657-663 -> WanVAE_.decode :544-568 -> Decoder3d.forward
:423-472) functionally over the reference-named state dict, as ONE pass over the whole latent sequence: the
reference's 21 single-frame chunks with a 2-frame feature cache are mathematically a causal network (every
CausalConv3d sees two zero frames in front), except that each `upsample3d` applies its time_conv to frames 1..T-1
only and passes frame 0 through (the 'Rep' branch, :106-131) — SURVEY.md Appendix B.  Pinned against outputs of the
reference's own chunked decode for T = 1, 2, 3, 5 (tests/golden/vae_*.pt, oracle/make_golden.py).

`encode` restates WanVAE.encode (:650-655 -> WanVAE_.encode :516-542 -> Encoder3d.forward :323-366) the same way: the
reference's 1 + 4 + 4 + ... frame chunks with the feature cache are a causal network, except that each `downsample3d`
passes frame 0 through and computes y_k = time_conv(x_{2k-2}, x_{2k-1}, x_{2k}) for k >= 1 (:143-159).  Pinned against
the reference's own chunked encode for T = 1, 5, 9, 13 (tests/golden/vae_encode.pt).

`rb` emulates the storage contract of the CUDA path (rb = f16_rt: fp16 activations between layers, fp16 conv operands,
fp32 accumulation; bf16_rt: the round-1 contract, kept for comparison); rb = ident gives the reference's fp32 semantics.
"""
import math

import torch
import torch.nn.functional as F

VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
            -0.1922, -0.9497, 0.2503, -0.2921]                                                   # vae.py:629-632
VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
           1.1253, 2.8251, 1.9160]                                                               # vae.py:633-636


def ident(x):
    return x


def bf16_rt(x):
    return x.to(torch.bfloat16).to(torch.float32)


def f16_rt(x):
    """fp16 round trip: the storage / operand contract of the CUDA decoder (csrc/vae_conv_sm100.cu)."""
    return x.to(torch.float16).to(torch.float32)


def causal_conv3d(x, w, b, rb=ident):
    """vae.py:17-36 with an empty cache: zero-pad W,H symmetrically, T in front only (2*pad_t)."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(rb(x), (kw // 2, kw // 2, kh // 2, kh // 2, 2 * (kt // 2), 0))
    return F.conv3d(x, rb(w.float()), None if b is None else b.float())


def rms_norm(x, gamma, channel_dim=1):
    """vae.py:39-54: F.normalize(x, dim=C) * sqrt(C) * gamma."""
    C = x.shape[channel_dim]
    shape = [1] * x.dim()
    shape[channel_dim] = C
    return F.normalize(x, dim=channel_dim) * math.sqrt(C) * gamma.float().reshape(shape)


def residual_block(sd, pre, x, rb=ident):
    """vae.py:186-220."""
    h = x
    if (pre + "shortcut.weight") in sd:
        h = rb(causal_conv3d(x, sd[pre + "shortcut.weight"], sd[pre + "shortcut.bias"], rb))
    y = rb(F.silu(rms_norm(x, sd[pre + "residual.0.gamma"])))
    y = rb(causal_conv3d(y, sd[pre + "residual.2.weight"], sd[pre + "residual.2.bias"], rb))
    y = rb(F.silu(rms_norm(y, sd[pre + "residual.3.gamma"])))
    y = causal_conv3d(y, sd[pre + "residual.6.weight"], sd[pre + "residual.6.bias"], rb)
    return rb(y + h)


def attention_block(sd, pre, x, rb=ident):
    """vae.py:223-262: per-frame single-head attention over the h*w positions."""
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    n = rb(rms_norm(xf, sd[pre + "norm.gamma"]))
    qkv = rb(F.conv2d(n, rb(sd[pre + "to_qkv.weight"].float()), sd[pre + "to_qkv.bias"].float()))
    q, k, v = qkv.reshape(b * t, 3 * c, h * w).permute(0, 2, 1).chunk(3, dim=-1)        # [bt, hw, c] each
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(c)
    p = rb(torch.softmax(s, dim=-1))
    o = rb(torch.matmul(p, v))                                                            # [bt, hw, c]
    o = o.permute(0, 2, 1).reshape(b * t, c, h, w)
    o = F.conv2d(o, rb(sd[pre + "proj.weight"].float()), sd[pre + "proj.bias"].float())
    o = o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return rb(o + x)


def resample_up(sd, pre, x, mode, rb=ident):
    """vae.py:66-141 for 'upsample2d' / 'upsample3d' in whole-sequence form."""
    b, c, t, h, w = x.shape
    if mode == "upsample3d" and t > 1:
        tail = rb(causal_conv3d(x[:, :, 1:], sd[pre + "time_conv.weight"], sd[pre + "time_conv.bias"], rb))
        tail = tail.reshape(b, 2, c, t - 1, h, w)
        tail = torch.stack((tail[:, 0], tail[:, 1]), 3).reshape(b, c, 2 * (t - 1), h, w)   # interleave :133-137
        x = torch.cat([x[:, :, :1], tail], dim=2)
    t2 = x.shape[2]
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t2, c, h, w)
    xf = F.interpolate(xf.float(), scale_factor=(2.0, 2.0), mode="nearest-exact")
    xf = F.conv2d(rb(xf), rb(sd[pre + "resample.1.weight"].float()), sd[pre + "resample.1.bias"].float(), padding=1)
    return rb(xf.reshape(b, t2, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4))


def decoder_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)):
    """Module list of Decoder3d.upsamples (vae.py:397-416): [('res', in, out) | ('up3d'|'up2d', dim)]."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    plan = []
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            in_dim = in_dim // 2
        for _ in range(num_res_blocks + 1):
            plan.append(("res", in_dim, out_dim))
            in_dim = out_dim
        if i != len(dim_mult) - 1:
            plan.append(("up3d" if temperal_upsample[i] else "up2d", out_dim))
    return dims, plan


def decode(sd, z, rb=ident, dim=96, z_dim=16):
    """WanVAE.decode for one latent z [z_dim, T, h, w] fp32 -> [3, 1+4(T-1), 8h, 8w] fp32 in [-1, 1]."""
    mean = torch.tensor(VAE_MEAN[:z_dim]).view(1, z_dim, 1, 1, 1)
    std = torch.tensor(VAE_STD[:z_dim]).view(1, z_dim, 1, 1, 1)
    x = z.float().unsqueeze(0) / (1.0 / std) + mean                                      # vae.py:547-551
    x = rb(causal_conv3d(x, sd["conv2.weight"], sd["conv2.bias"]))                       # :553 (fp32 in the kernel too)
    x = rb(causal_conv3d(x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"], rb))
    x = residual_block(sd, "decoder.middle.0.", x, rb)
    x = attention_block(sd, "decoder.middle.1.", x, rb)
    x = residual_block(sd, "decoder.middle.2.", x, rb)
    _, plan = decoder_plan(dim)
    for i, item in enumerate(plan):
        pre = "decoder.upsamples.%d." % i
        if item[0] == "res":
            x = residual_block(sd, pre, x, rb)
        else:
            x = resample_up(sd, pre, x, "upsample3d" if item[0] == "up3d" else "upsample2d", rb)
    x = rb(F.silu(rms_norm(x, sd["decoder.head.0.gamma"])))
    x = causal_conv3d(x, sd["decoder.head.2.weight"], sd["decoder.head.2.bias"], rb)
    return x[0].float().clamp_(-1, 1)


def resample_down(sd, pre, x, mode, rb=ident):
    """vae.py:66-160 for 'downsample2d' / 'downsample3d' in whole-sequence form: ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2)
    per frame; 'downsample3d' then applies time_conv (3,1,1) stride (2,1,1) WITHOUT temporal padding to
    [last frame of the previous chunk | chunk]; the first chunk (frame 0) skips it (:146-148)."""
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    xf = F.conv2d(F.pad(rb(xf), (0, 1, 0, 1)), rb(sd[pre + "resample.1.weight"].float()),
                  sd[pre + "resample.1.bias"].float(), stride=2)
    x = rb(xf.reshape(b, t, c, xf.shape[2], xf.shape[3]).permute(0, 2, 1, 3, 4))
    if mode == "downsample3d" and t > 1:
        tail = F.conv3d(rb(x), rb(sd[pre + "time_conv.weight"].float()), sd[pre + "time_conv.bias"].float(),
                        stride=(2, 1, 1))                          # frames (0,1,2), (2,3,4), ... = y_1, y_2, ...
        x = torch.cat([x[:, :, :1], rb(tail)], dim=2)
    return x


def encoder_plan(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    """Module list of Encoder3d.downsamples (vae.py:288-305): [('res', in, out) | ('down3d'|'down2d', dim)]."""
    dims = [dim * u for u in [1] + list(dim_mult)]
    plan = []
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            plan.append(("res", in_dim, out_dim))
            in_dim = out_dim
        if i != len(dim_mult) - 1:
            plan.append(("down3d" if temperal_downsample[i] else "down2d", out_dim))
    return dims, plan


def encode(sd, video, rb=ident, dim=96, z_dim=16):
    """WanVAE.encode for one video [3, 1+4k, H, W] fp32 in [-1, 1] -> mu [z_dim, 1+k, H/8, W/8] fp32, normalised."""
    assert (video.shape[1] - 1) % 4 == 0, "the reference drops trailing frames that do not fill a 4-frame chunk (:521)"
    x = rb(causal_conv3d(video.float().unsqueeze(0), sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], rb))
    _, plan = encoder_plan(dim)
    for i, item in enumerate(plan):
        pre = "encoder.downsamples.%d." % i
        if item[0] == "res":
            x = residual_block(sd, pre, x, rb)
        else:
            x = resample_down(sd, pre, x, "downsample3d" if item[0] == "down3d" else "downsample2d", rb)
    x = residual_block(sd, "encoder.middle.0.", x, rb)
    x = attention_block(sd, "encoder.middle.1.", x, rb)
    x = residual_block(sd, "encoder.middle.2.", x, rb)
    x = rb(F.silu(rms_norm(x, sd["encoder.head.0.gamma"])))
    x = rb(causal_conv3d(x, sd["encoder.head.2.weight"], sd["encoder.head.2.bias"], rb))
    mu = F.conv3d(x, sd["conv1.weight"].float(), sd["conv1.bias"].float())[:, :z_dim]   # :531 (fp32 in the kernel too)
    mean = torch.tensor(VAE_MEAN[:z_dim]).view(1, z_dim, 1, 1, 1)
    inv_std = (1.0 / torch.tensor(VAE_STD[:z_dim])).view(1, z_dim, 1, 1, 1)
    return ((mu - mean) * inv_std)[0]                                                   # :532-537
