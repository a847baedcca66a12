"""WanVAE with the reference's wrapper surface (wan/modules/vae.py:619-663) and parameter names, run by the
B200-native implicit-GEMM kernels (csrc/vae_conv_sm100.cu) through the C ABI.

WanVAE.decode (the hot path) and WanVAE.encode (§8f-4: preprocessing / i2v side, same kernels plus
strided convolutions).  Both walk the sequence in temporal chunks with a feature cache like the reference; the
networks are causal, so any chunking gives the same bits.  The irregularities are reproduced exactly: upsample3d's
time_conv starts at frame 1 (the 'Rep' branch), downsample3d passes frame 0 through and strides over
[last cached frame | chunk] (pinned by tests/golden/vae_decode.pt / vae_encode.pt).
Activations and conv operands are channels-last FP16 [T, H, W, C] (the 10 mantissa bits the reference's TF32 convs keep;
max-abs 4e-3 against the fp32 reference, bf16 storage measured 3e-2); accumulation is fp32.
"""
import logging
import math
import os

import torch
import torch.nn as nn

import movii_b200 as mv

__all__ = ["WanVAE"]

F16, F32 = torch.float16, torch.float32

VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
            -0.1922, -0.9497, 0.2503, -0.2921]
VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
           1.1253, 2.8251, 1.9160]


# ------------------------------------------------------------------------------------------------
# parameter holders (names == reference state-dict keys)
# ------------------------------------------------------------------------------------------------
class _Gamma(nn.Module):
    def __init__(self, dim, images):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones((dim, 1, 1) if images else (dim, 1, 1, 1)))


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.residual = nn.Sequential(_Gamma(cin, False), nn.Identity(), nn.Conv3d(cin, cout, 3), _Gamma(cout, False),
                                      nn.Identity(), nn.Identity(), nn.Conv3d(cout, cout, 3))
        self.shortcut = nn.Conv3d(cin, cout, 1) if cin != cout else nn.Identity()


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = _Gamma(dim, True)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)
        nn.init.zeros_(self.proj.weight)


class _Up(nn.Module):
    def __init__(self, dim, mode):
        super().__init__()
        self.mode = mode
        self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim // 2, 3, padding=1))
        if mode == "upsample3d":
            self.time_conv = nn.Conv3d(dim, dim * 2, (3, 1, 1))


class _Down(nn.Module):
    def __init__(self, dim, mode):
        super().__init__()
        self.mode = mode
        self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim, 3, stride=2))     # [0] = ZeroPad2d((0,1,0,1))
        if mode == "downsample3d":
            self.time_conv = nn.Conv3d(dim, dim, (3, 1, 1), stride=(2, 1, 1))


class _Encoder(nn.Module):
    """Parameter holder for Encoder3d (vae.py:265-321)."""

    def __init__(self, dim, z_dim, dim_mult, num_res_blocks, temperal_downsample):
        super().__init__()
        dims = [dim * u for u in [1] + list(dim_mult)]
        self.conv1 = nn.Conv3d(3, dims[0], 3)
        downs = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                downs.append(_Res(cin, cout))
                cin = cout
            if i != len(dim_mult) - 1:
                downs.append(_Down(cout, "downsample3d" if temperal_downsample[i] else "downsample2d"))
        self.downsamples = nn.Sequential(*downs)
        self.middle = nn.Sequential(_Res(cout, cout), _Attn(cout), _Res(cout, cout))
        self.head = nn.Sequential(_Gamma(cout, False), nn.Identity(), nn.Conv3d(cout, z_dim, 3))


class _Decoder(nn.Module):
    def __init__(self, dim, z_dim, dim_mult, num_res_blocks, temperal_upsample):
        super().__init__()
        dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
        self.conv1 = nn.Conv3d(z_dim, dims[0], 3)
        self.middle = nn.Sequential(_Res(dims[0], dims[0]), _Attn(dims[0]), _Res(dims[0], dims[0]))
        ups = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                cin = cin // 2
            for _ in range(num_res_blocks + 1):
                ups.append(_Res(cin, cout))
                cin = cout
            if i != len(dim_mult) - 1:
                ups.append(_Up(cout, "upsample3d" if temperal_upsample[i] else "upsample2d"))
        self.upsamples = nn.Sequential(*ups)
        self.head = nn.Sequential(_Gamma(dims[-1], False), nn.Identity(), nn.Conv3d(dims[-1], 3, 3))


class WanVAE_(nn.Module):
    """The reference's WanVAE_ (vae.py:481-514): `encoder` + `conv1`, `conv2` + `decoder`, same parameter names."""

    def __init__(self, dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_downsample=(False, True, True), dropout=0.0):
        super().__init__()
        if len(attn_scales) != 0:
            raise NotImplementedError("attn_scales must be empty (the shipped VAE config, vae.py:597-604)")
        self.dim, self.z_dim = dim, z_dim
        self.temperal_upsample = tuple(temperal_downsample)[::-1]
        self.encoder = _Encoder(dim, z_dim * 2, tuple(dim_mult), num_res_blocks, tuple(temperal_downsample))
        self.conv1 = nn.Conv3d(z_dim * 2, z_dim * 2, 1)
        self.conv2 = nn.Conv3d(z_dim, z_dim, 1)
        self.decoder = _Decoder(dim, z_dim, tuple(dim_mult), num_res_blocks, self.temperal_upsample)
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = VaeEngine(self)
        return self._engine

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine = None
        return r

    def load_state_dict(self, sd, strict=True, **k):
        """A state dict holding only one half (decoder + conv2, or encoder + conv1) leaves the other half untouched."""
        mine = self.state_dict()
        for half in (("decoder.", "conv2."), ("encoder.", "conv1.")):
            if not any(n.startswith(half) for n in sd):
                sd = {**{n: v for n, v in mine.items() if n.startswith(half)}, **sd}
        r = super().load_state_dict(sd, strict=strict, **k)
        self._engine = None
        return r

    def decode(self, z, scale=None):
        """z [1, z_dim, T, h, w] -> [1, 3, 1+4(T-1), 8h, 8w] (vae.py:544-568); scale is accepted for API parity and
        must be the standard (mean, 1/std) pair."""
        return self.engine().decode(z[0]).unsqueeze(0)

    def encode(self, x, scale=None):
        """x [1, 3, 1+4k, H, W] in [-1, 1] -> normalised mu [1, z_dim, 1+k, H/8, W/8] (vae.py:516-542)."""
        return self.engine().encode(x[0]).unsqueeze(0)


# ------------------------------------------------------------------------------------------------
# engine
# ------------------------------------------------------------------------------------------------
def _taps(kt, kh, kw):
    """(dt, dh, dw) per tap in weight order; temporal taps are causal (vae.py:24-36)."""
    return [(it - 2 * (kt // 2), ih - kh // 2, iw - kw // 2) for it in range(kt) for ih in range(kh) for iw in range(kw)]


class _Conv:
    """One packed convolution: fp16 [Cout_pad][taps][Cin] + fp32 bias + tap offsets."""

    def __init__(self, weight, bias, taps, device, cout_pad=None):
        cout, cin = weight.shape[0], weight.shape[1]
        w = weight.detach().to(F32).reshape(cout, cin, -1).permute(0, 2, 1).contiguous()      # [Cout, taps, Cin]
        self.cout_real = cout
        cp = cout_pad or ((cout + 15) // 16 * 16)
        if cp != cout:
            w = torch.cat([w, w.new_zeros(cp - cout, *w.shape[1:])])
        self.w = w.to(F16).contiguous().to(device)
        b = None if bias is None else bias.detach().to(F32)
        if b is not None and cp != cout:
            b = torch.cat([b, b.new_zeros(cp - cout)])
        self.b = None if b is None else b.contiguous().to(device)
        self.cin, self.cout = cin, cp
        self.taps = torch.tensor(taps, dtype=torch.int8).contiguous()
        self.ntaps = len(taps)


def _parity_weights(weight):
    """nearest-exact x2 followed by a 3x3 Conv2d == four 2x2 convolutions on the low-res grid, one per output
    parity (a, b) (vae.py:74-79).  Rows 2h+a-1, 2h+a, 2h+a+1 of the upsampled image are source rows
    {h-1, h, h} (a=0) or {h, h, h+1} (a=1); the kernel rows falling on the same source row are summed (in fp32).
    Returns {(a, b): (taps [(0, dh, dw)] * 4, weight [Co, Ci, 4, 1, 1])}."""
    groups = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}
    w = weight.detach().to(F32)                                                               # [Co, Ci, 3, 3]
    out = {}
    for a in (0, 1):
        for b in (0, 1):
            taps, mats = [], []
            for dh, khs in groups[a]:
                for dw, kws in groups[b]:
                    taps.append((0, dh, dw))
                    mats.append(sum(w[:, :, kh, kw] for kh in khs for kw in kws))
            out[(a, b)] = (taps, torch.stack(mats, dim=2).unsqueeze(-1).unsqueeze(-1))      # [Co, Ci, 4, 1, 1]
    return out


def _parity_convs(weight, bias, device):
    return {ab: _Conv(wp, bias, taps, device) for ab, (taps, wp) in _parity_weights(weight).items()}


def _head_tap_matrix(weight):
    """Head conv weight [3, C, 3, 3, 3] -> [112, C]: row 4*tap + co = W[co, :, tap] (tap in weight order (kt, kh, kw), co < 3;
    the fourth row of every tap and rows 108..111 are zero).  x @ rows gives, per voxel, the 27 x 3 partial sums that
    mv_vae_head_gather adds over the voxel's neighbours."""
    wh = weight.detach().to(F32)
    wt = wh.new_zeros(27, 4, wh.shape[1])
    wt[:, :3] = wh.reshape(3, wh.shape[1], 27).permute(2, 0, 1)                    # [tap, co, c]
    return torch.cat([wt.reshape(108, wh.shape[1]), wh.new_zeros(4, wh.shape[1])])   # 112 rows (a multiple of 16)


class VaeEngine:
    def __init__(self, model):
        dev = model.conv2.weight.device
        if dev.type != "cuda":
            raise RuntimeError("WanVAE must live on a CUDA (B200) device: movii_b200 has no CPU path")
        self.device = dev
        self.z_dim = model.z_dim
        d = model.decoder
        f32 = lambda t: t.detach().to(F32).contiguous().to(dev)  # noqa: E731
        self.w2 = f32(model.conv2.weight.reshape(self.z_dim, self.z_dim))
        self.b2 = f32(model.conv2.bias)
        self.mean = torch.tensor(VAE_MEAN[:self.z_dim], dtype=F32, device=dev)
        self.std = torch.tensor(VAE_STD[:self.z_dim], dtype=F32, device=dev)
        self.conv1 = _Conv(d.conv1.weight, d.conv1.bias, _taps(3, 3, 3), dev)
        self.layers = []
        for m in list(d.middle) + list(d.upsamples):
            if isinstance(m, _Res):
                self.layers.append(("res", self._res(m)))
            elif isinstance(m, _Attn):
                self.layers.append(("attn", dict(
                    gamma=f32(m.norm.gamma.reshape(-1)),
                    qkv=_Conv(m.to_qkv.weight.unsqueeze(2), m.to_qkv.bias, _taps(1, 1, 1), dev),
                    proj=_Conv(m.proj.weight.unsqueeze(2), m.proj.bias, _taps(1, 1, 1), dev))))
            else:
                up = dict(mode=m.mode, par=_parity_convs(m.resample[1].weight, m.resample[1].bias, dev))
                if m.mode == "upsample3d":
                    up["time"] = _Conv(m.time_conv.weight, m.time_conv.bias, _taps(3, 1, 1), dev)
                self.layers.append(("up", up))
        self.operand_dtype = "f16"
        self.head_gamma = f32(d.head[0].gamma.reshape(-1))
        self.head = _Conv(d.head[2].weight, d.head[2].bias, _taps(3, 3, 3), dev, cout_pad=16)
        # The same head conv as ONE 1x1x1 conv to 27 taps x (3 + 1 pad) partial sums per voxel + a gather over the 27
        # neighbours (mv_vae_head_gather): 6 MMAs per voxel tile instead of 162 sixteen-column ones (the 16-wide conv ran
        # at 219 TFLOP/s executed, 4.5 % of the decode).  MOVII_VAE_HEAD=conv keeps the direct conv (A/B).
        wt = _head_tap_matrix(d.head[2].weight)
        self.head_taps = _Conv(wt.reshape(112, wt.shape[1], 1, 1, 1), None, _taps(1, 1, 1), dev)
        self.head_bias = [float(v) for v in d.head[2].bias.detach().to(F32).cpu()]
        self.head_mode = os.environ.get("MOVII_VAE_HEAD", "gather")
        # temporal chunk (latent frames per pass): the decoder is a causal network, every temporal conv carries the
        # last two frames of its input to the next chunk — the reference's feature cache (vae.py:28-36,205-217) with
        # chunks of `chunk` latent frames instead of one.  Bounds the live activations (1080P: 3 stage-D tensors of
        # 16 frames instead of 81).
        self.chunk = max(1, int(os.environ.get("MOVII_VAE_CHUNK", "4")))
        self.cache = {}
        # A/B switch (tools/vae_bench.py): 0 = the block's last conv does not emit the consumer's norm (stand-alone pass)
        self.fuse_c6 = os.environ.get("MOVII_VAE_FUSE_C6", "1") != "0"
        # CUDA-graph replay of the decode: measured 73.3 fps vs 73.7 fps with direct launches on the same box — the host
        # keeps ahead of the millisecond-long kernels, there are no gaps to remove — so it is opt-in (and it pins the
        # 35 GB of activations of one decode in a private pool)
        self.use_graph = os.environ.get("MOVII_VAE_GRAPH", "0") == "1"
        self._graphs = {}
        self.model = model
        self.enc = None              # encoder plan, packed on the first encode()

    def _res(self, m):
        dev = self.device
        f32 = lambda t: t.detach().to(F32).contiguous().to(dev)  # noqa: E731
        r = m.residual
        return dict(g0=f32(r[0].gamma.reshape(-1)), c2=_Conv(r[2].weight, r[2].bias, _taps(3, 3, 3), dev),
                    g3=f32(r[3].gamma.reshape(-1)), c6=_Conv(r[6].weight, r[6].bias, _taps(3, 3, 3), dev),
                    sc=None if isinstance(m.shortcut, nn.Identity) else
                    _Conv(m.shortcut.weight, m.shortcut.bias, _taps(1, 1, 1), dev))

    # -- temporal feature cache ----------------------------------------------------------------------------
    def halo_buffer(self, key, n, H, W, C):
        """Input buffer of a temporal conv for an n-frame chunk: [k + n, H, W, C] whose first k (<= 2) frames are the
        cached tail of the previous chunk's input; the producer writes frames [k:].  Returns (buffer, k)."""
        cache = self.cache.get(key)
        k = 0 if cache is None else cache.shape[0]
        buf = torch.empty(k + n, H, W, C, dtype=F16, device=self.device)
        if k:
            buf[:k].copy_(cache)
        return buf, k

    def commit(self, key, buf):
        """Remember the last two frames of this conv's input (cache + chunk) for the next chunk (vae.py:205-217)."""
        self.cache[key] = buf[-2:].clone() if buf.shape[0] > 2 else buf.clone()

    # -- primitive launches -----------------------------------------------------------------------------
    def conv(self, x, c, res=None, out=None, t_off=0):
        """x [t_off + T,H,W,Cin] fp16 -> [T,H,W,Cout] fp16 (+res)."""
        Tin, H, W, _ = x.shape
        T = Tin - t_off
        if out is None:
            out = torch.empty(T, H, W, c.cout, dtype=F16, device=self.device)
        mv.vae_conv(x, c, out, res=res, o_base=0, os_t=H * W * c.cout, os_h=W * c.cout, os_w=c.cout, t_off=t_off)
        return out

    def normsilu(self, x, gamma, out=None, silu=True):
        out = x if out is None else out
        mv.vae_rmsnorm_silu(x, out, gamma, silu)
        return out

    # -- blocks -----------------------------------------------------------------------------------------------
    # Every conv of a ResidualBlock consumes silu(rms_norm(.)).  Where the producing conv holds a whole channel row
    # per thread (Cout <= 256: stages C and D, i.e. ~90 % of the bytes) that normalisation is fused into ITS epilogue
    # (mv_vae_conv_fused) and the stand-alone pass over HBM disappears.
    @staticmethod
    def _fusable(c):
        return c.cout <= 256 and c.cout == c.cout_real

    def resblock(self, x, p, key, a_in=None, nxt=None):
        """x -> x + conv(silu(norm(conv(silu(norm(x))))))  (vae.py:186-220).  a_in = (halo buffer, cached frames) when the
        PRODUCER of x already wrote silu(rms_norm(x) * g0) into this block's first conv input; nxt = (gamma, cache key)
        of the norm the CONSUMER applies to the result: when this block's last conv can hold a channel row per thread
        it emits that normalised tensor too (returned as a_next), and the stand-alone pass over HBM disappears."""
        n, H, W, _ = x.shape
        h = x if p["sc"] is None else self.conv(x, p["sc"])
        c2, c6 = p["c2"], p["c6"]
        if a_in is None:
            a, ka = self.halo_buffer(key + ".c2", n, H, W, c2.cin)
            self.normsilu(x, p["g0"], out=a[ka:])
        else:
            a, ka = a_in
        self.commit(key + ".c2", a)
        y, ky = self.halo_buffer(key + ".c6", n, H, W, c2.cout)
        if self._fusable(c2):
            mv.vae_conv_fused(a, c2, None, p["g3"], y[ky:], o_base=0, os_t=H * W * c2.cout, os_h=W * c2.cout, os_w=c2.cout,
                              t_off=ka)
        else:
            self.conv(a, c2, out=y[ky:], t_off=ka)
            self.normsilu(y[ky:], p["g3"])
        del a
        self.commit(key + ".c6", y)
        if nxt is not None and self._fusable(c6) and self.fuse_c6:
            gamma_n, key_n = nxt
            an, kn = self.halo_buffer(key_n, n, H, W, c6.cout)
            out = torch.empty(n, H, W, c6.cout, dtype=F16, device=self.device)
            mv.vae_conv_fused(y, c6, out, gamma_n, an[kn:], res=h, o_base=0, os_t=H * W * c6.cout, os_h=W * c6.cout,
                              os_w=c6.cout, t_off=ky)
            return out, (an, kn)
        return self.conv(y, c6, res=h, t_off=ky), None

    def attention(self, x, p):
        """vae.py:223-262: per-frame single-head attention, d = C, over the H*W positions."""
        T, H, W, C = x.shape
        hw = H * W
        n = self.normsilu(x, p["gamma"], out=torch.empty_like(x), silu=False)
        qkv = self.conv(n, p["qkv"]).view(T * hw, 3 * C)
        del n
        k8 = (hw + 7) // 8 * 8
        S = torch.empty(hw, (hw + 3) // 4 * 4, dtype=F32, device=self.device)
        P = torch.zeros(hw, k8, dtype=F16, device=self.device)
        vT = torch.zeros(C, k8, dtype=F16, device=self.device)
        O = torch.empty(T, H, W, C, dtype=F16, device=self.device)
        Of = O.view(T * hw, C)
        scale = 1.0 / math.sqrt(C)
        for f in range(T):
            rows = qkv[f * hw:(f + 1) * hw]
            mv.gemm_f16(rows[:, 0:C], rows[:, C:2 * C], None, S[:, :hw], mv.MV_EPI_F32)
            mv.softmax_rows(S[:, :hw], P, hw, scale)
            vT[:, :hw].copy_(rows[:, 2 * C:3 * C].t())       # layout change only (K-major operand for P.V)
            mv.gemm_f16(P, vT, None, Of[f * hw:(f + 1) * hw], mv.MV_EPI_BF16)
        return self.conv(O, p["proj"], res=x)

    def upsample(self, x, p, key, first, nxt=None):
        T, H, W, C = x.shape
        if p["mode"] == "upsample3d":
            # time_conv (3,1,1) C -> 2C + frame interleave; its stream starts at frame 1 of the sequence and never
            # sees frame 0, which passes through (the 'Rep' branch, vae.py:106-131)
            src = x[1:] if first else x
            nt = src.shape[0]
            T2 = 2 * nt + (1 if first else 0)
            y = torch.empty(T2, H, W, C, dtype=F16, device=self.device)
            fe = H * W * C
            if first:
                y[0].copy_(x[0])
            if nt > 0:
                xin, k = self.halo_buffer(key + ".t", nt, H, W, C)
                xin[k:].copy_(src)                               # stage A / B tensors: a small copy
                self.commit(key + ".t", xin)
                mv.vae_conv(xin, p["time"], y, res=None, o_base=fe if first else 0, os_t=2 * fe, os_h=W * C, os_w=C,
                            nsplit=C, nsplit_off=fe, t_off=k)
            x = y
            T = T2
        Co = C // 2
        out = torch.empty(T, 2 * H, 2 * W, Co, dtype=F16, device=self.device)
        a_next = None
        if nxt is not None and all(self._fusable(c) for c in p["par"].values()):
            an, kn = self.halo_buffer(nxt[1], T, 2 * H, 2 * W, Co)
            a_next = (an, kn)
        for (a, b), c in p["par"].items():
            kw = dict(o_base=(a * 2 * W + b) * Co, os_t=4 * H * W * Co, os_h=4 * W * Co, os_w=2 * Co)
            if a_next is not None:
                mv.vae_conv_fused(x, c, out, nxt[0], a_next[0][a_next[1]:], **kw)
            else:
                mv.vae_conv(x, c, out, res=None, **kw)
        return out, a_next

    # -- WanVAE.decode for one latent ---------------------------------------------------------------------------
    def decode(self, z):
        """z [z_dim, T, h, w] fp32 -> [3, 1+4(T-1), 8h, 8w] fp32 in [-1, 1].

        MOVII_VAE_GRAPH=1: the ~570 launches + ~150 device copies of one decode are captured in a CUDA graph per (latent
        shape, chunking) on the second decode of that shape and replayed afterwards: same kernels, same order, no host
        work between them.  Default: direct launches (measured equally fast)."""
        if not self.use_graph or mv.timing_enabled():
            return self._decode(z)
        key = (tuple(z.shape), self.chunk, self.head_mode, self.fuse_c6, mv.CONFIG_EPOCH)
        ent = self._graphs.get(key)
        if ent is None:                                    # first decode of this shape: direct launches (also warms up)
            self._graphs[key] = "seen"
            return self._decode(z)
        if ent == "seen":
            try:
                static_z = z.detach().to(self.device, F32).contiguous().clone()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0 = mv.LAUNCHES
                with torch.cuda.graph(graph):
                    static_out = self._decode(static_z)
                ent = (graph, static_z, static_out, mv.LAUNCHES - l0)
                self._graphs[key] = ent
            except Exception as ex:                           # capture refused: keep launching directly (same kernels)
                logging.warning("WanVAE.decode: CUDA graph capture failed (%r); using direct launches", ex)
                self.use_graph = False
                self.cache = {}
                return self._decode(z)
        graph, static_z, static_out, n_launch = ent
        static_z.copy_(z)
        graph.replay()
        mv.count_replayed_launches(n_launch)
        return static_out.clone()

    def _decode(self, z):
        Z, T, h, w = z.shape
        z = z.to(self.device, F32).contiguous()
        n_up3d = sum(1 for kind, p in self.layers if kind == "up" and p["mode"] == "upsample3d")
        sp_up = sum(1 for kind, p in self.layers if kind == "up")
        Tout = T
        for _ in range(n_up3d):
            Tout = 2 * Tout - 1
        Ho, Wo = h << sp_up, w << sp_up
        video = torch.empty(3, Tout, Ho, Wo, dtype=F32, device=self.device)
        self.cache = {}
        t_done = 0
        try:
            for c0 in range(0, T, self.chunk):
                n = min(self.chunk, T - c0)
                first = c0 == 0
                zc = z[:, c0:c0 + n].contiguous()
                x, k = self.halo_buffer("conv1", n, h, w, Z)
                mv.vae_latent_in(zc, self.w2, self.b2, self.mean, self.std, x[k:])
                self.commit("conv1", x)
                x = self.conv(x, self.conv1, t_off=k)
                a_in = None          # silu(rms_norm(x)) already written into the next block's first conv input
                nl = len(self.layers)
                for i, (kind, p) in enumerate(self.layers):
                    if i + 1 < nl:
                        nk, npar = self.layers[i + 1]
                        nxt = (npar["g0"], "L%d.c2" % (i + 1)) if nk == "res" else None
                    else:
                        nxt = (self.head_gamma, "head")
                    if kind == "res":
                        x, a_in = self.resblock(x, p, "L%d" % i, a_in=a_in, nxt=nxt)
                    elif kind == "attn":
                        x, a_in = self.attention(x, p), None
                    else:
                        x, a_in = self.upsample(x, p, "L%d" % i, first, nxt=nxt)
                nt, H, W, C = x.shape
                if a_in is None:
                    a, k = self.halo_buffer("head", nt, H, W, C)
                    self.normsilu(x, self.head_gamma, out=a[k:])
                else:
                    a, k = a_in
                del x, a_in
                if self.head_mode == "gather":
                    assert k == 0                                   # the temporal history lives in D, not in the input
                    D = self.conv(a, self.head_taps)                # [nt, H, W, 112] partial sums per (tap, channel)
                    del a
                    prev = self.cache.get("head.D")
                    mv.vae_head_gather(D, prev, self.head_bias, video, t_done)
                    if nt >= 2 or prev is None:
                        self.cache["head.D"] = D[-2:].clone()
                    else:
                        self.cache["head.D"] = torch.cat([prev[-1:], D])
                    del D, prev
                else:
                    self.commit("head", a)
                    mv.vae_conv(a, self.head, video, res=None, out_mode=1, o_base=t_done * H * W, os_t=Tout * H * W,
                                t_off=k)
                    del a
                t_done += nt
        finally:
            self.cache = {}
        assert t_done == Tout
        return video


    # -- WanVAE.encode ---------------------------------------------------------------------------
    def _pack_encoder(self):
        m, dev = self.model, self.device
        e = m.encoder
        f32 = lambda t: t.detach().to(F32).contiguous().to(dev)  # noqa: E731
        w1 = e.conv1.weight.detach()
        w1 = torch.cat([w1, w1.new_zeros(w1.shape[0], 16 - w1.shape[1], *w1.shape[2:])], dim=1)   # Cin 3 -> 16 (zeros)
        layers = []
        for mod in list(e.downsamples) + list(e.middle):
            if isinstance(mod, _Res):
                layers.append(("res", self._res(mod)))
            elif isinstance(mod, _Attn):
                layers.append(("attn", dict(
                    gamma=f32(mod.norm.gamma.reshape(-1)),
                    qkv=_Conv(mod.to_qkv.weight.unsqueeze(2), mod.to_qkv.bias, _taps(1, 1, 1), dev),
                    proj=_Conv(mod.proj.weight.unsqueeze(2), mod.proj.bias, _taps(1, 1, 1), dev))))
            else:
                # ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2): taps at (2h + kh, 2w + kw), zeros past the far edge
                down = dict(mode=mod.mode, sp=_Conv(mod.resample[1].weight.unsqueeze(2), mod.resample[1].bias,
                                                    [(0, kh, kw) for kh in range(3) for kw in range(3)], dev))
                if mod.mode == "downsample3d":
                    down["time"] = _Conv(mod.time_conv.weight, mod.time_conv.bias, _taps(3, 1, 1), dev)
                layers.append(("down", down))
        zz = 2 * self.z_dim
        self.enc = dict(conv1=_Conv(w1, e.conv1.bias, _taps(3, 3, 3), dev), layers=layers,
                        head_gamma=f32(e.head[0].gamma.reshape(-1)),
                        head=_Conv(e.head[2].weight, e.head[2].bias, _taps(3, 3, 3), dev),
                        w1=f32(m.conv1.weight.reshape(zz, zz)), b1=f32(m.conv1.bias),
                        inv_std=(1.0 / self.std).contiguous())
        # input frames per pass after the first frame (a multiple of 4: the two stride-2 time convs stay aligned)
        self.enc_chunk = max(4, int(os.environ.get("MOVII_VAE_ENC_CHUNK", "8")) // 4 * 4)

    def downsample(self, x, p, key, first):
        """Resample 'downsample2d' / 'downsample3d' (vae.py:92-104,140-159)."""
        T, H, W, C = x.shape
        Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
        if p["mode"] == "downsample2d":
            return mv.vae_conv_strided(x, p["sp"], torch.empty(T, Ho, Wo, C, dtype=F16, device=self.device), (1, 2, 2))
        # time_conv (3,1,1) stride (2,1,1), no temporal padding, over [last frame of the previous chunk | chunk]; the
        # first chunk (frame 0 alone) only fills the cache and passes through (:146-148)
        if first:
            assert T == 1
            y = mv.vae_conv_strided(x, p["sp"], torch.empty(1, Ho, Wo, C, dtype=F16, device=self.device), (1, 2, 2))
            self.cache[key + ".t"] = y.clone()
            return y
        assert T % 2 == 0
        xin = torch.empty(1 + T, Ho, Wo, C, dtype=F16, device=self.device)
        xin[:1].copy_(self.cache[key + ".t"])
        mv.vae_conv_strided(x, p["sp"], xin[1:], (1, 2, 2))
        self.cache[key + ".t"] = xin[-1:].clone()
        return mv.vae_conv_strided(xin, p["time"], torch.empty(T // 2, Ho, Wo, C, dtype=F16, device=self.device),
                                   (2, 1, 1), t_off=2)

    def encode(self, video):
        """video [3, 1+4k, H, W] fp32 in [-1, 1] -> normalised mu [z_dim, 1+k, H/8, W/8] fp32 (vae.py:516-542, 650-655)."""
        if self.enc is None:
            self._pack_encoder()
        E = self.enc
        _, T, H, W = video.shape
        if H % 8 or W % 8:
            raise ValueError("WanVAE.encode: H and W must be multiples of 8, got %dx%d" % (H, W))
        T = 1 + (T - 1) // 4 * 4                         # the reference drops frames that do not fill a chunk (:521-530)
        video = video.to(self.device, F32).contiguous()
        Tl = 1 + (T - 1) // 4
        mu = torch.empty(self.z_dim, Tl, H // 8, W // 8, dtype=F32, device=self.device)
        self.cache = {}
        t_done = 0
        try:
            c0 = 0
            while c0 < T:
                first = c0 == 0
                n = 1 if first else min(self.enc_chunk, T - c0)
                x, k = self.halo_buffer("E.conv1", n, H, W, 16)
                mv.vae_video_in(video, c0, n, x[k:])
                self.commit("E.conv1", x)
                x = self.conv(x, E["conv1"], t_off=k)
                a_in = None
                layers = E["layers"]
                for i, (kind, p) in enumerate(layers):
                    nxt = None
                    if i + 1 < len(layers):
                        if layers[i + 1][0] == "res":
                            nxt = (layers[i + 1][1]["g0"], "E%d.c2" % (i + 1))
                    else:
                        nxt = (E["head_gamma"], "E.head")
                    if kind == "res":
                        x, a_in = self.resblock(x, p, "E%d" % i, a_in=a_in, nxt=nxt)
                    elif kind == "attn":
                        x, a_in = self.attention(x, p), None
                    else:
                        x, a_in = self.downsample(x, p, "E%d" % i, first), None
                nt, h, w, C = x.shape
                if a_in is None:
                    a, k = self.halo_buffer("E.head", nt, h, w, C)
                    self.normsilu(x, E["head_gamma"], out=a[k:])
                else:
                    a, k = a_in
                del x, a_in
                self.commit("E.head", a)
                hd = self.conv(a, E["head"], t_off=k)
                mv.vae_latent_out(hd, E["w1"], E["b1"], self.mean, E["inv_std"], mu, t_done)
                t_done += nt
                c0 += n
        finally:
            self.cache = {}
        assert t_done == Tl
        return mu


class WanVAE:
    """wan/modules/vae.py:619-663.  vae_pth=None (or a missing file) -> random-init weights (benchmarks)."""

    def __init__(self, z_dim=16, vae_pth=None, dtype=torch.float, device="cuda"):
        self.dtype, self.device = dtype, device
        self.mean = torch.tensor(VAE_MEAN, dtype=dtype, device=device)
        self.std = torch.tensor(VAE_STD, dtype=dtype, device=device)
        self.scale = [self.mean, 1.0 / self.std]
        self.model = WanVAE_(dim=96, z_dim=z_dim, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                             temperal_downsample=(False, True, True))
        if vae_pth:
            logging.info("loading %s", vae_pth)
            self.model.load_state_dict(torch.load(vae_pth, map_location="cpu", weights_only=True))
        else:
            nn.init.normal_(self.model.decoder.middle[1].proj.weight, std=0.02)  # zero-init would hide the attention
            nn.init.normal_(self.model.encoder.middle[1].proj.weight, std=0.02)
        self.model.eval().requires_grad_(False).to(device)

    def encode(self, videos):
        """videos: list of [3, 1+4k, H, W] in [-1, 1] -> list of normalised latents mu [16, 1+k, H/8, W/8] fp32 (:650-655)."""
        eng = self.model.engine()
        return [eng.encode(u) for u in videos]

    def decode(self, zs):
        """zs: list of [16, T, h, w] latents -> list of [3, 1+4(T-1), 8h, 8w] fp32 in [-1, 1]."""
        eng = self.model.engine()
        return [eng.decode(u) for u in zs]
