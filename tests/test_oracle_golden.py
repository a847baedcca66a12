"""CPU: the oracle (oracle/dit_oracle.py) and the host-side restatements are pinned against golden vectors minted
from the reference's own code by oracle/make_golden.py (the reference has no tests of its own)."""
import os

import pytest
import torch

from oracle import dit_oracle as O
from oracle.fill import state_dict_like

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_oracle_block_config1_matches_reference():
    """BASELINE.json configs[0]: one WanAttentionBlock, dim 128, 256 tokens, fp32, CPU."""
    g = load("block_cfg1.pt")
    sd = {"blk." + k: v for k, v in state_dict_like(g["param_shapes"], g["seed"]).items()}
    angles = O.rope_table(g["grid"], 32, 256)
    y = O.block_forward(sd, "blk.", g["x"][0], g["e"][0], angles, g["ctx"][0], 4, 1e-6, O.ident, k_len=256)
    # fp32 vs fp32 (different summation orders only): 1e-5 relative
    assert ((y - g["y"][0]).norm() / g["y"][0].norm()).item() < 1e-5
    assert (y - g["y"][0]).abs().max().item() < 1e-4


@pytest.mark.parametrize("seq_len", [32, 40])
def test_oracle_tiny_model_matches_reference(seq_len):
    g = load("model_tiny_hd128.pt")
    sd = state_dict_like(g["param_shapes"], g["seed"])
    y = O.model_forward(sd, g["cfg"], g["x"], g["t"][0], g["ctx"], seq_len, O.ident)
    ref = g["y%d" % seq_len]
    assert y.shape == ref.shape
    assert ((y - ref).norm() / ref.norm()).item() < 1e-5


def test_bf16_emulation_is_close_to_fp32_reference():
    """The bf16 cast points the kernels implement stay within the §8c block tolerance of the fp32 reference."""
    g = load("model_tiny_hd128.pt")
    sd = state_dict_like(g["param_shapes"], g["seed"])
    y = O.model_forward(sd, g["cfg"], g["x"], g["t"][0], g["ctx"], 40, O.bf16_rt)
    rel = ((y - g["y40"]).norm() / g["y40"].norm()).item()
    assert rel < 1e-2, rel


def test_wan_model_state_dict_matches_reference_names():
    """Reference checkpoints must load: same parameter names and shapes."""
    from wan.modules.model import WanModel
    g = load("model_tiny_hd128.pt")
    cfg = dict(g["cfg"])
    m = WanModel(**cfg)
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ours == g["param_shapes"]


def test_unipc_scheduler_matches_reference():
    from wan.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    gold = load("unipc.pt")
    for steps, rec in gold.items():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        s.set_timesteps(steps, device="cpu", shift=5.0)
        assert torch.equal(s.timesteps, rec["timesteps"])
        assert torch.equal(s.sigmas, rec["sigmas"])
        x = rec["traj"][0]
        for i, t in enumerate(s.timesteps):
            x = s.step(rec["model_outputs"][i], t, x, return_dict=False)[0]
            ref = rec["traj"][i + 1]
            assert (x - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), (steps, i)


def test_rope_table_matches_engine_table():
    """The engine's fp32 cos/sin table (built from the complex128 freqs like the reference) equals the oracle angles."""
    from wan.modules.engine import rope_cos_sin
    from wan.modules.model import rope_params
    d = 128
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1)
    grid, seq_len = (3, 4, 5), 64
    cs = rope_cos_sin(freqs, grid, seq_len)
    ang = O.rope_table(grid, d, seq_len)
    assert cs.shape == (seq_len, d // 2, 2)
    assert (cs[..., 0].double() - torch.cos(ang)).abs().max() < 1e-6
    assert (cs[..., 1].double() - torch.sin(ang)).abs().max() < 1e-6
    part = rope_cos_sin(freqs, grid, seq_len, start=16, rows=16)
    assert torch.equal(part, cs[16:32])


@pytest.mark.parametrize("case", [(1, 4, 6), (2, 5, 9), (3, 4, 6), (5, 4, 4)])
def test_vae_oracle_matches_reference_chunked_decode(case):
    """The whole-sequence restatement equals the reference's chunked decode with feature cache (incl. the 'Rep' path)."""
    from oracle import vae_oracle as V
    g = load("vae_decode.pt")
    sd = state_dict_like(g["param_shapes"], g["seed"])
    rec = g["cases"][case]
    y = V.decode(sd, rec["z"])
    ref = rec["y"].float()
    assert y.shape == ref.shape == (3, 1 + 4 * (case[0] - 1), 8 * case[1], 8 * case[2])
    assert (y - ref).abs().max().item() <= 1e-3          # golden stored in fp16 (4.9e-4 quantisation on [-1, 1])


def test_vae_state_dict_matches_reference_names():
    from wan.modules.vae import WanVAE_
    m = WanVAE_()
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {**load("vae_decode.pt")["param_shapes"], **load("vae_encode.pt")["param_shapes"]}   # decoder+conv2, encoder+conv1
    assert ours == ref


@pytest.mark.parametrize("case", [(1, 16, 24), (5, 24, 40), (9, 16, 16), (13, 32, 16)])
def test_vae_oracle_matches_reference_chunked_encode(case):
    """The whole-sequence encoder restatement equals the reference's 1 + 4 + 4 ... chunked encode with feature cache
    (incl. downsample3d's pass-through of frame 0 and its one-frame cache); the fp16 storage contract of the CUDA path
    stays within 2e-3 of it on latents of magnitude ~1.5."""
    from oracle import vae_oracle as V
    g = load("vae_encode.pt")
    sd = state_dict_like(g["param_shapes"], g["seed"])
    rec = g["cases"][case]
    x = rec["x"].float()
    mu = V.encode(sd, x)
    T, H, W = case
    assert mu.shape == rec["mu"].shape == (16, 1 + (T - 1) // 4, H // 8, W // 8)
    assert (mu - rec["mu"]).abs().max().item() <= 1e-5
    assert (V.encode(sd, x, V.f16_rt) - rec["mu"]).abs().max().item() <= 2e-3


def test_dpmpp_scheduler_matches_reference():
    """sample_solver='dpm++' (text2video.py:214-223): sigma schedule + 2M midpoint trajectory of the reference."""
    from wan.utils.fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas, retrieve_timesteps
    gold = load("dpmpp.pt")
    for steps, rec in gold.items():
        s = FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        timesteps, n = retrieve_timesteps(s, device="cpu", sigmas=get_sampling_sigmas(steps, 5.0))
        assert n == steps and torch.equal(timesteps, rec["timesteps"]) and torch.equal(s.sigmas, rec["sigmas"])
        x = rec["traj"][0]
        for i, t in enumerate(timesteps):
            x = s.step(rec["model_outputs"][i], t, x, return_dict=False)[0]
            ref = rec["traj"][i + 1]
            assert (x - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), (steps, i)


# ---- umT5 text encoder ---------------------------------------------------------------------------
def _t5_cases():
    g = load("t5_encoder.pt")
    return g, sorted(g["cases"])


@pytest.mark.parametrize("case", [(24, 17), (300, 263), (40, 40)])
def test_t5_oracle_matches_reference(case):
    """oracle/t5_oracle.py against the reference's own T5Encoder (fp32, CPU): tiny umT5-style config, padded and
    unpadded prompts, relative distances past max_dist."""
    from oracle import t5_oracle as T
    from oracle.fill import t5_state_dict_like
    g = load("t5_encoder.pt")
    sd = t5_state_dict_like(g["param_shapes"], g["seed"])
    c = g["cases"][case]
    cfg = g["cfg"]
    y = T.t5_encoder_forward(sd, c["ids"], c["mask"], cfg["num_heads"], cfg["num_buckets"], cfg["shared_pos"])
    assert ((y - c["y"]).norm() / c["y"].norm()).item() < 1e-5
    # the CUDA path's arithmetic contract (bf16 GEMM I/O) stays within 1.5e-2 of the fp32 reference
    yb = T.t5_encoder_forward(sd, c["ids"], c["mask"], cfg["num_heads"], cfg["num_buckets"], cfg["shared_pos"],
                              rb=T.bf16_rt)
    assert ((yb - c["y"]).norm() / c["y"].norm()).item() < 1.5e-2
    # valid rows do not depend on the padded tail: encoding only the prefix gives the same rows (what
    # T5EncoderModel.__call__ relies on, t5.py:517)
    n = case[1]
    yp = T.t5_encoder_forward(sd, c["ids"][:n], None, cfg["num_heads"], cfg["num_buckets"], cfg["shared_pos"])
    assert ((yp - c["y"][:n]).norm() / c["y"][:n].norm()).item() < 1e-5


def test_t5_encoder_state_dict_and_bias_table_match_reference():
    """Reference T5 checkpoints must load (same names / shapes), and the per-distance bias table the kernel indexes
    equals the reference's materialised [N, Lq, Lk] bias."""
    from oracle import t5_oracle as T
    from wan.modules.t5 import T5Encoder, T5RelativeEmbedding
    g = load("t5_encoder.pt")
    m = T5Encoder(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == g["param_shapes"]
    rel = T5RelativeEmbedding(32, 4, bidirectional=True)
    torch.nn.init.normal_(rel.embedding.weight)
    for lq, lk in ((7, 7), (300, 300), (512, 512), (5, 9)):
        tab = rel.table(lq, lk)                                  # [N, lq + lk - 1]
        full = T.relative_bias(rel.embedding.weight.detach().float(), lq, lk, 32)
        i = torch.arange(lq).view(-1, 1)
        j = torch.arange(lk).view(1, -1)
        assert torch.equal(tab[:, (j - i) + (lq - 1)], full)


def test_tokenizer_cleaning_matches_reference_semantics():
    from wan.modules import tokenizers as tk
    assert tk.whitespace_clean("  a \n\t b  ") == "a b"
    assert tk.basic_clean(" &amp;amp; x ") == "& x"
    assert tk.canonicalize("Hello_World, it's  ME!") == "hello world its me"
    assert tk.canonicalize("a.b|c.d", keep_punctuation_exact_string="|") == "ab|cd"


def test_vae_head_conv_equals_tap_matrix_plus_gather():
    """Host logic of the decoder head (wan/modules/vae.py::_head_tap_matrix + csrc vae_head_gather_kernel): the causal
    3x3x3 conv 96 -> 3 equals ONE per-voxel matrix product to 27 x (3 + 1 pad) partial sums followed by the sum of each
    output voxel's 27 neighbours' partials (zero outside the image and before the first frame)."""
    import torch.nn.functional as F
    from oracle import vae_oracle as V
    from wan.modules.vae import _head_tap_matrix
    g = torch.Generator().manual_seed(21)
    C, T, H, W = 96, 4, 5, 7
    w, b = torch.randn(3, C, 3, 3, 3, generator=g) / 50, torch.randn(3, generator=g)
    x = torch.randn(1, C, T, H, W, generator=g)
    ref = V.causal_conv3d(x, w, b)[0]                                            # [3, T, H, W]
    wt = _head_tap_matrix(w)
    assert wt.shape == (112, C) and torch.count_nonzero(wt[3::4]) == 0 and torch.count_nonzero(wt[108:]) == 0
    D = x[0].permute(1, 2, 3, 0) @ wt.t()                                        # [T, H, W, 112]
    Dp = F.pad(D, (0, 0, 1, 1, 1, 1, 2, 0))                                      # two zero frames in front, zero ring
    out = torch.zeros(3, T, H, W)
    for it in range(3):
        for ih in range(3):
            for iw in range(3):
                tap = (it * 3 + ih) * 3 + iw
                out += Dp[it:it + T, ih:ih + H, iw:iw + W, 4 * tap:4 * tap + 3].permute(3, 0, 1, 2)
    out += b.view(3, 1, 1, 1)
    assert (out - ref).abs().max().item() <= 1e-5


def test_vae_subpixel_upsample_weights_equal_upsample_then_conv():
    """Host logic of Resample 'upsample2d/3d' (wan/modules/vae.py::_parity_weights): nearest-exact x2 + Conv2d(3x3, pad 1)
    equals four 2x2 convolutions on the low-resolution image, one per output parity, with the kernel rows / columns that
    land on the same source pixel pre-summed."""
    import torch.nn.functional as F
    from wan.modules.vae import _parity_weights
    g = torch.Generator().manual_seed(22)
    Ci, Co, H, W = 8, 4, 5, 7
    w, b = torch.randn(Co, Ci, 3, 3, generator=g), torch.randn(Co, generator=g)
    x = torch.randn(2, Ci, H, W, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=(2.0, 2.0), mode="nearest-exact"), w, b, padding=1)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for (a, bb), (taps, wp) in _parity_weights(w).items():
        acc = b.view(1, Co, 1, 1).expand(2, Co, H, W).clone()
        for i, (_, dh, dw) in enumerate(taps):
            acc += torch.einsum("nchw,oc->nohw", xp[:, :, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W], wp[:, :, i, 0, 0])
        out[:, :, a::2, bb::2] = acc
    assert (out - ref).abs().max().item() <= 1e-5

