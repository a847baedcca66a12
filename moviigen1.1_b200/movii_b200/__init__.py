"""ctypes binding of libmovii_b200.so (include/movii_b200.h).

PyTorch here is plumbing only: it owns device memory and streams; every operator below is a
hand-written sm_100a kernel reached through the C ABI.  There is no fallback: if the library is
missing or the device is not a B200 the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "lib", "libmovii_b200.so")

MV_EPI_BF16 = 0
MV_EPI_BF16_GELU = 1
MV_EPI_RESID_F32 = 2
MV_EPI_F32_ROUND = 3
MV_EPI_F32 = 4

_lib = None
_c = ctypes
_i64 = _c.c_int64
_int = _c.c_int
_f32 = _c.c_float
_ptr = _c.c_void_p

# name -> argtypes (restype is always int); mirrors include/movii_b200.h one to one
_SIGNATURES = {
    "mv_gemm_bf16": [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _int, _int, _int, _int, _ptr],
    "mv_gemm_f16": [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _int, _int, _int, _int, _ptr],
    "mv_attention_fwd": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _int, _int, _int, _f32, _ptr],
    "mv_ln_modulate": [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _int, _int, _f32, _int, _ptr],
    "mv_rmsnorm_rope": [_ptr, _i64, _ptr, _ptr, _int, _int, _int, _f32, _ptr],
    "mv_gemm_bf16_ksplit": [_ptr, _i64, _i64, _int, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _int, _int, _int, _int, _ptr],
    "mv_qkv_prepare": [_ptr, _i64, _ptr, _ptr, _ptr, _int, _int, _int, _int, _f32, _ptr],
    "mv_qkv_norm_rope": [_ptr, _i64, _ptr, _ptr, _ptr, _int, _int, _int, _f32, _ptr, _ptr, _ptr, _int, _int, _ptr],
    "mv_unipc_cfg_step": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _int, _ptr],
    "mv_modulation_table": [_ptr, _ptr, _ptr, _int, _int, _ptr],
    "mv_head_tokens": [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _int, _int, _f32, _ptr],
    "mv_unpatchify": [_ptr, _ptr, _int, _int, _int, _int, _int, _int, _ptr],
    "mv_vae_conv": [_ptr, _int, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _int, _int, _int, _int, _int, _int, _int,
                    _ptr, _i64, _i64, _i64, _i64, _int, _i64, _int, _ptr],
    "mv_vae_conv_fused": [_ptr, _int, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _int, _int, _int, _int, _int, _ptr, _i64,
                          _i64, _i64, _i64, _ptr, _ptr, _int, _ptr],
    "mv_vae_rmsnorm_silu": [_ptr, _ptr, _ptr, _i64, _int, _int, _ptr],
    "mv_vae_latent_in": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _i64, _ptr],
    "mv_softmax_rows": [_ptr, _i64, _ptr, _i64, _int, _int, _f32, _ptr],
    "mv_vae_conv_strided": [_ptr, _int, _int, _int, _int, _ptr, _ptr, _ptr, _int, _int, _int, _int, _int, _ptr, _int,
                            _int, _int, _int, _ptr],
    "mv_vae_video_in": [_ptr, _int, _int, _int, _int, _int, _ptr, _ptr],
    "mv_vae_head_gather": [_ptr, _ptr, _int, _int, _int, _int, _ptr, _ptr, _i64, _i64, _ptr],
    "mv_vae_latent_out": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _i64, _i64, _i64, _ptr],
    "mv_ipc_export": [_ptr, _ptr, _ptr, _ptr],
    "mv_ipc_open": [_ptr, _ptr],
    "mv_ipc_close": [_ptr],
    "mv_sp_barrier": [_ptr, _ptr, _int, _int, _c.c_uint, _ptr],
    "mv_qkv_prepare_p2p": [_ptr, _i64, _ptr, _ptr, _ptr, _int, _int, _int, _int, _int, _f32, _ptr],
    "mv_attention_fwd_scatter": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _int, _int, _int, _i64, _int, _int, _int,
                                 _f32, _ptr],
    "mv_patchify": [_ptr, _ptr, _int, _int, _int, _int, _int, _int, _ptr],
    "mv_head_unpatchify": [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _int, _int, _int, _int, _int, _int,
                           _f32, _ptr],
    "mv_linear_f32_vec": [_ptr, _ptr, _ptr, _ptr, _int, _int, _int, _ptr],
    "mv_sinusoid_embed": [_ptr, _int, _ptr, _int, _ptr],
    "mv_attention_fwd_trace": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _int, _int, _int, _f32, _ptr, _int, _ptr],
    "mv_attention_config": [_int, _int, _int, _int, _int, _int, _int],
    "mv_gemm_config": [_int],
    "mv_roles_config": [_int],
    "mv_vae_conv_config": [_int, _int, _int],
    "mv_t5_attention": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _int, _int, _int, _int, _int, _ptr],
    "mv_t5_rmsnorm": [_ptr, _i64, _ptr, _ptr, _i64, _int, _int, _f32, _ptr],
    "mv_embed_gather": [_ptr, _i64, _i64, _ptr, _ptr, _i64, _int, _int, _ptr],
    "mv_mul_bf16": [_ptr, _i64, _ptr, _i64, _ptr, _i64, _int, _int, _ptr],
}
EXPORTED_SYMBOLS = ["mv_last_error", "mv_version", "mv_device_check"] + sorted(_SIGNATURES)


def lib():
    """Loads the shared library (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is None:
        path = os.path.abspath(LIB_PATH)
        if not os.path.exists(path):
            raise RuntimeError(
                "libmovii_b200.so not found at %s — run `python moviigen1.1_b200/build.py` "
                "(there is no non-CUDA fallback)" % path)
        _lib = ctypes.CDLL(path)
        _lib.mv_last_error.restype = ctypes.c_char_p
        _lib.mv_version.restype = _int
        _lib.mv_device_check.restype = _int
        for name, args in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = _int
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().mv_last_error().decode()))


LAUNCHES = 0          # kernels launched through the C ABI by this process (bench.py reports it)
CONFIG_EPOCH = 0      # bumped by every *_config() call: captured CUDA graphs are keyed on it (a graph bakes the variant in)
_TIMED = {}           # entry point -> list of (start_event, end_event); filled while enabled via time_kernels()


def time_kernels(names):
    """Enable (list of entry-point names) or disable (None) per-launch CUDA-event timing on the launching stream."""
    _TIMED.clear()
    for n in names or ():
        _TIMED[n] = []
    return _TIMED


def _call(name, *args):
    global LAUNCHES
    rec = _TIMED.get(name)
    if rec is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib(), name)(*args)
        e.record()
        rec.append((s, e, args))
    else:
        rc = getattr(lib(), name)(*args)
    LAUNCHES += 1
    _check(rc, name)


def count_replayed_launches(n):
    """A CUDA-graph replay launches the n kernels that were captured through _call(): keep LAUNCHES truthful."""
    global LAUNCHES
    LAUNCHES += int(n)


def timing_enabled():
    return bool(_TIMED)


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (movii_b200 has no CPU path)" % name)
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream: a tensor on another GPU would be a peer access or a fault
        raise RuntimeError("%s lives on %s but the current CUDA device is %d: call torch.cuda.set_device(%d) "
                           "(generate.py:191-200 does) before using movii_b200 on that GPU"
                           % (name, t.device, torch.cuda.current_device(), t.device.index))
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))


def device_check():
    _check(lib().mv_device_check(), "mv_device_check")


def gemm(a, w, bias, out, epilogue, gate=None):
    """out = epilogue(a @ w.T + bias).  a [M,K] bf16 (row stride any multiple of 8), w [N,K] bf16."""
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w"); _req(bias, torch.float32, "bias")
    _req(gate, torch.float32, "gate")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.shape[0] == M and out.shape[1] == N
    want = torch.float32 if epilogue in (MV_EPI_RESID_F32, MV_EPI_F32_ROUND, MV_EPI_F32) else torch.bfloat16
    _req(out, want, "out")
    _call("mv_gemm_bf16", _p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), _p(gate),
                              M, N, K, epilogue, _stream())
    return out


def gemm_f16(a, w, bias, out, epilogue):
    """gemm() with fp16 operands; out fp16 (MV_EPI_BF16 slot) or fp32 (MV_EPI_F32).  WanVAE attention."""
    _req(a, torch.float16, "a"); _req(w, torch.float16, "w"); _req(bias, torch.float32, "bias")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.shape[0] == M and out.shape[1] == N
    assert epilogue in (MV_EPI_BF16, MV_EPI_F32)
    _req(out, torch.float32 if epilogue == MV_EPI_F32 else torch.float16, "out")
    _call("mv_gemm_f16", _p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), None, M, N, K,
          epilogue, _stream())
    return out


def attention(q, k, v, out, softmax_scale=None):
    """q [Lq,H,128], k/v [Lk,H,128] bf16 views (last two dims contiguous), out [Lq,H,128] bf16."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1 and t.stride(1) == 128, n
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    assert v.shape[0] == Lk and k.shape[1] == H and v.shape[1] == H and out.shape[0] == Lq
    if softmax_scale is None:
        softmax_scale = 128 ** -0.5
    _call("mv_attention_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                                  Lq, Lk, H, float(softmax_scale), _stream())
    return out


def ln_modulate(x, out, shift=None, scale=None, weight=None, bias=None, eps=1e-6, round_ln=False):
    _req(x, torch.float32, "x"); _req(out, torch.bfloat16, "out")
    for t, n in ((shift, "shift"), (scale, "scale"), (weight, "weight"), (bias, "bias")):
        _req(t, torch.float32, n)
        assert t is None or t.is_contiguous()
    assert x.dim() == 2 and x.stride(1) == 1 and out.stride(1) == 1 and out.shape == x.shape
    M, C = x.shape
    _call("mv_ln_modulate", _p(x), x.stride(0), _p(shift), _p(scale), _p(weight), _p(bias), _p(out),
                                out.stride(0), M, C, float(eps), int(round_ln), _stream())
    return out


def rmsnorm_rope(x, weight, cs=None, head_dim=128, eps=1e-6):
    """In place on x [M,C] bf16 (row stride multiple of 8).  cs: [M, head_dim/2, 2] fp32 or None."""
    _req(x, torch.bfloat16, "x"); _req(weight, torch.float32, "weight"); _req(cs, torch.float32, "cs")
    assert x.dim() == 2 and x.stride(1) == 1 and weight.is_contiguous()
    M, C = x.shape
    if cs is not None:
        assert cs.is_contiguous() and cs.shape == (M, head_dim // 2, 2), (cs.shape, M, head_dim)
    _call("mv_rmsnorm_rope", _p(x), x.stride(0), _p(weight), _p(cs), M, C, head_dim, float(eps), _stream())
    return x


def qkv_norm_rope(qkv, gain_q, gain_k, cs=None, head_dim=128, eps=1e-6, dst=None, n_dst=0, src_slot=0):
    """q / k RMSNorm (+ RoPE) of a fused QKV buffer [M, 3C] bf16 in ONE launch.  dst=None: in place.
    dst=(tab_q, tab_k, tab_v) (ptr_table()s of n_dst slab bases): Ulysses head scatter of q, k and v, see
    include/movii_b200.h::mv_qkv_norm_rope."""
    _req(qkv, torch.bfloat16, "qkv"); _req(gain_q, torch.float32, "gain_q"); _req(gain_k, torch.float32, "gain_k")
    _req(cs, torch.float32, "cs")
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[1] % 3 == 0
    assert gain_q.is_contiguous() and gain_k.is_contiguous()
    M, C = qkv.shape[0], qkv.shape[1] // 3
    if cs is not None:
        assert cs.is_contiguous() and cs.shape == (M, head_dim // 2, 2), (cs.shape, M, head_dim)
    tq, tk, tv = dst if dst is not None else (None, None, None)
    assert (dst is None) == (n_dst == 0)
    _call("mv_qkv_norm_rope", _p(qkv), qkv.stride(0), _p(gain_q), _p(gain_k), _p(cs), M, C, head_dim, float(eps),
          tq, tk, tv, int(n_dst), int(src_slot), _stream())
    return qkv


UNIPC_NCOEF = 20


def unipc_cfg_step(cond, uncond, sample, last_sample, hist, coef, x0_out, sample_out, prev_out):
    """Classifier-free guidance + one UniPC update (include/movii_b200.h::mv_unipc_cfg_step).  hist: up to three
    fp32 tensors (model_outputs[-1], [-2], [-3] before this step) or None entries; coef: UNIPC_NCOEF python floats."""
    ts = [cond, uncond, sample, last_sample] + list(hist) + [x0_out, sample_out, prev_out]
    n = cond.numel()
    for t in ts:
        _req(t, torch.float32, "latent")
        assert t is None or (t.is_contiguous() and t.numel() == n)
    assert len(hist) == 3 and len(coef) == UNIPC_NCOEF
    arr = (_c.c_float * UNIPC_NCOEF)(*[float(v) for v in coef])
    _call("mv_unipc_cfg_step", _p(cond), _p(uncond), _p(sample), _p(last_sample), _p(hist[0]), _p(hist[1]), _p(hist[2]),
          _p(x0_out), _p(sample_out), _p(prev_out), n, arr, UNIPC_NCOEF, _stream())
    return prev_out


def modulation_table(mods, e0, out):
    """out[l] = mods[l] + e0 (fp32; mods/out [layers, len], e0 [len])."""
    _req(mods, torch.float32, "mods"); _req(e0, torch.float32, "e0"); _req(out, torch.float32, "out")
    assert mods.is_contiguous() and e0.is_contiguous() and out.is_contiguous() and out.shape == mods.shape
    layers = mods.shape[0]
    ln = mods.numel() // layers
    assert e0.numel() == ln
    _call("mv_modulation_table", _p(mods), _p(e0), _p(out), layers, ln, _stream())
    return out


def patchify(latent, out, patch_hw=(2, 2)):
    _req(latent, torch.float32, "latent"); _req(out, torch.bfloat16, "out")
    assert latent.is_contiguous() and out.is_contiguous() and latent.dim() == 4
    C, F, H, W = latent.shape
    ph, pw = patch_hw
    assert out.shape == (F * (H // ph) * (W // pw), C * ph * pw)
    _call("mv_patchify", _p(latent), _p(out), C, F, H, W, ph, pw, _stream())
    return out


def head_unpatchify(x, shift, scale, w, b, out, grid, patch_hw=(2, 2), eps=1e-6):
    """x [>=L, C] fp32; w [ph*pw*Cout, C] fp32; out [Cout, F, Hp*ph, Wp*pw] fp32."""
    for t, n in ((x, "x"), (shift, "shift"), (scale, "scale"), (w, "w"), (b, "b"), (out, "out")):
        _req(t, torch.float32, n)
    F, Hp, Wp = grid
    ph, pw = patch_hw
    C = x.shape[1]
    Cout = out.shape[0]
    assert w.is_contiguous() and w.shape == (ph * pw * Cout, C) and out.is_contiguous()
    assert out.shape == (Cout, F, Hp * ph, Wp * pw) and x.shape[0] >= F * Hp * Wp and x.stride(1) == 1
    _call("mv_head_unpatchify", _p(x), x.stride(0), _p(shift), _p(scale), _p(w), _p(b), _p(out), F, Hp, Wp, ph,
                                    pw, Cout, C, float(eps), _stream())
    return out


def linear_f32_vec(x, w, b, out, act_in=0):
    for t, n in ((x, "x"), (w, "w"), (b, "b"), (out, "out")):
        _req(t, torch.float32, n)
    N, K = w.shape
    assert w.is_contiguous() and x.numel() == K and out.numel() == N
    _call("mv_linear_f32_vec", _p(x), _p(w), _p(b), _p(out), N, K, int(act_in), _stream())
    return out


def sinusoid_embed(t, out):
    _req(out, torch.float32, "out")
    assert t.is_cuda and t.numel() == 1 and t.dtype in (torch.int64, torch.float32)
    _call("mv_sinusoid_embed", _p(t), int(t.dtype == torch.int64), _p(out), out.numel(), _stream())
    return out


def gemm_ksplit(a_blocks, w, bias, out, epilogue, gate=None):
    """Like gemm() with A given as [nblk, M, Kb] slabs (A[m, j*Kb + c] = a_blocks[j, m, c]); Kb % 64 == 0."""
    _req(a_blocks, torch.bfloat16, "a_blocks"); _req(w, torch.bfloat16, "w"); _req(bias, torch.float32, "bias")
    _req(gate, torch.float32, "gate")
    assert a_blocks.dim() == 3 and a_blocks.stride(2) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    nblk, M, Kb = a_blocks.shape
    N, K = w.shape
    assert K == nblk * Kb and out.shape == (M, N)
    want = torch.float32 if epilogue in (MV_EPI_RESID_F32, MV_EPI_F32_ROUND) else torch.bfloat16
    _req(out, want, "out")
    _call("mv_gemm_bf16_ksplit", _p(a_blocks), a_blocks.stride(1), a_blocks.stride(0), Kb, _p(w), w.stride(0),
                                     _p(bias), _p(out), out.stride(0), _p(gate), M, N, K, epilogue, _stream())
    return out


def qkv_prepare(x, weight, cs, out, sp_world, head_dim=128, eps=1e-6):
    """rmsnorm (if weight) + rope (if cs) of x [M,C] bf16, written head-scattered to out [sp_world, M, C/sp_world]."""
    _req(x, torch.bfloat16, "x"); _req(weight, torch.float32, "weight"); _req(cs, torch.float32, "cs")
    _req(out, torch.bfloat16, "out")
    assert x.dim() == 2 and x.stride(1) == 1 and out.is_contiguous()
    M, C = x.shape
    assert out.numel() == M * C
    if cs is not None:
        assert cs.is_contiguous() and cs.shape == (M, head_dim // 2, 2)
    _call("mv_qkv_prepare", _p(x), x.stride(0), _p(weight), _p(cs), _p(out), sp_world, M, C, head_dim, float(eps),
                                _stream())
    return out


def head_tokens(x, shift, scale, w, b, out, eps=1e-6):
    for t, n in ((x, "x"), (shift, "shift"), (scale, "scale"), (w, "w"), (b, "b"), (out, "out")):
        _req(t, torch.float32, n)
    L, C = x.shape
    nout = w.shape[0]
    assert w.is_contiguous() and out.is_contiguous() and out.shape == (L, nout) and x.stride(1) == 1
    _call("mv_head_tokens", _p(x), x.stride(0), _p(shift), _p(scale), _p(w), _p(b), _p(out), L, nout, C,
                                float(eps), _stream())
    return out


def unpatchify(tokens, out, grid, patch_hw=(2, 2)):
    _req(tokens, torch.float32, "tokens"); _req(out, torch.float32, "out")
    F, Hp, Wp = grid
    ph, pw = patch_hw
    Cout = out.shape[0]
    assert tokens.is_contiguous() and out.is_contiguous() and out.shape == (Cout, F, Hp * ph, Wp * pw)
    assert tokens.shape[0] >= F * Hp * Wp and tokens.shape[1] == ph * pw * Cout
    _call("mv_unpatchify", _p(tokens), _p(out), F, Hp, Wp, ph, pw, Cout, _stream())
    return out


def vae_conv(x, conv, out, res=None, o_base=0, os_t=0, os_h=0, os_w=0, nsplit=0, nsplit_off=0, out_mode=0, t_off=0):
    """x [t_off + T,H,W,Cin] fp16 channels-last (contiguous; the first t_off frames are the cached tail of the previous
    temporal chunk); conv: packed weights (.w [Cout,taps,Cin] fp16, .b fp32, .taps int8 CPU [ntaps,3]); out: fp16
    channels-last (mode 0, any size — addressing by o_base/os_*) or the fp32 video (mode 1, see movii_b200.h)."""
    _req(x, torch.float16, "x"); _req(conv.w, torch.float16, "w"); _req(conv.b, torch.float32, "bias")
    _req(res, torch.float16, "res")
    assert x.dim() == 4 and x.is_contiguous() and out.is_contiguous() and conv.w.is_contiguous()
    Tin, H, W, Cin = x.shape
    T = Tin - int(t_off)
    assert T > 0 and Cin == conv.cin and conv.taps.device.type == "cpu" and conv.taps.dtype == torch.int8
    _req(out, torch.float32 if out_mode == 1 else torch.float16, "out")
    if res is not None:
        assert res.is_contiguous()
    _call("mv_vae_conv", _p(x), Tin, H, W, Cin, _p(conv.w), _p(conv.b), _p(res), _p(out), int(out_mode), T, H, W,
          conv.cout, conv.cout_real, conv.ntaps, conv.taps.data_ptr(), int(o_base), int(os_t), int(os_h), int(os_w),
          int(nsplit), int(nsplit_off), int(t_off), _stream())
    return out


def vae_rmsnorm_silu(x, out, gamma, silu=True):
    _req(x, torch.float16, "x"); _req(out, torch.float16, "out"); _req(gamma, torch.float32, "gamma")
    assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape and gamma.numel() == x.shape[-1]
    C = x.shape[-1]
    _call("mv_vae_rmsnorm_silu", _p(x), _p(out), _p(gamma), x.numel() // C, C, int(bool(silu)), _stream())
    return out


def vae_latent_in(z, w2, b2, mean, std, out):
    for t, n in ((z, "z"), (w2, "w2"), (b2, "b2"), (mean, "mean"), (std, "std")):
        _req(t, torch.float32, n)
        assert t.is_contiguous()
    _req(out, torch.float16, "out")
    Z = z.shape[0]
    nvox = z.numel() // Z
    assert out.is_contiguous() and out.numel() == z.numel() and out.shape[-1] == Z
    _call("mv_vae_latent_in", _p(z), _p(w2), _p(b2), _p(mean), _p(std), _p(out), Z, nvox, _stream())
    return out


def vae_conv_strided(x, conv, out, stride=(1, 2, 2), t_off=0):
    """Strided conv (encoder downsampling): x [Tin,H,W,Cin] fp16 channels-last -> out [T,Ho,Wo,Cout] fp16 channels-last,
    out voxel (t,h,w) reads x[st*t + t_off + dt, sh*h + dh, sw*w + dw]; reads past the far edge are zeros."""
    _req(x, torch.float16, "x"); _req(conv.w, torch.float16, "w"); _req(conv.b, torch.float32, "bias")
    _req(out, torch.float16, "out")
    assert x.dim() == 4 and out.dim() == 4 and x.is_contiguous() and out.is_contiguous() and conv.w.is_contiguous()
    Tin, H, W, Cin = x.shape
    T, Ho, Wo, Co = out.shape
    assert Cin == conv.cin and Co == conv.cout and conv.taps.device.type == "cpu" and conv.taps.dtype == torch.int8
    st, sh, sw = (int(v) for v in stride)
    _call("mv_vae_conv_strided", _p(x), Tin, H, W, Cin, _p(conv.w), _p(conv.b), _p(out), T, Ho, Wo, Co, conv.ntaps,
          conv.taps.data_ptr(), int(t_off), st, sh, sw, _stream())
    return out


def vae_head_gather(d_cur, d_prev, bias3, video, t0):
    """d_cur [n,H,W,112] fp16 (+ d_prev [k<=2,H,W,112] or None) -> video[:, t0:t0+n] fp32 [3,T,H,W]; bias3: 3 python floats."""
    _req(d_cur, torch.float16, "d_cur"); _req(d_prev, torch.float16, "d_prev"); _req(video, torch.float32, "video")
    assert d_cur.dim() == 4 and d_cur.shape[3] == 112 and d_cur.is_contiguous() and video.is_contiguous()
    n, H, W, _ = d_cur.shape
    k = 0 if d_prev is None else d_prev.shape[0]
    assert d_prev is None or (d_prev.is_contiguous() and tuple(d_prev.shape[1:]) == (H, W, 112) and k <= 2)
    assert video.dim() == 4 and video.shape[0] == 3 and tuple(video.shape[2:]) == (H, W) and t0 + n <= video.shape[1]
    arr = (_c.c_float * 3)(*[float(v) for v in bias3])
    _call("mv_vae_head_gather", _p(d_cur), _p(d_prev), k, n, H, W, arr, _p(video), video.shape[1] * H * W, int(t0) * H * W,
          _stream())
    return video


def vae_video_in(video, t0, n, out):
    """video [3,T,H,W] fp32 (contiguous) frames [t0, t0+n) -> out [n,H,W,16] fp16 channels-last (channels 3.. zero)."""
    _req(video, torch.float32, "video"); _req(out, torch.float16, "out")
    assert video.dim() == 4 and video.shape[0] == 3 and video.is_contiguous() and out.is_contiguous()
    _, T, H, W = video.shape
    assert tuple(out.shape) == (n, H, W, 16)
    _call("mv_vae_video_in", _p(video), T, int(t0), int(n), H, W, _p(out), _stream())
    return out


def vae_latent_out(head, w1, b1, mean, inv_std, mu, t0):
    """head [n,h,w,2Z] fp16 -> mu[:, t0:t0+n] of the fp32 channel-first latent [Z,T,h,w] (conv1 + chunk + normalise)."""
    _req(head, torch.float16, "head"); _req(mu, torch.float32, "mu")
    for t, nme in ((w1, "w1"), (b1, "b1"), (mean, "mean"), (inv_std, "inv_std")):
        _req(t, torch.float32, nme)
        assert t.is_contiguous()
    assert head.is_contiguous() and mu.is_contiguous() and head.dim() == 4 and mu.dim() == 4
    n, h, w, C = head.shape
    Z, T = mu.shape[0], mu.shape[1]
    assert C == 2 * Z and tuple(mu.shape[2:]) == (h, w) and t0 + n <= T and tuple(w1.shape) == (C, C)
    _call("mv_vae_latent_out", _p(head), _p(w1), _p(b1), _p(mean), _p(inv_std), _p(mu), Z, n * h * w, T * h * w,
          int(t0) * h * w, _stream())
    return mu


def softmax_rows(s, p, n, scale):
    _req(s, torch.float32, "s"); _req(p, torch.float16, "p")
    assert s.dim() == 2 and p.dim() == 2 and s.stride(1) == 1 and p.stride(1) == 1 and s.shape[0] == p.shape[0]
    assert s.shape[1] >= n and p.shape[1] >= n
    _call("mv_softmax_rows", _p(s), s.stride(0), _p(p), p.stride(0), s.shape[0], n, float(scale), _stream())
    return p


# ---- NVLink peer-to-peer plumbing (fused Ulysses exchange) -------------------------------------------------------
def ptr_table(ptrs):
    """Host array of device pointers (void* const*) for the *_p2p / *_scatter entry points."""
    arr = (_c.c_void_p * 8)()
    for i, v in enumerate(ptrs):
        arr[i] = int(v)
    return arr


def ipc_export(t):
    """(handle: bytes[64], offset: int, alloc_bytes: int) for the cudaMalloc allocation that holds tensor t."""
    h = (_c.c_ubyte * 64)()
    off, size = _c.c_int64(0), _c.c_int64(0)
    _check(lib().mv_ipc_export(_c.c_void_p(t.data_ptr()), h, _c.byref(off), _c.byref(size)), "mv_ipc_export")
    return bytes(h), off.value, size.value


def ipc_open(handle):
    base = _c.c_void_p(0)
    buf = (_c.c_ubyte * 64).from_buffer_copy(handle)
    _check(lib().mv_ipc_open(buf, _c.byref(base)), "mv_ipc_open")
    return base.value


def ipc_close(base):
    _check(lib().mv_ipc_close(_c.c_void_p(base)), "mv_ipc_close")


def sp_barrier(peer_flag_table, local_flags_ptr, rank, world, epoch):
    _call("mv_sp_barrier", peer_flag_table, _c.c_void_p(local_flags_ptr), rank, world, _c.c_uint(epoch & 0xffffffff),
          _stream())


def qkv_prepare_p2p(x, weight, cs, dst_table, src_rank, sp_world, head_dim=128, eps=1e-6):
    _req(x, torch.bfloat16, "x"); _req(weight, torch.float32, "weight"); _req(cs, torch.float32, "cs")
    assert x.dim() == 2 and x.stride(1) == 1
    M, C = x.shape
    _call("mv_qkv_prepare_p2p", _p(x), x.stride(0), _p(weight), _p(cs), dst_table, src_rank, sp_world, M, C, head_dim,
          float(eps), _stream())


def attention_scatter(q, k, v, o_table, n_dst, src_rank, rows_per_rank, ldo, softmax_scale=None):
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1 and t.stride(1) == 128, n
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    if softmax_scale is None:
        softmax_scale = 128 ** -0.5
    _call("mv_attention_fwd_scatter", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), o_table, n_dst,
          src_rank, rows_per_rank, int(ldo), Lq, Lk, H, float(softmax_scale), _stream())


def vae_conv_fused(x, conv, out, gamma, norm_out, res=None, o_base=0, os_t=0, os_h=0, os_w=0, t_off=0):
    """vae_conv (fp16 channels-last) that also writes norm_out = silu(rms_norm(out) * gamma); out may be None."""
    _req(x, torch.float16, "x"); _req(conv.w, torch.float16, "w"); _req(conv.b, torch.float32, "bias")
    _req(res, torch.float16, "res"); _req(out, torch.float16, "out"); _req(norm_out, torch.float16, "norm_out")
    _req(gamma, torch.float32, "gamma")
    assert x.dim() == 4 and x.is_contiguous() and norm_out.is_contiguous() and conv.w.is_contiguous()
    assert out is None or out.is_contiguous()
    Tin, H, W, Cin = x.shape
    T = Tin - int(t_off)
    assert T > 0 and Cin == conv.cin and gamma.numel() == conv.cout and conv.cout == conv.cout_real
    _call("mv_vae_conv_fused", _p(x), Tin, H, W, Cin, _p(conv.w), _p(conv.b), _p(res), _p(out), T, H, W, conv.cout,
          conv.ntaps, conv.taps.data_ptr(), int(o_base), int(os_t), int(os_h), int(os_w), _p(gamma), _p(norm_out),
          int(t_off), _stream())
    return norm_out


# ---- umT5 text encoder (wan/modules/t5.py) -------------------------------------------------------------------------
def t5_attention(q, k, v, out, bias=None, bias_center=0, kv_len=None):
    """q [Lq,H,64], k/v [Lk,H,64] bf16 views, out [Lq,H,64] bf16; bias fp32 [H, n] indexed by (j - i) + bias_center;
    keys >= kv_len are masked out.  No softmax scale (T5)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
        assert t.dim() == 3 and t.shape[2] == 64 and t.stride(2) == 1 and t.stride(1) == 64, n
    _req(bias, torch.float32, "bias")
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    assert v.shape[0] == Lk and k.shape[1] == H and v.shape[1] == H and out.shape[0] == Lq
    if bias is not None:
        assert bias.dim() == 2 and bias.shape[0] == H and bias.stride(1) == 1
    _call("mv_t5_attention", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
          _p(bias), bias.stride(0) if bias is not None else 0, int(bias_center), Lq, Lk,
          Lk if kv_len is None else int(kv_len), H, _stream())
    return out


def t5_rmsnorm(x, weight, out, eps=1e-6):
    _req(x, torch.float32, "x"); _req(weight, torch.float32, "weight"); _req(out, torch.bfloat16, "out")
    assert x.dim() == 2 and out.shape == x.shape and x.stride(1) == 1 and out.stride(1) == 1
    assert weight.numel() == x.shape[1] and weight.is_contiguous()
    _call("mv_t5_rmsnorm", _p(x), x.stride(0), _p(weight), _p(out), out.stride(0), x.shape[0], x.shape[1], float(eps),
          _stream())
    return out


def embed_gather(table, ids, out):
    _req(table, torch.bfloat16, "table"); _req(ids, torch.int64, "ids"); _req(out, torch.float32, "out")
    assert table.dim() == 2 and table.stride(1) == 1 and ids.dim() == 1 and ids.is_contiguous()
    assert out.shape == (ids.numel(), table.shape[1]) and out.stride(1) == 1
    _call("mv_embed_gather", _p(table), table.stride(0), table.shape[0], _p(ids), _p(out), out.stride(0), ids.numel(),
          table.shape[1], _stream())
    return out


def mul_bf16(a, b, out):
    for t, n in ((a, "a"), (b, "b"), (out, "out")):
        _req(t, torch.bfloat16, n)
        assert t.dim() == 2 and t.stride(1) == 1 and t.shape == a.shape, n
    _call("mv_mul_bf16", _p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), a.shape[0], a.shape[1],
          _stream())
    return out


def attention_trace(q, k, v, out, trace, softmax_scale=None):
    """Diagnostics: attention() on the 128-key-step kernel + clock64 stamps into trace (int64 [2, steps, 8], cuda)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1 and t.stride(1) == 128, n
    _req(trace, torch.int64, "trace")
    assert trace.dim() == 3 and trace.shape[0] == 2 and trace.shape[2] == 8 and trace.is_contiguous()
    Lq, H, _ = q.shape
    _call("mv_attention_fwd_trace", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
          Lq, k.shape[0], H, float(softmax_scale if softmax_scale is not None else 128 ** -0.5), _p(trace),
          trace.shape[1], _stream())
    return out


def roles_config(hi=-1):
    """Diagnostics: 1 = TMA / MMA role warps at the highest warp ids of every tcgen05 kernel, 0 = lowest."""
    global CONFIG_EPOCH
    CONFIG_EPOCH += 1
    _check(lib().mv_roles_config(int(hi)), "mv_roles_config")


def gemm_config(pair=-1):
    """Diagnostics: 1 = CTA-pair (cta_group::2) GEMM kernel for the large linears, 0 = single-CTA kernel."""
    global CONFIG_EPOCH
    CONFIG_EPOCH += 1
    _check(lib().mv_gemm_config(int(pair)), "mv_gemm_config")


def vae_conv_config(pair=-1, tiles_per_cta=-1, epi_regs=-1):
    """Diagnostics: pair 1 = CTA-pair convolution kernel for the tensor-bound WanVAE convs, 0 = single-CTA kernel;
    tiles_per_cta 0 = automatic, 1 | 2 | 4 forced; epi_regs 1 = fused norm epilogue in one TMEM pass (row kept in
    registers), 0 = two passes; -1 keeps, -2 restores the default."""
    global CONFIG_EPOCH
    CONFIG_EPOCH += 1
    _check(lib().mv_vae_conv_config(int(pair), int(tiles_per_cta), int(epi_regs)), "mv_vae_conv_config")


def attention_config(kstep=-1, emu=-1, stale=-1, pingpong=-1, skew=-1, wait_spin=-1, pack=-1):
    """Diagnostics: pick the attention kernel variant for subsequent launches (negative = keep); tools/ab_step.py."""
    global CONFIG_EPOCH
    CONFIG_EPOCH += 1
    _check(lib().mv_attention_config(int(kstep), int(emu), int(stale), int(pingpong), int(skew), int(wait_spin),
                                       int(pack)),
           "mv_attention_config")
