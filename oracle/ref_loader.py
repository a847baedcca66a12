"""Imports the REAL reference modules (wan/modules/{attention,model,vae}.py) from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not exist
on the GPU box.  Used by oracle/make_golden.py to mint tests/golden/*.pt and by the optional
`reference_available()` tests.  Nothing is copied: the files are executed where they lie.

Shims: `diffusers` is not installed -> stub ConfigMixin / ModelMixin /
register_to_config; flash_attention asserts CUDA -> on CPU it is swapped for an SDPA restatement
with the same [B, L, N, D] signature.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("MOVII_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "wan", "modules", "model.py"))


def _stub_diffusers():
    if "diffusers" in sys.modules:
        return
    cu = types.ModuleType("diffusers.configuration_utils")
    mu = types.ModuleType("diffusers.models.modeling_utils")

    class ConfigMixin:
        pass

    class ModelMixin(nn.Module):
        pass

    cu.ConfigMixin, cu.register_to_config, mu.ModelMixin = ConfigMixin, (lambda f: f), ModelMixin
    sys.modules["diffusers"] = types.ModuleType("diffusers")
    sys.modules["diffusers.configuration_utils"] = cu
    sys.modules["diffusers.models"] = types.ModuleType("diffusers.models")
    sys.modules["diffusers.models.modeling_utils"] = mu


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def sdpa_flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None,
                         causal=False, window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16,
                         version=None):
    """CPU stand-in for flash_attention (attention.py:24-130): same layout, key lengths honoured."""
    outs = []
    for i in range(q.size(0)):
        kl = int(k_lens[i]) if k_lens is not None else k.size(1)
        qi, ki, vi = (t.transpose(1, 2).float() for t in (q[i:i + 1], k[i:i + 1, :kl], v[i:i + 1, :kl]))
        o = torch.nn.functional.scaled_dot_product_attention(qi, ki, vi, is_causal=causal, scale=softmax_scale)
        outs.append(o.transpose(1, 2))
    return torch.cat(outs).to(q.dtype).contiguous()


_cache = {}


def load_reference(patch_attention_for_cpu=True):
    """Returns (attention_module, model_module, vae_module) of the reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise RuntimeError("reference not found at %s" % REF_ROOT)
    _stub_diffusers()
    for pkg, path in (("refwan", os.path.join(REF_ROOT, "wan")),
                      ("refwan.modules", os.path.join(REF_ROOT, "wan", "modules"))):
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    att = _load("refwan.modules.attention", os.path.join(REF_ROOT, "wan", "modules", "attention.py"))
    model = _load("refwan.modules.model", os.path.join(REF_ROOT, "wan", "modules", "model.py"))
    vae = _load("refwan.modules.vae", os.path.join(REF_ROOT, "wan", "modules", "vae.py"))
    if patch_attention_for_cpu and not torch.cuda.is_available():
        model.flash_attention = sdpa_flash_attention
    _cache["mods"] = (att, model, vae)
    return _cache["mods"]


def load_reference_t5():
    """The reference's wan/modules/t5.py (T5Encoder etc.).  Its `from .tokenizers import HuggingfaceTokenizer` is
    satisfied by a stub module (tokenizers.py needs `ftfy`, which is not installed, and is not arithmetic)."""
    if "t5" in _cache:
        return _cache["t5"]
    load_reference()
    tk = types.ModuleType("refwan.modules.tokenizers")

    class HuggingfaceTokenizer:  # never constructed by the golden script
        def __init__(self, *a, **k):
            raise RuntimeError("tokenizer stub")

    tk.HuggingfaceTokenizer = HuggingfaceTokenizer
    sys.modules["refwan.modules.tokenizers"] = tk
    # t5.py:478 evaluates torch.cuda.current_device() as a default argument at class-definition time
    real = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0
    try:
        _cache["t5"] = _load("refwan.modules.t5", os.path.join(REF_ROOT, "wan", "modules", "t5.py"))
    finally:
        torch.cuda.current_device = real
    return _cache["t5"]
