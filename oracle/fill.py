"""Deterministic parameter fill shared by oracle/make_golden.py (real reference, build container) and the tests
(our modules / the oracle, anywhere): same names + same seed -> same weights, so golden files only need to store
inputs and outputs.  TEST INFRASTRUCTURE ONLY."""
import math

import torch


def fill_parameters(module_or_named, seed):
    """Overwrites every parameter in sorted-name order with seeded CPU randoms (then copies to the param's device)."""
    named = module_or_named.named_parameters() if hasattr(module_or_named, "named_parameters") else module_or_named
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(named, key=lambda kv: kv[0]):
            shape = tuple(p.shape)
            if name.endswith("modulation"):
                v = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
            elif p.dim() >= 2:
                fan_in = p.numel() // shape[0]
                v = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            elif name.endswith("bias"):
                v = 0.1 * torch.randn(shape, generator=g)
            else:  # norm gains
                v = 1.0 + 0.1 * torch.randn(shape, generator=g)
            p.copy_(v.to(p.dtype).to(p.device))


def state_dict_like(shapes, seed):
    """Same fill for a {name: shape} description (used to build an oracle state dict without any nn.Module)."""
    params = {k: torch.empty(v) for k, v in shapes.items()}
    fill_parameters(list(params.items()), seed)
    return params


def fill_t5(module_or_named, seed):
    """fill_parameters, then T5-specific conditioning: T5 attention has no 1/sqrt(d) scaling, so unit-variance q/k
    rows would give near one-hot softmaxes; shrink q and make the relative-position bias O(1) so it matters."""
    named = list(module_or_named.named_parameters() if hasattr(module_or_named, "named_parameters")
                 else module_or_named)
    fill_parameters(named, seed)
    with torch.no_grad():
        for name, p in named:
            if name.endswith("attn.q.weight"):
                p.mul_(0.25)
            elif name.endswith("pos_embedding.embedding.weight"):
                p.mul_(3.0)
            elif name == "token_embedding.weight":
                p.mul_(math.sqrt(p.shape[1]))


def t5_state_dict_like(shapes, seed):
    params = {k: torch.empty(v) for k, v in shapes.items()}
    fill_t5(list(params.items()), seed)
    return params
