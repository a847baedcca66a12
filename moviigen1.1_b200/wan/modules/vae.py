"""WanVAE (decoder) — placeholder until the native 3-D causal-conv decoder lands (see DESIGN.md status table)."""
import torch


class _Dims:
    z_dim = 16


class WanVAE:
    def __init__(self, z_dim=16, vae_pth=None, dtype=torch.float, device="cuda"):
        self.device = device
        self.model = _Dims()

    def decode(self, zs):
        raise NotImplementedError("native WanVAE decoder not built yet")
