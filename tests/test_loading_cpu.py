"""CPU: diffusers-free checkpoint loading: WanModel.from_pretrained reads config.json + *.safetensors
exactly as the reference's ModelMixin.from_pretrained call (wan/text2video.py:87) expects them; WanVAE_ accepts a full
reference VAE state dict (both halves)."""
import json

import pytest
import torch

safetensors = pytest.importorskip("safetensors.torch")


def test_wan_model_from_pretrained_roundtrip(tmp_path):
    from wan.modules.model import WanModel
    cfg = dict(model_type="t2v", text_len=16, in_dim=16, dim=256, ffn_dim=512, freq_dim=64, out_dim=16, num_heads=2,
               num_layers=2, eps=1e-6)
    m = WanModel(text_dim=4096, **cfg)
    torch.nn.init.normal_(m.head.head.weight, std=0.02)
    sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    keys = sorted(sd)
    half = len(keys) // 2
    safetensors.save_file({k: sd[k] for k in keys[:half]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    safetensors.save_file({k: sd[k] for k in keys[half:]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    with open(tmp_path / "config.json", "w") as fh:
        json.dump(dict(cfg, _class_name="WanModel", _diffusers_version="0.30.0"), fh)
    m2 = WanModel.from_pretrained(str(tmp_path))
    sd2 = m2.state_dict()
    assert sorted(sd2) == keys
    for k in keys:
        assert torch.equal(sd[k], sd2[k]), k
    assert m2.num_layers == 2 and m2.dim == 256


def test_vae_accepts_full_reference_state_dict():
    from wan.modules.vae import WanVAE_
    m = WanVAE_()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["encoder.conv1.weight"] = torch.full((96, 3, 3, 3, 3), 0.5)   # Wan2.1_VAE.pth carries both halves
    sd["conv1.weight"] = torch.zeros(32, 32, 1, 1, 1)
    sd["decoder.head.2.bias"] = torch.full((3,), 0.25)
    m.load_state_dict(sd)
    assert torch.equal(m.decoder.head[2].bias.detach(), torch.full((3,), 0.25))
    assert torch.equal(m.encoder.conv1.weight.detach(), torch.full((96, 3, 3, 3, 3), 0.5))
    # a decoder-only state dict (what round-1 checkpoints of this repo hold) leaves the encoder half as it is
    m.load_state_dict({k: v for k, v in sd.items() if k.startswith(("decoder.", "conv2."))})
    assert torch.equal(m.encoder.conv1.weight.detach(), torch.full((96, 3, 3, 3, 3), 0.5))
