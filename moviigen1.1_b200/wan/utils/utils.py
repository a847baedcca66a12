"""Symbols generate.py imports from wan.utils.utils (generate.py:22).  Media I/O is outside the hot path: tensors are
saved with torch.save when no video encoder is installed."""
import argparse
import os

import torch

__all__ = ["cache_video", "cache_image", "str2bool"]


def str2bool(v):
    """argparse helper with the reference's accepted spellings (wan/utils/utils.py:96-118)."""
    if isinstance(v, bool):
        return v
    s = v.lower()
    if s in ("true", "t", "yes", "y", "1"):
        return True
    if s in ("false", "f", "no", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected (True/False)")


def cache_video(tensor, save_file=None, fps=30, suffix=".mp4", nrow=8, normalize=True, value_range=(-1, 1),
                retry=5):
    """Writes [B, C, T, H, W] in [-1, 1] as a video when imageio is available, else as a uint8 tensor file."""
    path = save_file or "movii_out" + suffix
    x = tensor.detach().float().clamp(*value_range)
    x = ((x - value_range[0]) / (value_range[1] - value_range[0]) * 255).round().to(torch.uint8)
    frames = x[0].permute(1, 2, 3, 0).cpu()  # T, H, W, C
    try:
        import imageio
        w = imageio.get_writer(path, fps=fps, codec="libx264", quality=8)
        for fr in frames.numpy():
            w.append_data(fr)
        w.close()
        return path
    except Exception:
        alt = os.path.splitext(path)[0] + ".pt"
        torch.save(frames, alt)
        return alt


def cache_image(tensor, save_file, nrow=8, normalize=True, value_range=(-1, 1), retry=5):
    path = os.path.splitext(save_file)[0] + ".pt"
    torch.save(tensor.detach().cpu(), path)
    return path
