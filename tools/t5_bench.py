"""umT5-XXL encode benchmark: random-init encoder of the real architecture (24 layers, dim 4096,
64 heads, gated-GELU ffn 10240, vocab 256384, bf16), one prompt of `n` valid tokens per call, as
T5EncoderModel.__call__ runs it (valid prefix only) and as the reference runs it (all 512 padded rows).
The encoder is weight-bandwidth bound at these M: roofline = layer-weight bytes / HBM copy bandwidth.  One JSON line
per case."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402
from wan.modules.t5 import umt5_xxl  # noqa: E402


def main():
    mv.device_check()
    torch.manual_seed(11)
    enc = umt5_xxl(encoder_only=True, dtype=torch.bfloat16, device="cuda").eval().requires_grad_(False)
    for b in enc.blocks:  # keep the random-init forward well conditioned (T5 has no 1/sqrt(d))
        b.attn.q.weight.mul_(0.25)
    layer_bytes = sum(p.numel() * 2 for n, p in enc.named_parameters() if n.startswith("blocks.") and p.dim() == 2
                      and "pos_embedding" not in n)
    peak = 6485.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1, 256384, (512,), generator=g).cuda()
    for n, rows in ((512, 512), (128, 128), (128, 512), (32, 32)):
        mask = (torch.arange(512, device="cuda") < n).long()
        if rows == n:
            fn = lambda: enc.encode_prefix(ids, n)              # noqa: E731  what T5EncoderModel.__call__ does
        else:
            fn = lambda: enc(ids[None], mask[None])             # noqa: E731  the reference's padded forward
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            l0 = mv.LAUNCHES
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
            launches = mv.LAUNCHES - l0
        ms = sorted(ts)[len(ts) // 2]
        flops = 24 * 2.0 * rows * (4 * 4096 * 4096 + 3 * 4096 * 10240) + 24 * 4.0 * rows * min(n, rows) * 4096
        print(json.dumps(dict(metric="umt5_encode_ms", valid_tokens=n, rows=rows, ms=round(ms, 3),
                              gpu_launches=launches, tflops=round(flops / ms / 1e9, 1),
                              roofline=dict(bound="hbm", achieved=round(layer_bytes / ms / 1e6, 1), peak=peak,
                                            unit="GB/s", frac=round(layer_bytes / ms / 1e6 / peak, 3),
                                            bytes=layer_bytes),
                              finite=bool(torch.isfinite(out.float()).all()))), flush=True)


if __name__ == "__main__":
    main()
