"""Per-shape timing of the WanVAE implicit-GEMM convolution (mv_vae_conv / mv_vae_conv_fused) at the decoder's 1080P
stage shapes, one temporal chunk of 4 latent frames: TFLOP/s per conv shape, so that the slow
shapes are visible; also the single-launch driver for `ncu --set full -k regex:conv_igemm`.
Usage: python tools/vae_conv_bench.py [stage ...]   stage in A B C D up2 (default: all)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402
from wan.modules.vae import _Conv, _parity_convs, _taps  # noqa: E402

DEV = "cuda"
# stage: (frames of one 4-latent-frame chunk, H, W, Cin, Cout)
STAGES = {"A": (4, 104, 240, 384, 384), "B": (8, 208, 480, 384, 384), "C": (16, 416, 960, 192, 192),
          "D": (16, 832, 1920, 96, 96)}


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


ONLY = [a for a in os.environ.get("CONV_VARIANTS", "").split(",") if a]   # e.g. CONV_VARIANTS=pair_nt2 for ncu


def main():
    which = sys.argv[1:] or ["A", "B", "C", "D", "up2"]
    mv.device_check()
    g = torch.Generator(device=DEV).manual_seed(0)
    for st in which:
        if st == "up2":        # upsamples.11: nearest-2x + Conv2d 192 -> 96 as four 2x2 sub-pixel convs (stage C -> D)
            T, H, W, Ci = 16, 416, 960, 192
            x = torch.randn(T, H, W, Ci, device=DEV, generator=g).half()
            w2 = torch.randn(Ci // 2, Ci, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Ci)
            par = _parity_convs(w2.cpu(), torch.zeros(Ci // 2), DEV)
            Co = Ci // 2
            out = torch.empty(T, 2 * H, 2 * W, Co, dtype=torch.float16, device=DEV)

            def run():
                for (a, b), c in par.items():
                    mv.vae_conv(x, c, out, o_base=(a * 2 * W + b) * Co, os_t=4 * H * W * Co, os_h=4 * W * Co, os_w=2 * Co)
            fl = 2.0 * 16 * Ci * Co * T * H * W            # 4 parities x 4 taps, as executed
            rec = dict(kind="vae_up2d_subpixel", shape=[T, H, W, Ci, Co])
            for label, pair, nt in (("1cta", 0, 0), ("pair_nt2", 1, 2), ("pair_nt4", 1, 4)):
                mv.vae_conv_config(pair, nt)
                rec[label + "_tflops_executed"] = round(fl / timeit(run) / 1e9, 1)
            mv.vae_conv_config(-2, -2)
            print(json.dumps(rec), flush=True)
            continue
        if st == "head":       # the decoder head: 1x1x1 conv 96 -> 27 x 4 partial sums + neighbour gather vs the direct 16-column conv
            T, H, W, Ci = 16, 832, 1920, 96
            x = torch.randn(T, H, W, Ci, device=DEV, generator=g).half()
            wt = torch.randn(112, Ci, 1, 1, 1, device=DEV, generator=g) / math.sqrt(Ci)
            c1 = _Conv(wt.cpu(), None, _taps(1, 1, 1), DEV)
            D = torch.empty(T, H, W, 112, dtype=torch.float16, device=DEV)
            video = torch.empty(3, T, H, W, device=DEV)
            ms1 = timeit(lambda: mv.vae_conv(x, c1, D, o_base=0, os_t=H * W * 112, os_h=W * 112, os_w=112))
            ms2 = timeit(lambda: mv.vae_head_gather(D, None, [0.0, 0.0, 0.0], video, 0))
            gb1 = T * H * W * (Ci + 112) * 2 / 1e9
            gb2 = (T * H * W * 112 * 2 + 3 * T * H * W * 4) / 1e9
            print(json.dumps(dict(kind="vae_head", shape=[T, H, W, Ci], conv1x1_ms=round(ms1, 3), conv1x1_gbs=round(gb1 / ms1 * 1e3, 1),
                                  gather_ms=round(ms2, 3), gather_gbs=round(gb2 / ms2 * 1e3, 1))), flush=True)
            continue
        if st.startswith("X:"):      # custom shape X:T:H:W:Cin:Cout (experiments on what bounds a channel plan)
            T, H, W, Ci, Co = (int(v) for v in st.split(":")[1:])
        else:
            T, H, W, Ci, Co = STAGES[st]
        x = torch.randn(T + 2, H, W, Ci, device=DEV, generator=g).half()
        wt = torch.randn(Co, Ci, 3, 3, 3, device=DEV, generator=g) / math.sqrt(27 * Ci)
        c = _Conv(wt.cpu(), torch.zeros(Co), _taps(3, 3, 3), DEV)
        out = torch.empty(T, H, W, Co, dtype=torch.float16, device=DEV)
        res = torch.randn(T, H, W, Co, device=DEV, generator=g).half()
        gamma = torch.ones(Co, device=DEV)
        kw = dict(o_base=0, os_t=H * W * Co, os_h=W * Co, os_w=Co, t_off=2)
        fl = 2.0 * 27 * Ci * Co * T * H * W
        rec = dict(kind="vae_conv3x3x3", stage=st, shape=[T, H, W, Ci, Co])
        nout = torch.empty_like(out)
        for label, pair, nt in (("1cta", 0, 0), ("pair_nt1", 1, 1), ("pair_nt2", 1, 2), ("pair_nt4", 1, 4)):
            if ONLY and label not in ONLY:
                continue
            mv.vae_conv_config(pair, nt)
            ms = timeit(lambda: mv.vae_conv(x, c, out, res=res, **kw))
            rec[label + "_tflops"] = round(fl / ms / 1e9, 1)
            if Co <= 256:
                for epi in (0, 1):
                    mv.vae_conv_config(epi_regs=epi)
                    tag = "_fused" + ("_regs" if epi else "")
                    ms2 = timeit(lambda: mv.vae_conv_fused(x, c, None, gamma, nout, **kw))
                    rec[label + tag + "_tflops"] = round(fl / ms2 / 1e9, 1)
                    ms3 = timeit(lambda: mv.vae_conv_fused(x, c, out, gamma, nout, res=res, **kw))
                    rec[label + tag + "_res_tflops"] = round(fl / ms3 / 1e9, 1)
                    if os.environ.get("CONV_SPLIT"):      # which part of the fused+res epilogue costs: residual load or second store
                        ms4 = timeit(lambda: mv.vae_conv_fused(x, c, None, gamma, nout, res=res, **kw))
                        rec[label + tag + "_resonly_tflops"] = round(fl / ms4 / 1e9, 1)
                        ms5 = timeit(lambda: mv.vae_conv_fused(x, c, out, gamma, nout, **kw))
                        rec[label + tag + "_rawonly_tflops"] = round(fl / ms5 / 1e9, 1)
        mv.vae_conv_config(-2, -2, -2)
        print(json.dumps(rec), flush=True)
        del x, out, res


if __name__ == "__main__":
    main()
