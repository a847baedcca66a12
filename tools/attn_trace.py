"""Per-step timeline of the 128-key attention kernel on CTA (0, head 0): where a Q tile's step goes (clock cycles).
Usage: python tools/attn_trace.py [L] [H]   (MV_ATTN_PINGPONG=1 traces the ping-pong variant)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 75600
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    steps = 256
    mv.device_check()
    q = torch.randn(L, H, 128, device="cuda").bfloat16()
    k = torch.randn(L, H, 128, device="cuda").bfloat16()
    v = torch.randn(L, H, 128, device="cuda").bfloat16()
    o = torch.empty_like(q)
    tr = torch.zeros(2, steps, 8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        mv.attention_trace(q, k, v, o, tr)
    torch.cuda.synchronize()
    t = tr.cpu().double()
    s0, s1 = 64, 250                     # steady-state window
    names = ["S visible -> S in regs", "-> row max done", "-> exps done", "-> P handed over",
             "P handed over -> seen by MMA warp", "MMA warp: P seen -> PV+QK issued"]
    for w in range(2):
        x = t[w]
        print("tile %d: period %.0f clk/step" % (w, ((x[s1, 0] - x[s0, 0]) / (s1 - s0)).item()))
        d = [x[s0:s1, 1] - x[s0:s1, 0], x[s0:s1, 2] - x[s0:s1, 1], x[s0:s1, 3] - x[s0:s1, 2], x[s0:s1, 4] - x[s0:s1, 3],
             x[s0:s1, 5] - x[s0:s1, 4], x[s0:s1, 6] - x[s0:s1, 5]]
        for n, dd in zip(names, d):
            print("   %-36s mean %7.0f  min %7.0f  max %7.0f" % (n, dd.mean().item(), dd.min().item(), dd.max().item()))
        nxt = x[s0 + 1:s1 + 1, 0] - x[s0:s1, 4]
        print("   %-36s mean %7.0f  min %7.0f  max %7.0f" % ("P handed over -> next S visible", nxt.mean().item(),
                                                                nxt.min().item(), nxt.max().item()))
    off = (t[1, s0:s1, 2] - t[0, s0:s1, 2])
    print("tile1 exp start - tile0 exp start: mean %.0f" % off.mean().item())


if __name__ == "__main__":
    main()
