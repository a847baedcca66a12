"""Diagnostic: per-(head, 32-row warp block) error map of the attention kernel vs an fp32 reference."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv
torch.manual_seed(0)
for (Lq, Lk, H) in [(256, 256, 2), (256, 128, 1), (2048, 2048, 4), (512, 320, 1), (256, 192, 1), (256, 512, 2)]:
    bad_total = 0
    for rep in range(3):
        q = torch.randn(Lq, H, 128, device="cuda").bfloat16()
        k = torch.randn(Lk, H, 128, device="cuda").bfloat16()
        v = torch.randn(Lk, H, 128, device="cuda").bfloat16()
        o = torch.empty_like(q)
        mv.attention(q, k, v, o)
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(0, 1), k.float().transpose(0, 1),
                                                               v.float().transpose(0, 1)).transpose(0, 1)
        err = (o.float() - ref).abs().amax(dim=2)          # [Lq, H]
        blocks = err.view(Lq // 32, 32, H).amax(dim=1)      # [Lq/32, H]
        bad = (blocks > 0.02).nonzero().tolist()
        bad_total += len(bad)
        if bad and rep == 0:
            print("  case", (Lq, Lk, H), "bad (rowblock, head):", bad[:24], "max", float(err.max()))
            # how wrong: compare with reference restricted to subsets of 64-key steps
            rb, hh = bad[0]
            rows = slice(rb * 32, rb * 32 + 32)
            qs = q[rows, hh].float(); ks = k[:, hh].float(); vs = v[:, hh].float()
            s = qs @ ks.t() / 128 ** 0.5
            p = torch.softmax(s, dim=-1)
            full = p @ vs
            n = Lk // 64
            for drop in range(n):
                pm = p.clone(); pm[:, drop * 64:(drop + 1) * 64] = 0
                alt = (pm @ vs)
                d = (o[rows, hh].float() - alt).abs().max().item()
                print("     drop step %d (no renorm): max diff %.4f" % (drop, d))
    print("case", (Lq, Lk, H), "dbg", os.environ.get("MV_ATTN_DBG", "0"), "bad blocks over 3 reps:", bad_total)
