"""One launch of every small / memory-bound kernel at a representative production shape, as the single-process driver of

    ncu --set full --clock-control none -k regex:'<names>' -o gpurun_out/small python tools/ncu_small_kernels.py

(profiles/r02_ncu_small_kernels.txt).  The Ulysses variants run with LOCAL destination tables (on a multi-GPU box the same
kernels store into IPC-mapped peer slabs; ncu is never run on a multi-rank command)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "moviigen1.1_b200"))
import movii_b200 as mv  # noqa: E402

DEV = "cuda"


def main():
    mv.device_check()
    g = torch.Generator(device=DEV).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)  # noqa: E731
    M, C = 75600, 5120                                            # 720P token count, 14B width
    # --- DiT row kernels
    x = rn(M, C)
    h = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    mv.ln_modulate(x, h, rn(C), rn(C))
    qkv = rn(M, 3 * C).bfloat16()
    cs = rn(M, 64, 2)
    mv.qkv_norm_rope(qkv, rn(C), rn(C), cs)                                               # q and k in place
    P = 8
    bufs = [torch.empty(P, M, C // P, dtype=torch.bfloat16, device=DEV) for _ in range(3)]
    tabs = tuple(mv.ptr_table([b[d].data_ptr() for d in range(P)]) for b in bufs)
    mv.qkv_norm_rope(qkv, rn(C), rn(C), cs, dst=tabs, n_dst=P, src_slot=0)               # + Ulysses head scatter (local slabs)
    mv.qkv_prepare_p2p(qkv[:, 2 * C:], None, None, tabs[2], 0, P, 128, 1e-6)              # v slab alone (pipelined mode)
    del bufs, qkv
    lat = [rn(16 * 21 * 90 * 160) for _ in range(9)]
    coef = [5.0, 0.9, 1.0, 2.0, 0.9, 0.1, 0.2, 0.5, 1.1, 0.0, 0.5, 0.0, 2.0, 0.9, 0.1, 0.2, 1.1, 0.0, 0.5, 0.0]
    mv.unipc_cfg_step(lat[0], lat[1], lat[2], lat[3], [lat[4], lat[5], None], coef, lat[6], lat[7], lat[8])
    mods, e0 = rn(40, 6, C), rn(6 * C)
    mv.modulation_table(mods, e0, torch.empty_like(mods))
    latent = rn(16, 21, 90, 160)
    patches = torch.empty(M, 64, dtype=torch.bfloat16, device=DEV)
    mv.patchify(latent, patches)
    out = torch.empty(16, 21, 90, 160, device=DEV)
    mv.head_unpatchify(x, rn(C), rn(C), rn(64, C) / math.sqrt(C), rn(64), out, (21, 45, 80))
    sin = torch.empty(256, device=DEV)
    mv.sinusoid_embed(torch.tensor([937], device=DEV), sin)
    e = torch.empty(C, device=DEV)
    mv.linear_f32_vec(sin, rn(C, 256), rn(C), e, act_in=0)
    mv.linear_f32_vec(e, rn(6 * C, C), rn(6 * C), torch.empty(6 * C, device=DEV), act_in=1)
    # --- Ulysses: attention with the scatter epilogue (5 heads = the N = 8 share, local destinations), flag barrier
    L, H, nd = 16384, 5, 8
    q, k, v = (rn(L, H, 128).bfloat16() for _ in range(3))
    rows = L // nd
    dsts = [torch.empty(nd, rows, H * 128, dtype=torch.bfloat16, device=DEV) for _ in range(nd)]
    mv.attention_scatter(q, k, v, mv.ptr_table([d.data_ptr() for d in dsts]), nd, 0, rows, H * 128)
    flags = torch.zeros(64, dtype=torch.int32, device=DEV)
    mv.sp_barrier(mv.ptr_table([flags.data_ptr()]), flags.data_ptr(), 0, 1, 1)
    del x, h, dsts
    # --- WanVAE small kernels at 1080P stage shapes
    a = rn(8, 832, 1920, 96).half()                                   # stage D, 8 frames
    mv.vae_rmsnorm_silu(a, torch.empty_like(a), torch.ones(96, device=DEV), True)
    b = rn(8, 208, 480, 384).half()                                   # stage B
    mv.vae_rmsnorm_silu(b, torch.empty_like(b), torch.ones(384, device=DEV), True)
    hw = 104 * 240                                                    # middle attention: one frame's score matrix
    S = rn(4096, hw)
    Pm = torch.empty(4096, hw, dtype=torch.float16, device=DEV)
    mv.softmax_rows(S, Pm, hw, 1.0 / math.sqrt(384))
    z = rn(16, 4, 104, 240)
    mv.vae_latent_in(z, rn(16, 16), rn(16), rn(16), rn(16).abs() + 0.5, torch.empty(4, 104, 240, 16, dtype=torch.float16, device=DEV))
    D = (rn(8, 832, 1920, 112) * 0.1).half()
    mv.vae_head_gather(D, None, [0.0, 0.1, -0.1], torch.empty(3, 8, 832, 1920, device=DEV), 0)
    vid = torch.rand(3, 8, 832, 1920, device=DEV) * 2 - 1
    mv.vae_video_in(vid, 0, 8, torch.empty(8, 832, 1920, 16, dtype=torch.float16, device=DEV))
    hd = rn(4, 104, 240, 32).half()
    mv.vae_latent_out(hd, rn(32, 32), rn(32), rn(16), rn(16).abs() + 0.5, torch.empty(16, 4, 104, 240, device=DEV), 0)
    # --- umT5 row kernels (512 tokens, dim 4096 / ffn 10240)
    xt = rn(512, 4096)
    mv.t5_rmsnorm(xt, rn(4096), torch.empty(512, 4096, dtype=torch.bfloat16, device=DEV))
    table = rn(4096, 4096).bfloat16()
    ids = torch.randint(0, 4096, (512,), device=DEV)
    mv.embed_gather(table, ids, torch.empty(512, 4096, device=DEV))
    f1, f2 = rn(512, 10240).bfloat16(), rn(512, 10240).bfloat16()
    mv.mul_bf16(f1, f2, torch.empty_like(f1))
    qt, kt, vt = (rn(512, 64, 64).bfloat16() for _ in range(3))
    mv.t5_attention(qt, kt, vt, torch.empty_like(qt), bias=rn(64, 1023), bias_center=511)
    torch.cuda.synchronize()
    print("ok: %d launches" % mv.LAUNCHES)


if __name__ == "__main__":
    main()
