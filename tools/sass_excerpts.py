"""Hot-loop SASS excerpts per kernel (north_star: "ncu-captured utilisation committed next to its SASS").

For every kernel of lib/obj/*.o: the mnemonic histogram and the code regions dense in the instructions that matter for
that kernel (UTCHMMA / UTMALDG / LDTM / STTM / UTCBAR / MUFU / LDG / STG ...), encodings stripped.
Runs on the build box (cuobjdump only, no GPU).  Usage: python tools/sass_excerpts.py [out_dir]  (default profiles/r02_sass)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "moviigen1.1_b200", "lib", "obj")
KEY = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|MUFU|LDG|STG|LDS|STS|SHFL|REDUX|HMMA|"
                 r"SYNCS|UCGABAR|ATOM|RED)\b")
# (label, object file, regex on the demangled-ish function name, max loops to print)
KERNELS = [
    ("attention_fwd_k128 (default: fixed-reference softmax; also the Ulysses scatter epilogue)", "attention_sm100.o",
     r"attention_fwd_k128_kernelILi0ELb0ELb0ELb1ELb0ELb0E", 4),
    ("attention_fwd (64-key variant)", "attention_sm100.o", r"attention_fwd_kernelILi1E", 3),
    ("gemm_bf16 single CTA, residual epilogue", "gemm_sm100.o", r"gemm_bf16_kernelILi2ELi256ELb0E", 4),
    ("gemm_bf16 single CTA, bf16+GELU epilogue", "gemm_sm100.o", r"gemm_bf16_kernelILi1ELi256ELb0E", 3),
    ("gemm_bf16 CTA pair (cta_group::2), residual epilogue", "gemm_sm100.o", r"gemm_bf16_pair_kernelILi2ELb0E", 4),
    ("conv_igemm_pair BK=64 NT=1 (WanVAE stages A-C, cta_group::2)", "vae_conv_sm100.o", r"conv_igemm_pair_kernelILi64ELi1ELb0E", 4),
    ("conv_igemm_pair BK=32 NT=2 (WanVAE stage D, cta_group::2)", "vae_conv_sm100.o", r"conv_igemm_pair_kernelILi32ELi2ELb0E", 4),
    ("conv_igemm BK=64 (single CTA: 1x1x1 convs, strided encoder convs)", "vae_conv_sm100.o", r"conv_igemm_kernelILi64ELb0E", 4),
    ("conv_igemm BK=32 (single CTA, Cin=96)", "vae_conv_sm100.o", r"conv_igemm_kernelILi32ELb0E", 3),
    ("conv_igemm BK=16 (16-channel stems)", "vae_conv_sm100.o", r"conv_igemm_kernelILi16ELb0E", 3),
    ("rmsnorm_silu_cl (WanVAE)", "vae_conv_sm100.o", r"rmsnorm_silu_cl_kernelILi32ELi1E", 2),
    ("softmax_rows_reg (WanVAE attention, register-resident row)", "vae_conv_sm100.o", r"softmax_rows_reg_kernel", 3),
    ("vae_head_gather (decoder head: 27-neighbour gather of partial sums)", "vae_conv_sm100.o", r"vae_head_gather_kernel", 2),
    ("vae_latent_in", "vae_conv_sm100.o", r"vae_latent_in_kernel", 1),
    ("vae_video_in (encoder)", "vae_conv_sm100.o", r"vae_video_in_kernel", 1),
    ("vae_latent_out (encoder)", "vae_conv_sm100.o", r"vae_latent_out_kernel", 1),
    ("ln_modulate", "rowops.o", r"ln_modulate_kernel", 2),
    ("qkv_norm_rope (C = 5120; also the Ulysses p2p head scatter)", "rowops.o", r"qkv_norm_rope_kernelILi20E", 2),
    ("unipc_cfg_step", "rowops.o", r"unipc_cfg_step_kernel", 1),
    ("modulation_table", "rowops.o", r"modulation_table_kernel", 1),
    ("head_unpatchify", "rowops.o", r"head_unpatchify_kernel", 3),
    ("patchify", "rowops.o", r"\d+patchify_kernel", 1),
    ("linear_f32_vec", "rowops.o", r"linear_f32_vec_kernel", 2),
    ("sp_barrier", "p2p_sm100.o", r"sp_barrier_kernel", 1),
    ("t5_attention", "t5_sm100.o", r"t5_attention_kernel", 3),
    ("t5_rmsnorm", "t5_sm100.o", r"t5_rmsnorm_kernel", 1),
]
LINE = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;\s*/\*")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = LINE.match(ln)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def loops(ins):
    """(start_idx, end_idx) of every backward branch's body, innermost first."""
    addr2idx = {a: i for i, (a, _) in enumerate(ins)}
    res = []
    for i, (a, txt) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.[A-Z0-9_.]+)?\s+(?:!?U?P\d+,\s*)?`?\(?\.?(?:L_x_\d+|0x([0-9a-f]+))", txt)
        if not m or not txt.lstrip("@!UP0123456789 ").startswith("BRA"):
            continue
        tgt = re.search(r"0x([0-9a-f]+)", txt)
        if tgt is None:
            continue
        t = int(tgt.group(1), 16)
        if t <= a and t in addr2idx:
            res.append((addr2idx[t], i))
    res.sort(key=lambda se: se[1] - se[0])
    return res


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass")
    os.makedirs(out_dir, exist_ok=True)
    cache = {}
    index = []
    for label, objname, pat, nloops in KERNELS:
        obj = os.path.join(OBJ, objname)
        if objname not in cache:
            cache[objname] = functions(obj)
        names = [n for n in cache[objname] if re.search(pat, n)]
        if not names:
            index.append("%-60s NOT FOUND (%s)" % (label, pat))
            continue
        name = names[0]
        ins = cache[objname][name]
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
        fn = re.sub(r"[^A-Za-z0-9]+", "_", label).strip("_").lower() + ".txt"
        with open(os.path.join(out_dir, fn), "w") as fh:
            fh.write("%s\nfunction %s\nobject   moviigen1.1_b200/lib/obj/%s (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)\n"
                     "%d SASS instructions\n\n" % (label, name, objname, len(ins)))
            fh.write("mnemonic histogram (top 24): " + ", ".join("%s %d" % kv for kv in hist.most_common(24)) + "\n")
            # regions = maximal runs where an instruction of interest occurs at least every 24 instructions
            hot = [i for i, (_, t) in enumerate(ins)
                   if (m := KEY.search(t)) and m.group(1) not in ("SYNCS", "UCGABAR", "SHFL", "REDUX", "LDS", "STS")]
            regions, cur = [], None
            for i in hot:
                w = 50 if re.search(r"\b(UTCHMMA|UTMALDG|UTCBAR|LDTM|STTM)\b", ins[i][1]) else 1   # tensor / TMA / TMEM first
                if cur is not None and i - cur[1] <= 24:
                    cur[1] = i
                    cur[2] += w
                else:
                    cur = [i, i, w]
                    regions.append(cur)
            regions.sort(key=lambda r: -r[2])
            for s_, e_, w in sorted(regions[:nloops], key=lambda r: r[0]):
                lo, hi = max(0, s_ - 6), min(len(ins) - 1, e_ + 6)
                body = ins[lo:hi + 1]
                keys = collections.Counter(m.group(1) for _, t in body for m in [KEY.search(t)] if m)
                fh.write("\n---- region @0x%04x..0x%04x (%d instructions): %s\n" %
                         (body[0][0], body[-1][0], len(body), ", ".join("%s x%d" % kv for kv in keys.most_common())))
                show = body if len(body) <= 140 else body[:70] + [(-1, "... (%d instructions elided) ..." % (len(body) - 140))] + body[-70:]
                for a_, t in show:
                    fh.write(("        %s\n" % t) if a_ < 0 else ("  %04x  %s\n" % (a_, t)))
        index.append("%-60s %-52s %5d instr  %s" % (label, fn, len(ins), " ".join(
            "%s:%d" % (k, hist[k]) for k in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "MUFU", "HMMA") if hist.get(k))))
    with open(os.path.join(out_dir, "INDEX.txt"), "w") as fh:
        fh.write("Hot-loop SASS excerpts per kernel (tools/sass_excerpts.py).  HMMA (legacy mma.sync) must be absent everywhere.\n\n")
        fh.write("\n".join(index) + "\n")
    print("\n".join(index))


if __name__ == "__main__":
    main()
