"""Ulysses sequence parallelism for the DiT self-attention over NCCL (one process per GPU).

What the reference does through xfuser (wan/distributed/xdit_context_parallel.py:155-198 ->
xFuserLongContextAttention; data movement spelled out in scripts/train/model/model_seq.py:231-256 and
SURVEY.md Appendix C):  rank r owns tokens [r L/P, (r+1) L/P);  after QKV + full-row RMSNorm + RoPE (global
positions) the heads are scattered / the sequence gathered with an all-to-all, every rank attends over the full
sequence for its 40/P heads, and a second all-to-all brings the outputs back to the token owners.

B200-native realisation here:
  * the head scatter is fused into the RMSNorm/RoPE pass (mv_qkv_prepare writes the send layout
    [dst][local token][local heads x 128] directly),
  * q, k, v travel as ONE grouped NCCL operation (3 all_to_all_single calls inside a coalescing group are
    pairwise send/recv over NVLink 5 / NVSwitch),
  * attention output is produced directly in the return send layout ([dst][token][heads x 128] is just the
    row-major [L, heads x 128] result), and the o projection consumes the received
    [src][local token][heads x 128] slabs through a 3-D TMA tensor map (mv_gemm_bf16_ksplit) — no transpose
    pass on either side.
Only the collective itself goes through torch.distributed (plumbing).
"""
import torch
import torch.distributed as dist


class UlyssesGroup:
    """Process-group facts + communication buffers for one (rows, dim) shape."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._buf = {}

    def buffers(self, rows, dim, device, dtype=torch.bfloat16):
        key = (rows, dim, str(device))
        b = self._buf.get(key)
        if b is None:
            self._buf.clear()
            P = self.world
            mk = lambda: torch.empty(P, rows, dim // P, dtype=dtype, device=device)  # noqa: E731
            b = dict(q_s=mk(), k_s=mk(), v_s=mk(), q_r=mk(), k_r=mk(), v_r=mk(), o_s=mk(), o_r=mk())
            self._buf[key] = b
        return b

    # -- collectives (plumbing) -----------------------------------------------------------------------
    def all_to_all(self, outs, ins):
        """outs[i] <- all_to_all(ins[i]) for equally shaped [P, ...] tensors."""
        if self.world == 1:
            for o, i in zip(outs, ins):
                o.copy_(i)
            return
        for o, i in zip(outs, ins):
            dist.all_to_all_single(o.view(-1), i.view(-1), group=self.group)

    def all_gather_rows(self, local):
        """[rows, n] -> [P*rows, n] in rank order (xdit_context_parallel.py:148)."""
        if self.world == 1:
            return local
        out = torch.empty(self.world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out


def token_range(seq_len, world, rank):
    """Contiguous rank-major token chunk (torch.chunk semantics, xdit_context_parallel.py:137-139); seq_len is a
    multiple of world (text2video.py:164-166)."""
    if seq_len % world != 0:
        raise ValueError("seq_len %d is not a multiple of the sequence-parallel size %d" % (seq_len, world))
    n = seq_len // world
    return rank * n, n


def sp_self_attention(mv, grp, ws, rows, bw, cs, kv_len_total, prepare=None, attend=None):
    """Self-attention core for the local `rows` tokens whose fused QKV GEMM output sits in ws.qkv.

    Leaves the attention output as K-split slabs in the returned tensor [P, rows, C/P] (consumed by
    mv.gemm_ksplit).  `prepare` / `attend` are injection points for the gloo CPU tests of this choreography;
    in production they are the sm_100a kernels."""
    P = grp.world
    C = bw.dim
    Hl = bw.num_heads // P
    b = grp.buffers(rows, C, ws.qkv.device, ws.qkv.dtype)
    prepare = prepare or (lambda x, w, c, out: mv.qkv_prepare(x, w, c, out, P, 128, bw.eps))
    qkv = ws.qkv[:rows]
    prepare(qkv[:, 0:C], bw.g_q, cs, b["q_s"])
    prepare(qkv[:, C:2 * C], bw.g_k, cs, b["k_s"])
    prepare(qkv[:, 2 * C:3 * C], None, None, b["v_s"])
    grp.all_to_all([b["q_r"], b["k_r"], b["v_r"]], [b["q_s"], b["k_s"], b["v_s"]])
    L = P * rows
    q = b["q_r"].view(L, Hl, C // bw.num_heads)
    k = b["k_r"].view(L, Hl, C // bw.num_heads)[:kv_len_total]
    v = b["v_r"].view(L, Hl, C // bw.num_heads)[:kv_len_total]
    o = b["o_s"].view(L, Hl, C // bw.num_heads)
    if attend is None:
        mv.attention(q, k, v, o)
    else:
        attend(q, k, v, o)
    grp.all_to_all([b["o_r"]], [b["o_s"]])
    return b["o_r"]
