"""Ulysses sequence parallelism for the DiT self-attention over NCCL (one process per GPU).

What the reference does through xfuser (wan/distributed/xdit_context_parallel.py:155-198 ->
xFuserLongContextAttention; data movement spelled out in scripts/train/model/model_seq.py:231-256 and
SURVEY.md Appendix C):  rank r owns tokens [r L/P, (r+1) L/P);  after QKV + full-row RMSNorm + RoPE (global
positions) the heads are scattered / the sequence gathered with an all-to-all, every rank attends over the full
sequence for its 40/P heads, and a second all-to-all brings the outputs back to the token owners.

B200-native realisation here:
  * the head scatter of q, k AND v is fused into ONE RMSNorm/RoPE pass over the fused QKV rows (mv_qkv_norm_rope
    writes the send layout [dst][local token][local heads x 128] directly — or, in p2p mode, the peers' HBM),
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
    """Process-group facts + communication buffers for one (rows, dim) shape.

    mode 'p2p' (default on CUDA, MOVII_SP_MODE=nccl to disable): the exchange is FUSED into the producing kernels —
    every rank maps its peers' receive buffers over NVLink (CUDA IPC) and mv_qkv_prepare_p2p / mv_attention_fwd_scatter
    store into them directly; two flag barriers per layer order producers and consumers (mv_sp_barrier).
    mode 'nccl': 3 + 1 all_to_all_single calls per layer (the baseline; also what the gloo CPU tests exercise)."""

    def __init__(self, group=None, mode=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        import os
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._buf = {}
        self.mode = mode or os.environ.get("MOVII_SP_MODE", "p2p")
        self._p2p = {}
        self.epoch = 0
        # p2p mode, MOVII_SP_PIPELINE=1: the QKV GEMM runs as three column-slab GEMMs and the NVLink scatter of q (k) is
        # issued on a side stream while the k (v) GEMM runs.  Bit-identical; measured at N = 2: non-attention time -8 ms,
        # attention +16 ms per step (the step is energy-bound: hiding a low-power phase lowers the clock of what
        # follows) -> off by default (profiles/r02_ab_sp_pipeline_n2.jsonl)
        self.pipeline = os.environ.get("MOVII_SP_PIPELINE", "0") == "1"
        self._side = None

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

    # -- NVLink peer mapping (p2p mode) --------------------------------------------------------------------------
    def p2p_buffers(self, mv, rows, dim, device):
        """Receive slabs q_r, k_r, v_r, o_r [P, rows, dim/P] + flags on every rank, mapped into every peer."""
        key = (rows, dim, str(device))
        st = self._p2p.get(key)
        if st is not None:
            return st
        for old in self._p2p.values():
            for base in old["opened"]:
                mv.ipc_close(base)
        self._p2p.clear()
        P, r = self.world, self.rank
        n = rows * dim                                   # elements per slab
        # Every rank takes the same decision at two agreement points (export, map); any failure anywhere switches the
        # whole group to the NCCL exchange (still this library's kernels — there is no CPU or eager fallback).
        slab, info, err = None, None, None
        try:
            slab = torch.zeros(4 * n * 2 + 256, dtype=torch.uint8, device=device)
            handle, off, _ = mv.ipc_export(slab)
            info = (handle, off)
        except Exception as ex:
            err = repr(ex)
        infos = [None] * P
        dist.all_gather_object(infos, info, group=self.group)
        if any(i is None for i in infos):
            self._p2p_error = err or "a peer could not export its buffer"
            self.mode = "nccl"
            return None
        views = [slab[i * n * 2:(i + 1) * n * 2].view(torch.bfloat16).view(P, rows, dim // P) for i in range(4)]
        flags_off = 4 * n * 2
        bases, opened = [], []
        try:
            for i, (h, o) in enumerate(infos):
                if i == r:
                    bases.append(slab.data_ptr())
                else:
                    b = mv.ipc_open(h)
                    opened.append(b)
                    bases.append(b + o)
        except Exception as ex:
            err = repr(ex)
        torch.cuda.synchronize()
        ok = torch.full((1,), 0.0 if err else 1.0, device=device)
        dist.all_reduce(ok, group=self.group)        # also: every rank has zeroed + mapped before anyone writes
        if int(round(ok.item())) != P:
            for b in opened:
                mv.ipc_close(b)
            self._p2p_error = err or "a peer could not map the buffers"
            self.mode = "nccl"
            return None
        st = dict(slab=slab, q_r=views[0], k_r=views[1], v_r=views[2], o_r=views[3], opened=opened,
                  tab=[mv.ptr_table([b + i * n * 2 for b in bases]) for i in range(4)],
                  flags=mv.ptr_table([b + flags_off for b in bases]), local_flags=slab.data_ptr() + flags_off)
        self._p2p[key] = st
        return st

    def side_stream(self, device):
        """(side stream, three reusable events) for the pipelined q / k / v scatter."""
        if self._side is None:
            self._side = (torch.cuda.Stream(device=device), [torch.cuda.Event() for _ in range(3)])
        return self._side

    def p2p_barrier(self, mv, st):
        self.epoch += 1
        mv.sp_barrier(st["flags"], st["local_flags"], self.rank, self.world, self.epoch)

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


def sp_self_attention(mv, grp, ws, rows, bw, cs, kv_len_total, prepare=None, attend=None, qkv_gemm=None):
    """Self-attention core for the local `rows` tokens.  qkv_gemm=None: the fused QKV GEMM output already sits in ws.qkv;
    otherwise qkv_gemm(i) computes column slab i (0 = q, 1 = k, 2 = v) and qkv_gemm(None) all of it — which lets the p2p
    mode hide the NVLink scatter of q and k behind the GEMMs of k and v.

    Leaves the attention output as K-split slabs in the returned tensor [P, rows, C/P] (consumed by
    mv.gemm_ksplit).  `prepare` / `attend` are injection points for the gloo CPU tests of this choreography;
    in production they are the sm_100a kernels."""
    P = grp.world
    C = bw.dim
    Hl = bw.num_heads // P
    if grp.mode == "p2p" and prepare is None and attend is None:
        # fused exchange: scatter stores go straight into the peers' receive buffers over NVLink
        st = grp.p2p_buffers(mv, rows, C, ws.qkv.device)
    if grp.mode == "p2p" and prepare is None and attend is None:
        qkv = ws.qkv[:rows]
        if qkv_gemm is not None and grp.pipeline:
            # q GEMM | k GEMM + scatter(q) | v GEMM + scatter(k) | scatter(v): the scatter kernels (norm + RoPE + peer
            # stores, NVLink-bound) run on a side stream next to the compute-bound GEMMs of the following slab
            main = torch.cuda.current_stream()
            side, ev = grp.side_stream(qkv.device)
            for i in range(3):
                qkv_gemm(i)
                slab = qkv[:, i * C:(i + 1) * C]
                if i < 2:
                    ev[i].record(main)
                    side.wait_event(ev[i])
                    with torch.cuda.stream(side):
                        mv.qkv_prepare_p2p(slab, (bw.g_q, bw.g_k)[i], cs, st["tab"][i], grp.rank, P, 128, bw.eps)
                else:
                    mv.qkv_prepare_p2p(slab, None, None, st["tab"][2], grp.rank, P, 128, bw.eps)
            ev[2].record(side)
            main.wait_event(ev[2])
        else:
            if qkv_gemm is not None:
                qkv_gemm(None)
            # ONE pass over the local QKV rows: q/k RMSNorm + RoPE, and the head groups of q, k and v stored straight
            # into slab `rank` of every destination's receive buffers
            mv.qkv_norm_rope(qkv, bw.g_q, bw.g_k, cs, 128, bw.eps, dst=(st["tab"][0], st["tab"][1], st["tab"][2]),
                             n_dst=P, src_slot=grp.rank)
        grp.p2p_barrier(mv, st)                    # all q/k/v slabs complete everywhere
        L = P * rows
        hd = C // bw.num_heads
        q = st["q_r"].view(L, Hl, hd)
        k = st["k_r"].view(L, Hl, hd)[:kv_len_total]
        v = st["v_r"].view(L, Hl, hd)[:kv_len_total]
        mv.attention_scatter(q, k, v, st["tab"][3], P, grp.rank, rows, Hl * hd)
        grp.p2p_barrier(mv, st)                    # all o slabs complete; also fences q/k/v reuse by the next layer
        return st["o_r"]
    if qkv_gemm is not None:
        qkv_gemm(None)
    b = grp.buffers(rows, C, ws.qkv.device, ws.qkv.dtype)
    qkv = ws.qkv[:rows]
    if prepare is None:
        # one launch: normalise + rotate q, k and write the send layout [dst][row][C/P] of all three tensors
        if "tabs" not in b:
            b["tabs"] = tuple(mv.ptr_table([b[n][d].data_ptr() for d in range(P)]) for n in ("q_s", "k_s", "v_s"))
        mv.qkv_norm_rope(qkv, bw.g_q, bw.g_k, cs, 128, bw.eps, dst=b["tabs"], n_dst=P, src_slot=0)
    else:
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
