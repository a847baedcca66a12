"""CPU (gloo, world_size 2 and 4): the Ulysses choreography of wan/distributed/ulysses.py — head scatter, q/k/v
all-to-all, head-parallel attention, return all-to-all, K-split consumption — reproduces single-rank attention.
The CUDA kernels are replaced by torch stand-ins here (no GPU); the -m gpu multi-GPU test checks the real path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dit_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _BW:
    def __init__(self, dim, nh):
        self.dim, self.num_heads, self.eps = dim, nh, 1e-6
        g = torch.Generator().manual_seed(1)
        self.g_q = 1 + 0.1 * torch.randn(dim, generator=g)
        self.g_k = 1 + 0.1 * torch.randn(dim, generator=g)


class _WS:
    pass


def _torch_prepare(P, hd):
    def prepare(x, w, cs, out):
        M, C = x.shape
        y = x.float()
        if w is not None:
            y = O.rms_norm(x, w, 1e-6, O.bf16_rt)
        if cs is not None:
            y = O.rope_apply(y.view(M, C // hd, hd), cs).reshape(M, C)
        y = O.bf16_rt(y).to(x.dtype)
        out.copy_(y.view(M, P, C // P).transpose(0, 1))
    return prepare


def _attend(q, k, v, o):
    o.copy_(O.attention(q, k, v, O.bf16_rt).to(o.dtype))


def _worker(rank, world, port, L, dim, nh, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wan.distributed.ulysses import UlyssesGroup, sp_self_attention, token_range
        hd = dim // nh
        g = torch.Generator().manual_seed(5)
        qkv_full = torch.randn(L, 3 * dim, generator=g).bfloat16()
        ang = O.rope_table((1, 1, L), hd, L)
        grp = UlyssesGroup()
        start, rows = token_range(L, world, rank)
        ws = _WS()
        ws.qkv = qkv_full[start:start + rows].clone()
        bw = _BW(dim, nh)
        slabs = sp_self_attention(None, grp, ws, rows, bw, ang[start:start + rows], L,
                                  prepare=_torch_prepare(world, hd), attend=_attend)
        # K-split slabs [P(src), rows, C/P] -> [rows, C]
        local = slabs.transpose(0, 1).reshape(rows, dim)
        full = grp.all_gather_rows(local.contiguous())
        if rank == 0:
            ret["out"] = full.float()
            ret["qkv"] = qkv_full
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ulysses_equals_single_rank(world):
    L, dim, nh = 48, 256, 8
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, L, dim, nh, ret), nprocs=world, join=True)
    out, qkv = ret["out"], ret["qkv"]
    hd = dim // nh
    bw = _BW(dim, nh)
    ang = O.rope_table((1, 1, L), hd, L)
    q = O.bf16_rt(O.rope_apply(O.rms_norm(qkv[:, :dim], bw.g_q, 1e-6, O.bf16_rt).view(L, nh, hd), ang))
    k = O.bf16_rt(O.rope_apply(O.rms_norm(qkv[:, dim:2 * dim], bw.g_k, 1e-6, O.bf16_rt).view(L, nh, hd), ang))
    v = qkv[:, 2 * dim:].float().view(L, nh, hd)
    ref = O.attention(q, k, v, O.bf16_rt).reshape(L, dim)
    assert torch.equal(out, ref), (out - ref).abs().max()


def test_token_range():
    from wan.distributed.ulysses import token_range
    assert [token_range(131040, 8, r) for r in (0, 7)] == [(0, 16380), (114660, 16380)]
    with pytest.raises(ValueError):
        token_range(10, 4, 0)
