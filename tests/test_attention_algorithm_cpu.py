"""CPU model of the softmax bookkeeping of the default attention kernel (attention_fwd_k128_kernel, STALE = true in
moviigen1.1_b200/csrc/attention_sm100.cu): 128-key steps, a FIXED reference taken from the exact row max of the first
step, P = bf16(2^(s * scale_log2 - ref)), fp32 row sums, and the overflow guard — a step whose exponentials sum to more
than 2^64 is redone with its exact row max after O and l have been rescaled.  No GPU: this checks that the scheme is
exact (equal to the reference softmax(QK^T)V of wan/modules/attention.py:24-130 as restated by oracle/dit_oracle.py)
for benign scores, for a row max that keeps growing, for jumps beyond the guard, and for a decreasing profile; the
`-m gpu` tests in test_kernels_gpu.py run the same cases through the kernel.
"""
import numpy as np
import pytest
import torch

from oracle import dit_oracle as O

LOG2E = 1.4426950408889634
GUARD = np.float32(2.0 ** 64)


def _bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def fixed_reference_attention(q, k, v, scale, step=128, warp_rows=32):
    """q [Lq, d], k, v [Lk, d] (bf16-representable fp32).  Returns (out [Lq, d], n_redo, n_steps)."""
    Lq, d = q.shape
    Lk = k.shape[0]
    sl2 = np.float32(scale * LOG2E)
    out = np.zeros((Lq, d), np.float32)
    n_redo = 0
    n_kv = (Lk + step - 1) // step
    for r0 in range(0, Lq, warp_rows):           # the redo decision is taken per warp (32 rows)
        rows = slice(r0, min(Lq, r0 + warp_rows))
        S_all = (q[rows] @ k.T).astype(np.float32)   # fp32 accumulation of bf16 products (tcgen05.mma kind::f16)
        n = S_all.shape[0]
        acc = np.zeros((n, d), np.float32)
        l = np.zeros(n, np.float32)
        ref2 = np.zeros(n, np.float32)
        for j in range(n_kv):
            w = min(step, Lk - j * step)
            s = np.full((n, step), -np.inf, np.float32)
            s[:, :w] = S_all[:, j * step:j * step + w]
            vt = np.zeros((step, d), np.float32)
            vt[:w] = v[j * step:j * step + w]
            if j == 0:
                ref2 = (s.max(1) * sl2).astype(np.float32)
            for attempt in range(2):                 # at most ONE redo: with the exact max nothing finite overflows
                with np.errstate(over="ignore", invalid="ignore"):
                    e = np.exp2((s * sl2 - ref2[:, None]).astype(np.float32)).astype(np.float32)
                    row_sum = e.sum(1, dtype=np.float32)
                if attempt == 1 or not np.any(~(row_sum <= GUARD)):   # warp vote; !(x <= t) also catches inf / NaN
                    break
                n_redo += 1
                with np.errstate(invalid="ignore"):
                    up = np.maximum(np.fmax.reduce(s, axis=1) * sl2 - ref2, 0).astype(np.float32)   # max.f32 skips NaN
                alpha = np.exp2(-up).astype(np.float32)
                l *= alpha
                acc *= alpha[:, None]
                ref2 = (ref2 + up).astype(np.float32)
            l += row_sum
            acc += _bf16(e) @ vt                      # P is rounded to bf16 before P.V, the sum is not
        out[rows] = acc / l[:, None]
    return _bf16(out), n_redo, n_kv


def _case(jumps, offset, seed=19, Lq=64):
    g = torch.Generator().manual_seed(seed)
    Lk = 128 * len(jumps) - 37                     # ragged last block
    q = torch.ones(Lq, 1, 128) + 0.05 * torch.randn(Lq, 1, 128, generator=g)
    k = 0.2 * torch.randn(Lk, 1, 128, generator=g)
    for b, c in enumerate(jumps):
        k[128 * b + offset:128 * b + offset + 4] += c
    v = torch.randn(Lk, 1, 128, generator=g)
    return q.bfloat16(), k.bfloat16(), v.bfloat16()


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("offset", [5, 70])
@pytest.mark.parametrize("jumps,expect_redo", [
    ((0.0, 0.3, 0.6, 0.9, 1.2, 1.5, 1.8, 2.1), False),      # max grows by 2^34 in total: below the guard, no redo
    ((0.0, 0.0, 5.0, 5.0, 0.5, 12.0, 12.0, 1.0), True),     # jumps of 2^82 and 2^114: two exact redos per warp
    ((3.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0), False),      # reference far above every later score
])
def test_fixed_reference_softmax_is_exact(jumps, expect_redo, offset):
    q, k, v = _case(jumps, offset)
    ref = O.attention(q, k, v, O.bf16_rt)[:, 0].numpy()
    out, n_redo, _ = fixed_reference_attention(q[:, 0].float().numpy(), k[:, 0].float().numpy(), v[:, 0].float().numpy(),
                                               128 ** -0.5)
    assert np.isfinite(out).all()
    assert _rel(out, ref) <= 5e-3                  # the tolerance of the kernel test (bf16 P, bf16 output)
    assert (n_redo > 0) == expect_redo


def test_fixed_reference_matches_on_random_scores():
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(96, 1, 128, generator=g) * 2).bfloat16()
    k = (torch.randn(333, 1, 128, generator=g) * 2).bfloat16()
    v = torch.randn(333, 1, 128, generator=g).bfloat16()
    ref = O.attention(q, k, v, O.bf16_rt, scale=0.5)[:, 0].numpy()
    out, n_redo, n_steps = fixed_reference_attention(q[:, 0].float().numpy(), k[:, 0].float().numpy(),
                                                     v[:, 0].float().numpy(), 0.5)
    assert n_steps == 3
    assert _rel(out, ref) <= 5e-3


def test_fixed_reference_terminates_on_nan_scores():
    """Non-finite input: one bounded redo, then the NaN propagates into the rows that saw it (never a spin)."""
    g = torch.Generator().manual_seed(5)
    q = torch.randn(40, 1, 128, generator=g).bfloat16()[:, 0].float().numpy()
    k = torch.randn(300, 1, 128, generator=g).bfloat16()[:, 0].float().numpy()
    v = torch.randn(300, 1, 128, generator=g).bfloat16()[:, 0].float().numpy()
    k[200, 7] = np.nan
    with np.errstate(invalid="ignore"):
        out, n_redo, _ = fixed_reference_attention(q, k, v, 128 ** -0.5)
    assert np.isnan(out).all() and n_redo == 2          # one redo per warp of 32 rows (40 rows = 2 warps)
