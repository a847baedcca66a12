"""CPU: the host half of the fused CFG + UniPC step (FlowUniPCMultistepScheduler.step_cfg) — coefficient packing,
history rotation, orders — against the REFERENCE trajectory (tests/golden/unipc.pt, minted from the reference's
scheduler).  The sm_100a kernel itself (mv_unipc_cfg_step) is replaced by a line-by-line torch emulation here; the
-m gpu test (tests/test_model_gpu.py::test_unipc_fused_kernel_matches_reference) runs the real kernel on the same data."""
import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def emulate_kernel(cond, uncond, sample, last_sample, hist, coef, x0_out, sample_out, prev_out):
    """csrc/rowops.cu::unipc_cfg_step_kernel restated with torch fp32 element-wise ops in the same order."""
    c = [torch.tensor(v, dtype=torch.float32) for v in coef]
    f = lambda v: c[v]  # noqa: E731
    noise = uncond + f(0) * (cond - uncond)
    x0 = sample - f(1) * noise
    s = sample
    m0, m1, m2 = hist
    if coef[2] != 0.0:
        co = int(coef[3])
        xt = f(4) * last_sample - f(5) * m0
        res = (x0 - m0) * f(7)
        if co >= 2:
            res = res + ((m1 - m0) / f(8)) * f(10)
        if co >= 3:
            res = res + ((m2 - m0) / f(9)) * f(11)
        s = xt - f(6) * res
    po = int(coef[12])
    xn = f(13) * s - f(14) * x0
    if po >= 2:
        res = ((m0 - x0) / f(16)) * f(18)
        if po >= 3:
            res = res + ((m1 - x0) / f(17)) * f(19)
        xn = xn - f(15) * res
    x0_out.copy_(x0)
    sample_out.copy_(s)
    prev_out.copy_(xn)
    return prev_out


def _run(monkeypatch, guide, solver_order=2):
    import movii_b200 as mv
    from wan.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    monkeypatch.setattr(mv, "unipc_cfg_step", emulate_kernel)
    gold = torch.load(os.path.join(GOLD, "unipc.pt"), weights_only=False)
    worst = 0.0
    for steps, rec in gold.items():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False,
                                        solver_order=solver_order)
        s.set_timesteps(steps, device="cpu", shift=5.0)
        ref_s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False,
                                            solver_order=solver_order)
        ref_s.set_timesteps(steps, device="cpu", shift=5.0)
        x = rec["traj"][0]
        xr = x.clone()
        for i, t in enumerate(s.timesteps):
            v = rec["model_outputs"][i]
            if guide is None:                       # the scheduler-only API: cond == uncond == model output, guide 0
                x = s.step_cfg(v, v, 0.0, t, x)[0]
                noise = v
            else:                                   # a real CFG pair whose combination is v up to rounding
                un = v - 0.3 * torch.cos(5.0 * v)
                x = s.step_cfg(v + (guide - 1) * (v - un) / guide, un, guide, t, x)[0]
                noise = un + guide * ((v + (guide - 1) * (v - un) / guide) - un)
            xr = ref_s.step(noise, t, xr, return_dict=False)[0]      # the torch restatement pinned to the reference
            assert torch.equal(x, xr), (steps, i, (x - xr).abs().max().item())
            if guide is None and solver_order == 2:
                ref = rec["traj"][i + 1]
                worst = max(worst, ((x - ref).abs().max() / max(1.0, ref.abs().max().item())).item())
    return worst


def test_fused_step_equals_reference_trajectory(monkeypatch):
    assert _run(monkeypatch, None) <= 1e-5


def test_fused_step_with_guidance_equals_unfused(monkeypatch):
    _run(monkeypatch, 5.0)


def test_fused_step_order3(monkeypatch):
    _run(monkeypatch, None, solver_order=3)
