"""Flow-matching DPM-Solver++ (2M, midpoint) scheduler, diffusers-free.

Restates what WanT2V.generate uses of the reference's wan/utils/fm_solvers.py when `sample_solver='dpm++'`
(text2video.py:214-223): get_sampling_sigmas (:22-26), retrieve_timesteps (:29-66) and
FlowDPMSolverMultistepScheduler (:69-858) in its shipped configuration — solver_order 2, algorithm 'dpmsolver++',
solver_type 'midpoint', flow_prediction, lower_order_final, final sigma 0.  Scalars are float32 CPU tensors exactly
as in the reference; only latent-sized linear combinations touch the device.
"""
import numpy as np
import torch

__all__ = ["FlowDPMSolverMultistepScheduler", "get_sampling_sigmas", "retrieve_timesteps"]


def get_sampling_sigmas(sampling_steps, shift):
    """:22-26."""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """:29-66 for the two call forms the pipeline uses."""
    if timesteps is not None:
        raise ValueError("custom timesteps are not supported; pass sigmas")
    if sigmas is not None:
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, len(scheduler.timesteps)


class FlowDPMSolverMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction", shift=1.0,
                 use_dynamic_shifting=False, thresholding=False, algorithm_type="dpmsolver++", solver_type="midpoint",
                 lower_order_final=True, euler_at_final=False, final_sigmas_type="zero", **unused):
        if (prediction_type != "flow_prediction" or thresholding or use_dynamic_shifting or solver_order != 2 or
                algorithm_type != "dpmsolver++" or solver_type != "midpoint" or final_sigmas_type != "zero"):
            raise NotImplementedError("only the dpmsolver++ / midpoint / order-2 / flow_prediction configuration used by "
                                      "WanT2V is supported")
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.lower_order_final, self.euler_at_final = lower_order_final, euler_at_final
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()
        self.num_inference_steps = None
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """:226-290."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.num_train_timesteps
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = [int(v) for v in self.timesteps.tolist()]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self._step_index = None

    def _lam(self, i):
        s = self.sigmas[i]
        return torch.log(1 - s) - torch.log(s)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict=True):
        """:706-798."""
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        if self._step_index is None:
            t = int(timestep)
            idx = [k for k, v in enumerate(self._timesteps_host) if v == t]
            self._step_index = idx[1] if len(idx) > 1 else idx[0]
        i, n = self._step_index, len(self._timesteps_host)
        lower_final = (i == n - 1)                              # final_sigmas_type == "zero" (:739-742)
        lower_second = (i == n - 2) and self.lower_order_final and n < 15
        m0 = sample - float(self.sigmas[i]) * model_output      # convert_model_output, flow_prediction (:392-394)
        self.model_outputs = [self.model_outputs[1], m0]
        sample = sample.to(torch.float32)
        sigma_t, sigma_s0 = self.sigmas[i + 1], self.sigmas[i]
        alpha_t = 1 - sigma_t
        h = self._lam(i + 1) - self._lam(i)
        c_x = float(sigma_t / sigma_s0)
        c_d0 = float(alpha_t * (torch.exp(-h) - 1.0))
        if self.lower_order_nums < 1 or lower_final:
            prev = c_x * sample - c_d0 * m0                     # first order (:455-459)
        else:                                                   # second order multistep, midpoint (:520-549)
            del lower_second                                    # solver_order == 2: same branch either way
            h0 = self._lam(i) - self._lam(i - 1)
            r0 = h0 / h
            d1 = float(1.0 / r0) * (m0 - self.model_outputs[0])
            prev = c_x * sample - c_d0 * m0 - 0.5 * c_d0 * d1
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        prev = prev.to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return {"prev_sample": prev}

    def scale_model_input(self, sample, *a, **k):
        return sample
