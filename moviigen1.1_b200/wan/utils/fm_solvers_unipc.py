"""Flow-matching UniPC multistep scheduler, diffusers-free.

Restates the behaviour of the reference's FlowUniPCMultistepScheduler
(wan/utils/fm_solvers_unipc.py:76-742) for the configuration WanT2V.generate uses
(text2video.py:206-213): solver_order 2, solver_type 'bh2', predict_x0, flow_prediction,
lower_order_final, final sigma 0, no thresholding.  Scalar coefficients are computed on the host in
float32 exactly as the reference does (its sigmas live on the CPU, :226-227); only the latent-sized
linear combinations run on the device.  The reference's per-step debug prints (:318,331,690) are
not replicated.
"""
import numpy as np
import torch


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction", shift=1.0,
                 use_dynamic_shifting=False, thresholding=False, predict_x0=True, solver_type="bh2",
                 lower_order_final=True, disable_corrector=(), final_sigmas_type="zero", **unused):
        if prediction_type != "flow_prediction" or not predict_x0 or thresholding or use_dynamic_shifting:
            raise NotImplementedError("only the flow_prediction / predict_x0 configuration used by WanT2V is supported")
        if solver_type not in ("bh1", "bh2"):
            solver_type = "bh2"
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.solver_type = solver_type
        self.lower_order_final = lower_order_final
        self.disable_corrector = list(disable_corrector)
        self.final_sigmas_type = final_sigmas_type
        self.shift = shift
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas
        self.timesteps = sigmas * num_train_timesteps
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self.num_inference_steps = None
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.solver_order
        self.timestep_list = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = 1
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """:160-227: sigma_i = shift*s/(1+(shift-1)s), s = linspace(sigma_max, sigma_min, N+1)[:-1]; t = int64(1000 sigma)."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        if self.final_sigmas_type != "zero":
            raise NotImplementedError("final_sigmas_type must be 'zero'")
        timesteps = sigmas * self.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)  # stays on the CPU
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._timesteps_host = [int(v) for v in self.timesteps.tolist()]
        self._reset()

    # -- scalar helpers (float32 torch scalars on the CPU, as in the reference) ------------------------
    def _lam(self, idx):
        s = self.sigmas[idx]
        return torch.log(1 - s) - torch.log(s)

    def _coeffs(self, idx_t, idx_s0, prev_idxs, order):
        """Common part of :413-450 and :552-592: returns (sigma_t/sigma_s0, alpha_t, h_phi_1, B_h, rks, R, b)."""
        sigma_t, sigma_s0 = self.sigmas[idx_t], self.sigmas[idx_s0]
        alpha_t = 1 - sigma_t
        h = self._lam(idx_t) - self._lam(idx_s0)
        rks = [((self._lam(i) - self._lam(idx_s0)) / h) for i in prev_idxs] + [1.0]
        rks_t = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks_t, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return sigma_t / sigma_s0, alpha_t, h_phi_1, B_h, rks, torch.stack(R), torch.tensor(b)

    def _predict(self, sample, order):
        """multistep_uni_p_bh_update (:351-485)."""
        i = self._step_index
        m0 = self.model_outputs[-1]
        prev = [i - k for k in range(1, order)]
        ratio, alpha_t, h_phi_1, B_h, rks, R, b = self._coeffs(i + 1, i, prev, order)
        x_t = float(ratio) * sample - float(alpha_t * h_phi_1) * m0
        if order > 1:
            if order == 2:
                rhos = [0.5]
            else:
                rhos = torch.linalg.solve(R[:-1, :-1], b[:-1]).tolist()
            res = None
            for k in range(1, order):
                d1 = (self.model_outputs[-(k + 1)] - m0) / float(rks[k - 1])
                res = d1 * float(rhos[k - 1]) if res is None else res + d1 * float(rhos[k - 1])
            x_t = x_t - float(alpha_t * B_h) * res
        return x_t.to(sample.dtype)

    def _correct(self, model_t, last_sample, order):
        """multistep_uni_c_bh_update (:487-627)."""
        i = self._step_index
        m0 = self.model_outputs[-1]
        prev = [i - (k + 1) for k in range(1, order)]
        ratio, alpha_t, h_phi_1, B_h, rks, R, b = self._coeffs(i, i - 1, prev, order)
        x_t = float(ratio) * last_sample - float(alpha_t * h_phi_1) * m0
        rhos = [0.5] if order == 1 else torch.linalg.solve(R, b).tolist()
        res = (model_t - m0) * float(rhos[-1])
        for k in range(1, order):
            d1 = (self.model_outputs[-(k + 1)] - m0) / float(rks[k - 1])
            res = res + d1 * float(rhos[k - 1])
        return (x_t - float(alpha_t * B_h) * res).to(last_sample.dtype)

    def _init_step_index(self, timestep):
        t = int(timestep)
        idx = [k for k, v in enumerate(self._timesteps_host) if v == t]
        self._step_index = idx[1] if len(idx) > 1 else idx[0]                 # :629-640

    # -- fused device path: CFG + convert_model_output + UniC + UniP in ONE sm_100a kernel (mv_unipc_cfg_step) -------
    def _predict_coefs(self, order):
        i = self._step_index
        prev = [i - k for k in range(1, order)]
        ratio, alpha_t, h_phi_1, B_h, rks, R, b = self._coeffs(i + 1, i, prev, order)
        rhos = []
        if order == 2:
            rhos = [0.5]
        elif order > 2:
            rhos = torch.linalg.solve(R[:-1, :-1], b[:-1]).tolist()
        return (float(ratio), float(alpha_t * h_phi_1), float(alpha_t * B_h), [float(r) for r in rks[:order - 1]],
                [float(r) for r in rhos[:order - 1]])

    def _correct_coefs(self, order):
        i = self._step_index
        prev = [i - (k + 1) for k in range(1, order)]
        ratio, alpha_t, h_phi_1, B_h, rks, R, b = self._coeffs(i, i - 1, prev, order)
        rhos = [0.5] if order == 1 else torch.linalg.solve(R, b).tolist()
        return (float(ratio), float(alpha_t * h_phi_1), float(alpha_t * B_h), float(rhos[-1]),
                [float(r) for r in rks[:order - 1]], [float(r) for r in rhos[:order - 1]])

    def step_cfg(self, cond, uncond, guide_scale, timestep, sample):
        """The latent-sized arithmetic of one sampling step (text2video.py:245-254 + step(), :656-742) as ONE kernel:
        noise = uncond + guide_scale*(cond - uncond), x0, corrector, predictor.  CUDA fp32 tensors only; the host
        computes the scalar coefficients exactly as step() does.  Returns (prev_sample, x0)."""
        import movii_b200 as mv
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        if self.solver_order > 3:
            raise NotImplementedError("mv_unipc_cfg_step supports solver_order <= 3")
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        shape = sample.shape
        f = lambda t: None if t is None else t.reshape(-1)  # noqa: E731
        cond, uncond, sample = (t.contiguous() for t in (cond, uncond, sample))
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        coef = [0.0] * mv.UNIPC_NCOEF
        coef[0], coef[1], coef[2] = float(guide_scale), float(self.sigmas[i]), 1.0 if use_corrector else 0.0
        if use_corrector:
            co = self.this_order
            ratio, ca, cb, rho_last, rks, rhos = self._correct_coefs(co)
            coef[3:8] = [float(co), ratio, ca, cb, rho_last]
            for k, (rk, rho) in enumerate(zip(rks, rhos)):
                coef[8 + k], coef[10 + k] = rk, rho
        order = min(self.solver_order, len(self._timesteps_host) - i) if self.lower_order_final else self.solver_order
        this_order = min(order, self.lower_order_nums + 1)
        ratio, pa, pb, rks, rhos = self._predict_coefs(this_order)
        coef[12:16] = [float(this_order), ratio, pa, pb]
        for k, (rk, rho) in enumerate(zip(rks, rhos)):
            coef[16 + k], coef[18 + k] = rk, rho
        hist = [self.model_outputs[-1 - k] if k < self.solver_order else None for k in range(3)]
        x0 = torch.empty_like(sample)          # fresh tensors, like the reference (callers may keep them)
        s_out = torch.empty_like(sample)
        prev = torch.empty_like(sample)
        mv.unipc_cfg_step(f(cond), f(uncond), f(sample), f(self.last_sample), [f(h) for h in hist], coef, f(x0), f(s_out),
                          f(prev))
        for k in range(self.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
            self.timestep_list[k] = self.timestep_list[k + 1]
        self.model_outputs[-1] = x0.view(shape)
        self.timestep_list[-1] = timestep
        self.this_order = this_order
        self.last_sample = s_out.view(shape)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return prev.view(shape), self.model_outputs[-1]

    def step(self, model_output, timestep, sample, return_dict=True, generator=None):
        """:656-742.  Returns (prev_sample, x0_pred) when return_dict=False.  CUDA fp32 tensors run through the fused
        kernel (guide 0 makes the CFG stage the identity: m + 0*(m - m)); CPU tensors (host-side tests against the
        reference trajectories) through the equivalent torch expressions below."""
        if model_output.is_cuda and model_output.dtype == torch.float32 and sample.dtype == torch.float32 \
                and self.solver_order <= 3:
            prev, x0 = self.step_cfg(model_output, model_output, 0.0, timestep, sample)
            return (prev, x0) if not return_dict else {"prev_sample": prev}
        if self.num_inference_steps is None:
            raise ValueError("run set_timesteps first")
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        x0 = sample - float(self.sigmas[i]) * model_output                    # convert_model_output :319-332
        if use_corrector:
            sample = self._correct(x0, self.last_sample, self.this_order)
        for k in range(self.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
            self.timestep_list[k] = self.timestep_list[k + 1]
        self.model_outputs[-1] = x0
        self.timestep_list[-1] = timestep
        order = min(self.solver_order, len(self._timesteps_host) - i) if self.lower_order_final else self.solver_order
        self.this_order = min(order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev_sample = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev_sample, x0)
        return {"prev_sample": prev_sample}

    def scale_model_input(self, sample, *a, **k):
        return sample
