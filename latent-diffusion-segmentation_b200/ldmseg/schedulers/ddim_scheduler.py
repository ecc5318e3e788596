"""DDIM noise scheduler -- drop-in for `ldmseg.schedulers.DDIMNoiseScheduler`
(reference: /root/reference/ldmseg/schedulers/ddim_scheduler.py:26-291).

Same constructor, attributes and methods.  `step` runs as ONE fused CUDA kernel
(ldmseg_ddim_step / ldmseg_ddim_step_indexed in include/ldmseg_b200.h) instead of ~10 element-wise
launches, and when the timestep is a CUDA tensor the alpha table is indexed on the device, so the
three device->host synchronisations the reference pays per call (ddim_scheduler.py:234-235)
disappear.  There is no CPU fallback for `step`: CPU tensors raise.
The beta / alpha tables, the timestep grid and the loss weights are host logic (numpy / torch CPU).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import numpy as np
import torch

from ldmseg.utils import OutputDict
from ldmseg import _native as nat


class DDIMNoiseSchedulerOutput(OutputDict):
    prev_sample: torch.FloatTensor
    pred_original_sample: Optional[torch.FloatTensor] = None


_PTYPE = {"epsilon": 0, "sample": 1, "v_prediction": 2}


class DDIMNoiseScheduler(object):
    def __init__(
        self,
        num_train_timesteps: int = 1000,
        beta_start: float = 0.0001,
        beta_end: float = 0.02,
        beta_schedule: str = "linear",
        clip_sample: bool = True,
        set_alpha_to_one: bool = True,
        steps_offset: int = 0,
        prediction_type: str = "epsilon",
        thresholding: bool = False,
        dynamic_thresholding_ratio: float = 0.995,
        clip_sample_range: float = 1.0,
        sample_max_value: float = 1.0,
        weight: str = "none",
        max_snr: float = 5.0,
        device: Union[str, torch.device] = None,
        verbose: bool = True,
    ):
        n = num_train_timesteps
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = self.get_betas_for_alpha_bar(n)
        elif beta_schedule == "sigmoid":
            self.betas = torch.sigmoid(torch.linspace(-6, 6, n)) * (beta_end - beta_start) + beta_start
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")

        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

        self.compute_loss_weights(mode=weight, max_snr=max_snr)
        self.weights = self.weights.to(device)

        self.num_train_timesteps = n
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, n)[::-1].copy().astype(np.int64))
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.prediction_type = prediction_type
        self.thresholding = thresholding
        self.dynamic_thresholding_ratio = dynamic_thresholding_ratio
        self.steps_offset = steps_offset
        self.beta_schedule = beta_schedule
        self.beta_start = beta_start
        self.beta_end = beta_end
        self.init_noise_sigma = 1.0
        self.verbose = verbose
        self._acp_dev = {}  # device -> alphas_cumprod copy used by the indexed kernel

    # ------------------------------------------------------------------ host logic
    def compute_loss_weights(self, mode="max_clamp_snr", max_snr=5.0):
        assert mode in ["inverse_log_snr", "max_clamp_snr", "linear", "fixed", "none"]
        self.weight_mode = mode
        snr = self.alphas_cumprod / (1 - self.alphas_cumprod)
        if mode == "inverse_log_snr":
            self.weights = torch.log(1.0 / snr).clamp(min=1)
            self.weights /= self.weights[-1]
        elif mode == "max_clamp_snr":
            self.weights = snr.clamp(max=max_snr) / snr
        elif mode == "fixed":
            self.weights = snr
            self.weights[: len(self.weights) // 4] = 0.1
        elif mode == "linear":
            self.weights = torch.arange(1, len(snr) + 1) / len(snr)
        else:
            self.weights = torch.ones_like(snr)

    def set_timesteps_inference(self, num_inference_steps: int, device=None, tmin: int = 0):
        """Timestep grid with the final step always included; note that this overwrites
        `steps_offset` with ratio - 1 exactly like the reference (ddim_scheduler.py:126-127)."""
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        self.steps_offset = step_ratio - 1
        grid = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(grid).to(device)
        self.timesteps += self.steps_offset
        self.timesteps = self.timesteps[self.timesteps >= tmin]

    def move_timesteps_to(self, device):
        self.timesteps = self.timesteps.to(device)

    def get_betas_for_alpha_bar(self, num_diffusion_timesteps, max_beta=0.999) -> torch.Tensor:
        def alpha_bar(s):
            return math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2

        t = num_diffusion_timesteps
        return torch.tensor([min(1 - alpha_bar((i + 1) / t) / alpha_bar(i / t), max_beta) for i in range(t)],
                            dtype=torch.float32)

    def _gather(self, timesteps, like: torch.Tensor):
        acp = self.alphas_cumprod.to(device=like.device, dtype=like.dtype)
        a = acp[timesteps.to(like.device)].flatten()
        view = (-1,) + (1,) * (like.dim() - 1)
        return (a ** 0.5).view(view), ((1 - a) ** 0.5).view(view)

    def _acp_on(self, dev):
        acp = self._acp_dev.get(dev)
        if acp is None:
            acp = self.alphas_cumprod.to(device=dev, dtype=torch.float32).contiguous()
            self._acp_dev[dev] = acp
        return acp

    def _mix_cuda(self, x, noise, timesteps, scale, mode):
        """CUDA tensors: one fused kernel, per-sample timesteps gathered on the device (ldmseg_noise_mix)."""
        xs, nz = x.contiguous().float(), noise.contiguous().float()
        nb = xs.shape[0]
        t = timesteps.to(device=xs.device, dtype=torch.int64).reshape(-1)
        if t.numel() == 1 and nb > 1:
            t = t.expand(nb)
        t = t.contiguous()
        if t.numel() != nb:
            raise RuntimeError("add_noise / remove_noise: one timestep per sample")
        out = torch.empty_like(xs)
        nat.noise_mix(xs, nz, t, self._acp_on(xs.device), nb, xs[0].numel(), float(scale), mode, out)
        return out if x.dtype == torch.float32 else out.to(x.dtype)

    def add_noise(self, original_samples, noise, timesteps, scale: float = 1.0,
                  mask_noise_perc: Optional[float] = None):
        if mask_noise_perc is not None:
            noise *= torch.rand_like(original_samples) < mask_noise_perc
        if original_samples.is_cuda:
            return self._mix_cuda(original_samples, noise, timesteps, scale, 0)
        sa, sb = self._gather(timesteps, original_samples)
        return sa * scale * original_samples + sb * noise

    @torch.no_grad()
    def remove_noise(self, noisy_samples, noise, timesteps, scale: float = 1.0):
        if noisy_samples.is_cuda:
            return self._mix_cuda(noisy_samples, noise, timesteps, scale, 1)
        sa, sb = self._gather(timesteps, noisy_samples)
        return (noisy_samples - sb * noise) / (sa * scale)

    # ------------------------------------------------------------------ the hot call
    def step(self, model_output, timestep, sample, use_clipped_model_output: bool = False
             ) -> DDIMNoiseSchedulerOutput:
        if self.prediction_type not in _PTYPE:
            raise NotImplementedError
        if self.thresholding:
            raise NotImplementedError
        nat.require_cuda(model_output, sample)
        mo = model_output.contiguous()
        xs = sample.contiguous()
        if mo.dtype != torch.float32 or xs.dtype != torch.float32:
            mo, xs = mo.float(), xs.float()
        prev = torch.empty_like(xs)
        x0 = torch.empty_like(xs)
        ratio = self.num_train_timesteps // self.num_inference_steps
        ptype = _PTYPE[self.prediction_type]
        if torch.is_tensor(timestep) and timestep.is_cuda:
            acp = self._acp_on(xs.device)
            t = timestep.reshape(-1)[:1].to(torch.int64)
            nat.ddim_step_indexed(mo, xs, t, acp, ratio, float(self.final_alpha_cumprod), ptype,
                                  self.clip_sample, float(self.clip_sample_range),
                                  use_clipped_model_output, prev, x0)
        else:
            t = int(timestep)
            tp = t - ratio
            a_t = float(self.alphas_cumprod[t])
            a_p = float(self.alphas_cumprod[tp]) if tp >= 0 else float(self.final_alpha_cumprod)
            nat.ddim_step(mo, xs, a_t, a_p, ptype, self.clip_sample, float(self.clip_sample_range),
                          use_clipped_model_output, prev, x0)
        return DDIMNoiseSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    # ---- extension (not in the reference): ancestral / DDPM step = DDIM with eta = 1
    def step_ddpm(self, model_output, timestep, sample, noise) -> DDIMNoiseSchedulerOutput:
        nat.require_cuda(model_output, sample, noise)
        t = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        tp = t - ratio
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[tp]) if tp >= 0 else float(self.final_alpha_cumprod)
        var = max((1 - a_p) / (1 - a_t) * (1 - a_t / a_p), 0.0)
        prev = torch.empty_like(sample)
        x0 = torch.empty_like(sample)
        nat.ddim_step(model_output.contiguous(), sample.contiguous(), a_t, a_p, 0, False, 1.0, False, prev,
                      x0, sigma=var ** 0.5, noise=noise.contiguous())
        return DDIMNoiseSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def __str__(self) -> str:
        w = self.weights if self.verbose else "VerboseDisabled"
        return (f"DDIMScheduler(num_inference_steps={self.num_inference_steps}, "
                f"num_train_timesteps={self.num_train_timesteps}, prediction_type={self.prediction_type}, "
                f"beta_start={self.beta_start}, beta_end={self.beta_end}, beta_schedule={self.beta_schedule}, "
                f"clip_sample={self.clip_sample}, clip_sample_range={self.clip_sample_range}, "
                f"thresholding={self.thresholding}, "
                f"dynamic_thresholding_ratio={self.dynamic_thresholding_ratio}, "
                f"steps_offset={self.steps_offset}, weight_mode={self.weight_mode}, weights={w})")

    def __repr__(self) -> str:
        return self.__str__()

    def __len__(self) -> int:
        return self.num_train_timesteps
