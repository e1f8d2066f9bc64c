"""TEST INFRASTRUCTURE — CPU restatement of the DDIM scheduler arithmetic the head uses.

PARITY UNPINNED at this boundary: the reference imports
``diffusers.schedulers.DDIMScheduler`` (transfuser_model_v2.py:10) but pins no
version (docs/install.md:7 says ``pip install diffusers einops``; the package is
absent from requirements.txt / environment.yml / setup.py), ships no test or
golden vector for it, and ``diffusers`` is not installed in this image.  This
file restates the published algorithm of
``diffusers/schedulers/scheduling_ddim.py`` (stable across diffusers 0.11-0.3x)
for exactly the call pattern the head makes:

* ctor  ``DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear",
  prediction_type="sample")`` (transfuser_model_v2.py:447-451); all other
  arguments at their defaults: beta_start=1e-4, beta_end=0.02, clip_sample=True,
  clip_sample_range=1.0, set_alpha_to_one=True, steps_offset=0,
  timestep_spacing="leading", thresholding=False.
* ``set_timesteps(1000, device)``  (:584)  =>  step_ratio = 1000 // 1000 = 1, so
  ``prev_timestep = timestep - 1``.
* ``add_noise(original_samples, noise, timesteps)``  (:595-597, :535-539).
* ``step(model_output, timestep, sample).prev_sample``  (:634-636) with eta = 0,
  use_clipped_model_output = False.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline leg may
import this module; the product path never does.
"""
from __future__ import annotations

import torch


class DDIMSchedulerRestated:
    """Subset of ``diffusers.DDIMScheduler`` used by ``TrajectoryHead``."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4,
                 beta_end: float = 0.02, beta_schedule: str = "scaled_linear",
                 prediction_type: str = "sample", clip_sample: bool = True,
                 clip_sample_range: float = 1.0, set_alpha_to_one: bool = True, **_unused):
        assert beta_schedule == "scaled_linear" and prediction_type == "sample"
        # scaled_linear: linspace in sqrt-space, squared; fp32 cumprod (bit-identical
        # table values require torch fp32 here, not a double recomputation).
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = (torch.tensor(1.0) if set_alpha_to_one
                                    else self.alphas_cumprod[0])
        self.num_train_timesteps = num_train_timesteps
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor,
                  timesteps: torch.Tensor) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=original_samples.device,
                                    dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sqrt_ac = ac[timesteps] ** 0.5
        sqrt_1m = (1 - ac[timesteps]) ** 0.5
        sqrt_ac = sqrt_ac.flatten()
        sqrt_1m = sqrt_1m.flatten()
        while sqrt_ac.dim() < original_samples.dim():
            sqrt_ac = sqrt_ac.unsqueeze(-1)
            sqrt_1m = sqrt_1m.unsqueeze(-1)
        return sqrt_ac * original_samples + sqrt_1m * noise

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        ac_t = self.alphas_cumprod[t]
        ac_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - ac_t
        # prediction_type == "sample"
        pred_x0 = model_output
        pred_eps = (sample - ac_t ** 0.5 * pred_x0) / beta_prod_t ** 0.5
        if self.clip_sample:
            pred_x0 = pred_x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        # eta = 0  =>  std_dev_t = 0, no variance noise; use_clipped_model_output=False
        # => pred_eps is NOT recomputed from the clipped x0.
        pred_dir = (1 - ac_prev) ** 0.5 * pred_eps
        prev_sample = ac_prev ** 0.5 * pred_x0 + pred_dir

        class _Out:
            pass
        out = _Out()
        out.prev_sample = prev_sample
        out.pred_original_sample = pred_x0
        return out
