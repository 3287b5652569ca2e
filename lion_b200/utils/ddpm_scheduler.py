"""Minimal DDPM scheduler with the interface `models/lion.py` uses (reference: models/lion.py:24-26
constructs `diffusers.DDPMScheduler(clip_sample=False, beta_start, beta_end, beta_schedule,
num_train_timesteps, variance_type)`, :39-40 `set_timesteps(1000, device)` / `.timesteps`,
:55,:70 `.step(noise_pred, t, x).prev_sample`, :90 `.alphas_cumprod`).

diffusers (pinned 0.11.1 in the reference's env.yaml) is not vendored in the reference and not
installed here, so the published algorithm of its scheduling_ddpm.py is restated -- PARITY
UNPINNED by the letter (SURVEY.md 8c iii; the CPU restatement in oracle/scheduler.py reproduces the
known-answer values of diffusers' own scheduler tests, tests/test_scheduler_route.py): epsilon prediction, fp32 `linspace` betas, `cumprod` in fp32,
  x0 = (x - sqrt(1-abar_t) eps) / sqrt(abar_t)
  prev = sqrt(abar_{t-1}) beta_t / (1-abar_t) * x0 + sqrt(alpha_t) (1-abar_{t-1}) / (1-abar_t) * x
  x' = prev + sqrt(var_t) z  (t > 0; z drawn with torch.randn on the sample's device), x' = prev (t = 0)
  var_t = (1-abar_{t-1})/(1-abar_t) beta_t, clamped at 1e-20 for 'fixed_small', beta_t for 'fixed_large'.
Quirk kept: LION passes cfg.ddpm.model_var_type = 'fixedlarge' (no underscore), which matches none
of the scheduler's variance types, so the un-clamped posterior variance is used.

The arithmetic of step() runs in the library (lion_scheduler_step); there is no CPU path.
"""
from collections import namedtuple

import torch

from .. import _lib as L

SchedulerOutput = namedtuple("SchedulerOutput", ["prev_sample"])


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True):
        if beta_schedule != "linear":
            raise NotImplementedError("lion_b200: only the 'linear' beta schedule of the shipped configs is provided")
        if clip_sample:
            raise NotImplementedError("lion_b200: LION constructs the scheduler with clip_sample=False")
        if variance_type in ("fixed_small_log", "fixed_large_log", "learned", "learned_range"):
            raise NotImplementedError("lion_b200: variance_type %r is not used by LION" % variance_type)
        self.num_train_timesteps = num_train_timesteps
        self.variance_type = variance_type
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(0, num_train_timesteps).flip(0)
        self._tables = {}

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps != self.num_train_timesteps:
            raise NotImplementedError("lion_b200: LION.sample runs all %d training timesteps" % self.num_train_timesteps)
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.arange(0, self.num_train_timesteps).flip(0).to(device)

    def _get_variance(self, t):
        ab = self.alphas_cumprod[t]
        ab_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        variance = (1 - ab_prev) / (1 - ab) * self.betas[t]
        if self.variance_type == "fixed_small":
            variance = torch.clamp(variance, min=1e-20)
        elif self.variance_type == "fixed_large":
            variance = self.betas[t]
        return variance          # any other string (LION's 'fixedlarge'): the un-clamped posterior variance

    def step_tables(self, device):
        """[T][8] fp32 rows {sqrt(1-abar_t), sqrt(abar_t), c0, c1, sqrt(var_t), 0, 0, 0} for lion_scheduler_step."""
        key = str(device)
        if key not in self._tables:
            rows = []
            for t in range(self.num_train_timesteps):
                ab = self.alphas_cumprod[t]
                ab_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
                bp, bp_prev = 1 - ab, 1 - ab_prev
                c0 = (ab_prev ** 0.5 * self.betas[t]) / bp
                c1 = self.alphas[t] ** 0.5 * bp_prev / bp
                sig = self._get_variance(t) ** 0.5 if t > 0 else torch.tensor(0.0)
                z = torch.tensor(0.0)
                rows.append(torch.stack([bp ** 0.5, ab ** 0.5, c0, c1, sig, z, z, z]))
            self._tables[key] = torch.stack(rows).to(torch.float32).contiguous().to(device)
        return self._tables[key]

    @torch.no_grad()
    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        t = int(timestep)
        x = sample.detach().to(torch.float32).contiguous()
        eps = model_output.detach().to(torch.float32).contiguous()
        dev = x.device
        tables = self.step_tables(dev)
        noise = None
        if t > 0:
            noise = torch.randn(eps.shape, generator=generator, device=dev, dtype=eps.dtype)
        step = torch.tensor([t], dtype=torch.int32, device=dev)
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            L.check(L.lib().lion_scheduler_step(L.ptr(x), L.ptr(eps), L.ptr(noise), L.ptr(out), L.ptr(tables), L.ptr(step),
                                                x.numel(), L.stream()), "scheduler_step")
        return SchedulerOutput(prev_sample=out) if return_dict else (out,)
