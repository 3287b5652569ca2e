"""CPU restatement of the DDPM scheduler on the `LION.sample` route (models/lion.py:24-26,:37-80).

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.

PARITY UNPINNED by the letter of the rule: the arithmetic lives in the third-party dependency
diffusers==0.11.1 (env.yaml:155; `DDPMScheduler`, src/diffusers/schedulers/scheduling_ddpm.py), which is
neither vendored under /root/reference nor installed here, and the reference holds no test or golden
vector for this route.  The published algorithm is restated (epsilon prediction, fp32 linspace betas +
fp32 cumprod, optional clip_sample); the call sites it is anchored on are models/lion.py:24-26
(constructor arguments), :39-40 (set_timesteps(1000) -> t = 999..0), :52-55 / :67-70 (model called with
t+1, `.step(noise_pred, t, x).prev_sample`).
What it IS checked against: the known-answer values of the dependency's OWN test-suite for this class
(diffusers tests, DDPMSchedulerTest: `test_variance` -- _get_variance(0 / 487 / 999) = 0.0 / 0.00979 / 0.02
at 1e-5 -- and `test_full_loop_no_noise` -- 1000 steps of a deterministic dummy model on a deterministic
sample, torch.manual_seed(0) noise, clip_sample=True, fixed_small: sum|x| = 258.9070, mean|x| = 0.3374 at
the test's own 1e-2 / 1e-3).  Those constants are quoted from the published 0.11-line test file from
memory -- this container has no network to re-fetch it -- and the restatement reproduces all five
(tests/test_scheduler_route.py::test_scheduler_restatement_reproduces_diffusers_known_answers).
"""
import torch


def make_scheduler(num_steps=1000, beta_1=1e-4, beta_T=0.02):
    betas = torch.linspace(beta_1, beta_T, num_steps, dtype=torch.float32)
    alphas = 1.0 - betas
    return dict(betas=betas, alphas=alphas, alphas_cumprod=torch.cumprod(alphas, dim=0))


def variance(s, t, variance_type="fixedlarge"):
    ab = s["alphas_cumprod"][t]
    ab_prev = s["alphas_cumprod"][t - 1] if t > 0 else torch.tensor(1.0)
    v = (1 - ab_prev) / (1 - ab) * s["betas"][t]
    if variance_type == "fixed_small":
        v = torch.clamp(v, min=1e-20)
    elif variance_type == "fixed_large":
        v = s["betas"][t]
    return v      # LION's 'fixedlarge' matches no branch: un-clamped posterior variance


def step(s, eps, t, x, noise, variance_type="fixedlarge", clip_sample=False):
    ab = s["alphas_cumprod"][t]
    ab_prev = s["alphas_cumprod"][t - 1] if t > 0 else torch.tensor(1.0)
    bp, bp_prev = 1 - ab, 1 - ab_prev
    x0 = (x - bp ** 0.5 * eps) / ab ** 0.5
    if clip_sample:                      # (LION constructs the scheduler with clip_sample=False, models/lion.py:24-26)
        x0 = torch.clamp(x0, -1, 1)
    c0 = (ab_prev ** 0.5 * s["betas"][t]) / bp
    c1 = s["alphas"][t] ** 0.5 * bp_prev / bp
    prev = c0 * x0 + c1 * x
    if t > 0:
        prev = prev + (variance(s, t, variance_type) ** 0.5) * noise
    return prev


def run(model_fn, s, x_T, noises, variance_type="fixedlarge"):
    """noises[k] is the k-th draw (made at t = T-1 .. 1; none at t = 0)."""
    T = s["betas"].shape[0]
    x = x_T
    for k, t in enumerate(reversed(range(T))):
        eps = model_fn(x, torch.ones(x.shape[0]) * (t + 1))
        x = step(s, eps, t, x, noises[k] if t > 0 else None, variance_type)
    return x
