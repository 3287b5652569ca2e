"""CPU restatement of the discrete DDPM sampler on LION's hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.

  make_schedule        utils/diffusion.py:52-53 (linear) + utils/diffusion_pvd.py:118-142
  ddpm_step            utils/diffusion_pvd.py:475-486 (get_q_posterior_mean) + :283-296
  run_denoising        utils/diffusion_pvd.py:223-303 (run_denoising_diffusion)
  sample_2prior        trainers/train_2prior.py:49-127 (generate_samples_vada_2prior,
                       ode_sample=0, ddim_step=0) + models/vae_adain.py:301-333 (sample)

Parity status: pinned by tests/golden/schedule.npz and ddpm10.npz (made by running the
reference's own DiffusionDiscretized on CPU, tests/golden/make_golden.py).
"""
import numpy as np
import torch


def make_schedule(num_steps=1000, beta_1=1e-4, beta_T=0.02, mode="linear"):
    """float64 numpy constants cast to fp32 torch (diffusion_pvd.py:118-140)."""
    if mode != "linear":
        raise NotImplementedError(mode)  # every shipped prior config uses 'linear'
    betas = torch.linspace(beta_1, beta_T, num_steps, dtype=torch.float64).numpy()
    alphas = 1.0 - betas
    alpha_bars = np.cumprod(alphas)
    f = lambda a: torch.from_numpy(a).float()
    return dict(betas=f(betas), alphas=f(alphas), alpha_bars=f(alpha_bars))


def ddpm_step(sched, x, eps, t, noise, temp=1.0):
    """One ancestral step t -> t-1 (t in 0..T-1 as in the reference's loop variable).

    t>0: mean = 1/sqrt(alpha_t) * (x - beta_t*eps/sqrt(1-abar_t)); x' = mean + exp(0.5*log beta_t)*z*temp
    t=0: x' = 1/sqrt(abar_0) * (x - sqrt(1-abar_0)*eps), no noise."""
    if t == 0:
        ab = sched["alpha_bars"][0]
        return 1.0 / torch.sqrt(ab) * (x - torch.sqrt(1.0 - ab) * eps)
    mean = 1.0 / torch.sqrt(sched["alphas"][t]) * (
        x - sched["betas"][t] * eps / torch.sqrt(1.0 - sched["alpha_bars"][t]))
    log_scale = 0.5 * torch.log(sched["betas"][t])
    return mean + torch.exp(log_scale) * noise * temp


def run_denoising(model_fn, sched, x_T, noises, temp=1.0, steps=None):
    """model_fn(x, t_float[B]) -> eps.  noises[t] is consumed at loop variable t (the
    reference indexes given_noise[1][t], diffusion_pvd.py:289).  `steps` optionally restricts
    the loop to the last `steps` values of t (teacher-forced short horizons)."""
    T = sched["betas"].shape[0]
    x = x_T
    traj = []
    ts = list(reversed(range(T)))
    if steps is not None:
        ts = ts[-steps:]
    for t in ts:
        tt = torch.ones(x.shape[0]) * (t + 1)              # the model sees 1..T (:257-258)
        eps = model_fn(x, tt)
        x = ddpm_step(sched, x, eps, t, noises[t] if t > 0 else None, temp)
        traj.append(x)
    return x, traj


def sample_2prior(global_fn, local_fn, decoder_fn, sched, noise_g, noise_l):
    """generate_samples_vada_2prior: global prior loop -> style (style_mlp is '' => identity,
    vae_adain.py:120-127) -> local prior loop conditioned on it -> decoder."""
    z_g, _ = run_denoising(lambda x, t: global_fn(x, t), sched, noise_g[0], noise_g[1])
    style = z_g.reshape(z_g.shape[0], -1)
    z_l, _ = run_denoising(lambda x, t: local_fn(x, t, style), sched, noise_l[0], noise_l[1])
    pts = decoder_fn(z_l.reshape(z_l.shape[0], -1), style)
    return pts, z_g, z_l
