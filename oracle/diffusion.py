"""CPU restatement of the discrete DDPM sampler on LION's hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.

  make_schedule        utils/diffusion.py:52-53 (linear) + utils/diffusion_pvd.py:118-142
  ddpm_step            utils/diffusion_pvd.py:475-486 (get_q_posterior_mean) + :283-296
  run_denoising        utils/diffusion_pvd.py:223-303 (run_denoising_diffusion)
  ddim_taus / ddim_coeffs / ddim_step / run_ddim
                       utils/diffusion_pvd.py:389-473 (run_ddim; the DDIM route of
                       generate_samples_vada_2prior, trainers/train_2prior.py:87-93)
  sample_2prior        trainers/train_2prior.py:49-127 (generate_samples_vada_2prior,
                       ode_sample=0, ddim_step=0) + models/vae_adain.py:301-333 (sample)

Parity status: pinned by tests/golden/schedule.npz, ddpm10.npz and ddim5.npz (made by running
the reference's own DiffusionDiscretized on CPU, tests/golden/make_golden.py and
make_golden_ddim.py).
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


# ------------------------------------------------------------------------------------------
# DDIM (utils/diffusion_pvd.py:389-473)
# ------------------------------------------------------------------------------------------
def ddim_taus(T, S, skip_type="uniform"):
    """Sub-sequence of loop variables, descending (:411-423).  'uniform': floor(i*(T-1)/(S-1));
    'quad': int(linspace(0, sqrt(0.8 T), S)**2)."""
    if skip_type == "uniform":
        c = (T - 1.0) / (S - 1.0)
        taus = [int(np.floor(i * c)) for i in range(S)]
    elif skip_type == "quad":
        taus = [int(s) for s in list(np.linspace(0, np.sqrt(T * 0.8), S) ** 2)]
    else:
        raise NotImplementedError(skip_type)
    return sorted(taus, reverse=True)


def ddim_coeffs(sched, taus, i, kappa=1.0):
    """fp32 0-dim tensors (a, c, sigma) of step i (:437-451):
       alpha_next = abar[tau_{i+1}] (1 at the last step), sigma = kappa*sqrt((1-an)/(1-ab)*(1-ab/an)),
       a = sqrt(an/ab), c = sqrt(1-an-sigma^2) - sqrt(1-ab)*sqrt(an/ab);   x' = x*a + (c*eps + sigma*z)."""
    ab = sched["alpha_bars"][taus[i]]
    if i == len(taus) - 1:
        assert taus[i] == 0
        an, sigma = torch.tensor(1.0), torch.tensor(0.0)
    else:
        an = sched["alpha_bars"][taus[i + 1]]
        sigma = kappa * torch.sqrt((1 - an) / (1 - ab) * (1 - ab / an))
    a = torch.sqrt(an / ab)
    c = torch.sqrt(1 - an - sigma ** 2) - torch.sqrt(1 - ab) * torch.sqrt(an / ab)
    return a, c, sigma


def ddim_step(x, eps, noise, a, c, sigma):
    """x = x_noisy*a;  x += c*eps + sigma*z   (:450,:464-465)."""
    x = x * a
    x = x + (c * eps + sigma * noise)
    return x


def run_ddim(model_fn, sched, x_T, noises, S, skip_type="uniform", kappa=1.0):
    """noises[i] is the i-th draw (the reference draws torch.randn(size) on the CPU generator
    once per step, including the last one where sigma = 0, :464-465)."""
    T = sched["betas"].shape[0]
    taus = ddim_taus(T, S, skip_type)
    x = x_T
    traj = []
    for i, t in enumerate(taus):
        tt = torch.ones(x.shape[0]) * (t + 1)
        eps = model_fn(x, tt)
        a, c, sigma = ddim_coeffs(sched, taus, i, kappa)
        x = ddim_step(x, eps, noises[i], a, c, sigma)
        traj.append(x)
    return x, traj
