"""Host-side logic that feeds the update kernels (no GPU needed): the per-step scalar tables of the
DDPM / DDIM / scheduler routes are built with the reference's fp32 expressions and must equal the
oracle's scalars bit for bit."""
import torch

from oracle import diffusion as OD
from oracle import scheduler as OS


def _diff(T):
    from lion_b200.config import default_prior_cfg
    from lion_b200.utils.diffusion_pvd import DiffusionDiscretized
    cfg = default_prior_cfg(num_steps=T)
    return DiffusionDiscretized(cfg.sde, None, cfg)


def test_ddpm_tables_match_oracle_scalars():
    d = _diff(1000)
    sched = OD.make_schedule(1000, 1e-4, 0.02)
    assert torch.equal(d._betas_init.cpu(), sched["betas"]) and torch.equal(d._alpha_bars.cpu(), sched["alpha_bars"])
    tab = d._step_tables(torch.device("cpu"))
    assert tab.shape == (1000, 4) and tab.dtype == torch.float32
    for t in (999, 500, 1):
        assert tab[t, 0] == 1.0 / torch.sqrt(sched["alphas"][t])
        assert tab[t, 1] == sched["betas"][t]
        assert tab[t, 2] == torch.sqrt(1.0 - sched["alpha_bars"][t])
        assert tab[t, 3] == torch.exp(0.5 * torch.log(sched["betas"][t]))
    assert tab[0, 0] == 1.0 / torch.sqrt(sched["alpha_bars"][0]) and tab[0, 1] == torch.sqrt(1.0 - sched["alpha_bars"][0])
    assert tab[0, 2] == 1.0 and tab[0, 3] == 0.0


def test_ddim_tables_match_oracle_scalars():
    d = _diff(1000)
    sched = OD.make_schedule(1000, 1e-4, 0.02)
    for skip, kappa, S in (("uniform", 1.0, 100), ("quad", 0.5, 25), ("uniform", 0.0, 7)):
        taus = OD.ddim_taus(1000, S, skip)
        assert taus[-1] == 0 and all(a >= b for a, b in zip(taus, taus[1:]))
        tab = d._ddim_tables(taus, kappa, torch.device("cpu"))
        assert tab.shape == (S, 4)
        for i in (0, 1, S // 2, S - 2, S - 1):
            a, c, sigma = OD.ddim_coeffs(sched, taus, i, kappa)
            assert tab[i, 0] == a and tab[i, 1] == c and tab[i, 2] == sigma and tab[i, 3] == taus[i] + 1


def test_scheduler_tables_match_oracle_scalars():
    from lion_b200.utils.ddpm_scheduler import DDPMScheduler
    sc = DDPMScheduler(clip_sample=False, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", num_train_timesteps=50,
                       variance_type="fixedlarge")
    s = OS.make_scheduler(50)
    tab = sc.step_tables(torch.device("cpu"))
    assert tab.shape == (50, 8)
    x, e, z = torch.randn(64), torch.randn(64), torch.randn(64)
    for t in (49, 20, 1, 0):
        r = tab[t]
        x0 = (x - r[0] * e) / r[1]
        prev = r[2] * x0 + r[3] * x
        out = prev + r[4] * z if t > 0 else prev
        assert torch.equal(out, OS.step(s, e, t, x, z if t > 0 else None))
    assert float(tab[0, 4]) == 0.0
    for bad in ({"clip_sample": True}, {"beta_schedule": "cosine", "clip_sample": False}):
        try:
            DDPMScheduler(**bad)
            assert False, "expected NotImplementedError"
        except NotImplementedError:
            pass


def test_vpsde_closed_forms_cpu():
    """utils/diffusion_continuous.py:571-621 restated: var / g2 / f / e2int_f / inv_var of the linear-beta VPSDE."""
    import torch
    from lion_b200.config import default_prior_cfg
    from lion_b200.utils.diffusion_continuous import make_diffusion
    d = make_diffusion(default_prior_cfg().sde)
    t = torch.tensor([0.0, 0.3, 0.7], dtype=torch.float64)
    assert torch.allclose(d.var(t), 1.0 - torch.exp(-0.1 * t - 0.5 * 19.9 * t * t))
    assert torch.allclose(d.g2(t), 0.1 + 19.9 * t) and torch.allclose(d.f(t), -0.5 * d.g2(t))
    assert torch.allclose(d.inv_var(d.var(t[1:])), t[1:], atol=1e-9)
    assert torch.allclose(d.e2int_f(t) ** 2, 1.0 - d.var(t), atol=1e-12)


def test_cuda_order_mean_emulation_is_a_mean():
    """oracle/point_ops.py::cuda_mean_lastdim reorders a float32 summation (pinned bit-for-bit against torch-CUDA on the
    GPU box); on the CPU it must at least be a correct mean for every block shape the configuration formula yields."""
    import numpy as np
    import torch
    from oracle import point_ops as P
    g = torch.Generator().manual_seed(3)
    for B, N in [(1, 2048), (32, 1024), (2, 64), (3, 700), (2, 33), (7, 128), (5, 4096), (1, 5)]:
        x = torch.randn(B, 3, N, generator=g)
        m = P.cuda_mean_lastdim(x.numpy())
        ref = x.double().mean(2).numpy()
        assert np.abs(m - ref).max() < 5e-7, (B, N)
    nc, vox = P.voxel_coords_cuda_order(torch.randn(2, 3, 256, generator=g).numpy(), 8)
    assert vox.dtype == torch.int32 and int(vox.min()) >= 0 and int(vox.max()) <= 7
