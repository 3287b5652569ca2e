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


def _conv_items_sched1(U, per_nt, grid, G):
    """Python model of the round-based work distribution of the convolution kernel (csrc/conv_tc.cu, `Items`, sched 1):
    CTA i owns T_i = floor(U (i+1) / grid) - floor(U i / grid) tiles in m = ceil(max T / G) items; round k is one contiguous
    stretch of the flat tile space holding every CTA's k-th item; an item is cut at n-tile boundaries."""
    q, n_p = U // grid, U - (U // grid) * grid
    n_q, m = grid - n_p, -(-(q + (1 if n_p else 0)) // G)
    per = [[] for _ in range(grid)]
    for i in range(grid):
        u_begin, u_endr = U * i // grid, U * (i + 1) // grid
        c_i, big = u_begin - q * i, (u_endr - u_begin) > q
        k, u, u_end = 0, 0, 0
        while True:
            while u >= u_end and k < m:
                a0, a1, b0, b1 = q * k // m, q * (k + 1) // m, (q + 1) * k // m, (q + 1) * (k + 1) // m
                u = n_q * a0 + n_p * b0 + (i - c_i) * (a1 - a0) + c_i * (b1 - b0)
                u_end = u + ((b1 - b0) if big else (a1 - a0))
                k += 1
            if u >= u_end:
                break
            nt = u // per_nt
            e = min(u_end, (nt + 1) * per_nt)
            per[i].append((nt, u - nt * per_nt, e - u))
            u = e
    return per


def test_conv_interleaved_schedule_covers_every_tile_once():
    # (tiles per shape, shapes, n-tiles, SMs, G): the step's shape classes at B = 32 plus ragged small cases
    for ntile, B, n_nt, sms, G in ((307, 32, 1, 148, 8), (46, 32, 1, 148, 4), (8, 32, 1, 148, 4), (46, 32, 2, 148, 4),
                                   (5, 3, 2, 148, 8), (1, 1, 1, 148, 8), (16, 32, 3, 148, 4), (307, 1, 1, 148, 8), (41, 32, 1, 148, 8), (9, 7, 1, 148, 2)):
        per_nt, U = B * ntile, B * ntile * n_nt
        per_cta = -(-U // sms)
        grid = -(-U // per_cta)                       # conv_tc_run: fewest CTAs that reach the minimal maximum
        per = _conv_items_sched1(U, per_nt, grid, G)
        seen = []
        for items in per:
            for nt, v0, n in items:
                assert 1 <= n <= G and v0 + n <= per_nt        # fits the TMEM accumulators, never crosses an n-tile
                seen += [nt * per_nt + v0 + k for k in range(n)]
        assert sorted(seen) == list(range(U))
        totals = [sum(n for _, _, n in items) for items in per]
        assert max(totals) == per_cta and max(totals) - min(totals) <= 1
