"""the probability-flow ODE sampler (DiffusionVPSDE.sample_model_ode, host-side adaptive RK45 around
the network forward) against tests/golden/ode_sample.npz, produced by the reference's OWN utils/diffusion_continuous.py +
vendored torchdiffeq scipy wrapper on CPU (tests/golden/make_golden_ode.py).  Short integration spans: with random weights
the full span is a diverging ODE (12 000 evaluations), useless as a pin."""
import json
import os

import numpy as np
import pytest
import torch

from tests.synth import synth_state_dict
from tests.util import assert_close

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = json.load(open(os.path.join(G, "keys.json")))


def _setup():
    from lion_b200.config import default_prior_cfg
    from lion_b200.models.latent_points_ada_localprior import PVCNN2Prior
    from lion_b200.models.score_sde.resnet import PriorSEDrop
    from lion_b200.utils.diffusion_continuous import make_diffusion
    cfg = default_prior_cfg()
    gp = PriorSEDrop(cfg.sde, 128, cfg)
    gp.load_state_dict(synth_state_dict(KEYS["global"], 14))
    lp = PVCNN2Prior(cfg.sde, 1, cfg)
    lp.load_state_dict(synth_state_dict(KEYS["prior"], 11))
    return cfg, gp.cuda().eval(), lp.cuda().eval(), make_diffusion(cfg.sde)


def test_vpsde_closed_forms():
    _, _, _, d = _setup()
    t = torch.tensor([0.0, 0.3, 1.0], device="cuda")
    assert torch.allclose(d.var(t), 1.0 - torch.exp(-0.1 * t - 0.5 * 19.9 * t * t))
    assert torch.allclose(d.g2(t), 0.1 + 19.9 * t) and torch.allclose(d.f(t), -0.5 * d.g2(t))
    assert torch.allclose(d.inv_var(d.var(t[1:2])), t[1:2], atol=1e-4)      # (var(1) rounds to 1.0 in fp32: not invertible there)
    assert torch.allclose(d.e2int_f(t) ** 2, 1.0 - d.var(t), atol=1e-6)


def test_ode_sampler_global_prior_golden():
    z = np.load(os.path.join(G, "ode_sample.npz"))
    _, gp, _, d = _setup()
    out, nfe, _ = d.sample_model_ode(gp, 2, [128, 1, 1], 1e-5, float(z["g_tol"]), False, 1.0, noise=torch.from_numpy(z["g_noise"]).cuda(),
                                     init_t=float(z["g_init_t"]))
    assert abs(nfe - int(z["g_nfe"])) <= max(12, int(z["g_nfe"]) // 4), (nfe, int(z["g_nfe"]))
    assert_close(out, torch.from_numpy(z["g_out"]), 2e-2, "ODE sample of the global prior vs the reference's solver")


def test_ode_sampler_point_prior_golden():
    z = np.load(os.path.join(G, "ode_sample.npz"))
    _, _, lp, d = _setup()
    out, nfe, _ = d.sample_model_ode(lp, 1, [8192, 1, 1], 1e-5, float(z["l_tol"]), False, 1.0, noise=torch.from_numpy(z["l_noise"]).cuda(),
                                     condition_input=torch.from_numpy(z["l_style"]).cuda(), init_t=float(z["l_init_t"]))
    assert abs(nfe - int(z["l_nfe"])) <= max(12, int(z["l_nfe"]) // 2), (nfe, int(z["l_nfe"]))
    # adaptive steps at solver tolerance 1e-2 on a TF32 network against an fp32 CPU run: agreement to ~10x the solver tolerance
    assert_close(out, torch.from_numpy(z["l_out"]), 1e-1, "ODE sample of the latent-point prior vs the reference's solver")


def test_generate_samples_ode_route():
    """generate_samples_vada_2prior(ode_sample=1): global latent -> style -> point latent -> decoder (train_2prior.py:64-80)."""
    from lion_b200.models.vae_adain import Model
    from lion_b200.trainers.train_2prior import generate_samples_vada_2prior
    from lion_b200.utils.diffusion_continuous import DiffusionVPSDE
    cfg, gp, lp, _ = _setup()

    class ShortSpan(DiffusionVPSDE):              # full-span integration of random-weight priors diverges; keep the test short
        def sample_model_ode(self, *a, **k):
            k["init_t"] = 0.05
            return super().sample_model_ode(*a, **k)

    vae = Model(cfg)
    vae.decoder.load_state_dict(synth_state_dict(KEYS["decoder"], 13))
    vae = vae.cuda().eval()
    torch.manual_seed(9)
    img, nfe, t_ode, t_all, out = generate_samples_vada_2prior(vae.latent_shape(), torch.nn.ModuleList([gp, lp]), ShortSpan(cfg.sde), vae, 2,
                                                               False, ode_sample=1, ode_eps=1e-5, ode_solver_tol=1e-2)
    assert img.shape == (2, 2048, 3) and torch.isfinite(img).all() and float(nfe) > 0 and out["sampled_eps"].shape == (2, 8192, 1, 1)
