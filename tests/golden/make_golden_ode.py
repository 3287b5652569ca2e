"""Generate tests/golden/ode_sample.npz from the UNMODIFIED reference (container only):

    python tests/golden/make_golden_ode.py

`DiffusionVPSDE.sample_model_ode` (utils/diffusion_continuous.py:178-249, scipy RK45 through the reference's vendored
torchdiffeq wrapper) on CPU with key-seeded synthetic weights:
  * the global prior (PriorSEDrop), 2 samples, span t = 0.15 -> 1e-5, tolerance 1e-3 (with random weights the full span
    t = 1 -> 0 is a diverging ODE: 12 000 evaluations and |z| ~ 1e7, useless as a pin);
  * the latent-point prior (PVCNN2Prior), 1 sample, span t = 0.12 -> 1e-5, tolerance 1e-2 (a CPU forward takes ~1 s).
SURVEY.md 8f rank 4."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import ref_import as R  # noqa: E402

R.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.golden.make_golden import load_synth, gen  # noqa: E402

torch.set_num_threads(8)


def main():
    cfg = R.load_cfg()
    from models.latent_points_ada_localprior import PVCNN2Prior
    from models.score_sde.resnet import PriorSEDrop
    from utils.diffusion_continuous import make_diffusion
    diff = make_diffusion(cfg.sde)
    out = {}
    with torch.no_grad():
        gp = PriorSEDrop(cfg.sde, cfg.latent_pts.style_dim, cfg)
        load_synth(gp, 14)
        noise = gen(501, 2, 128, 1, 1)
        z, nfe, _ = diff.sample_model_ode(gp, 2, [128, 1, 1], 1e-5, 1e-3, False, 1.0, noise=noise, init_t=0.15)
        out.update(g_noise=noise.numpy(), g_out=z.numpy(), g_nfe=np.int32(nfe), g_tol=np.float32(1e-3), g_init_t=np.float32(0.15))
        print("global: nfe", nfe, float(z.abs().mean()))
        lp = PVCNN2Prior(cfg.sde, 1, cfg)
        load_synth(lp, 11)
        noise_l = gen(502, 1, 8192, 1, 1)
        style = gen(503, 1, 128, 1, 1)
        zl, nfe_l, _ = diff.sample_model_ode(lp, 1, [8192, 1, 1], 1e-5, 1e-2, False, 1.0, noise=noise_l, condition_input=style,
                                             init_t=0.12)
        out.update(l_noise=noise_l.numpy(), l_style=style.numpy(), l_out=zl.numpy(), l_nfe=np.int32(nfe_l), l_tol=np.float32(1e-2),
                   l_init_t=np.float32(0.12))
        print("local: nfe", nfe_l, float(zl.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "ode_sample.npz"), **out)


if __name__ == "__main__":
    main()
