"""Generate tests/golden/ddim5.npz from the UNMODIFIED reference (container only):

    python tests/golden/make_golden_ddim.py

DiffusionDiscretized(num_steps=10).run_ddim(ddim_step=5) on the global prior (uniform skip,
kappa=1) and on the local prior (uniform skip kappa=1, and quad skip kappa=0.5), same synthetic
weights as make_golden.py.  The reference draws x_T on 'cuda' (patched to CPU by ref_import) and
the per-step noise with torch.randn(size) on the CPU generator; torch.manual_seed before each run
fixes both, and the draws are replayed here to store them next to the outputs.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import ref_import as R  # noqa: E402

R.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.golden.make_golden import load_synth  # noqa: E402

torch.set_num_threads(8)


def replay(seed, size, S):
    torch.manual_seed(seed)
    x_T = torch.randn(size=size)
    z = [torch.randn(size) for _ in range(S)]
    return x_T, torch.stack(z)


def main():
    cfg10 = R.load_cfg(overrides=["ddpm.num_steps", 10])
    from models.latent_points_ada_localprior import PVCNN2Prior
    from models.score_sde.resnet import PriorSEDrop
    from utils.diffusion_pvd import DiffusionDiscretized
    S = 5
    out = {}
    with torch.no_grad():
        diff = DiffusionDiscretized(cfg10.sde, None, cfg10)
        gp = PriorSEDrop(cfg10.sde, cfg10.latent_pts.style_dim, cfg10)
        load_synth(gp, 14)
        prior = PVCNN2Prior(cfg10.sde, 1, cfg10)
        load_synth(prior, 11)

        torch.manual_seed(201)
        z_g, lst_g = diff.run_ddim(gp, 2, [128, 1, 1], ddim_step=S, skip_type='uniform', kappa=1.0)
        xT, z = replay(201, [2, 128, 1, 1], S)
        out.update(g_xT=xT.numpy(), g_z=z.numpy(), g_out=z_g.numpy(), g_traj=torch.stack(lst_g).numpy())

        cond = z_g[:1].clone()
        torch.manual_seed(202)
        z_l, lst_l = diff.run_ddim(prior, 1, [8192, 1, 1], condition_input=cond, ddim_step=S, skip_type='uniform', kappa=1.0)
        xT, z = replay(202, [1, 8192, 1, 1], S)
        out.update(l_cond=cond.numpy(), l_xT=xT.numpy(), l_z=z.numpy(), l_out=z_l.numpy(),
                   l_traj=torch.stack(lst_l).numpy()[:, 0, :, 0, 0])

        torch.manual_seed(203)
        z_q, lst_q = diff.run_ddim(gp, 2, [128, 1, 1], ddim_step=S, skip_type='quad', kappa=0.5)
        xT, z = replay(203, [2, 128, 1, 1], S)
        out.update(q_xT=xT.numpy(), q_z=z.numpy(), q_out=z_q.numpy(), q_traj=torch.stack(lst_q).numpy())
        out.update(alpha_bars=diff._alpha_bars.numpy())
    np.savez_compressed(os.path.join(HERE, "ddim5.npz"), **out)
    print("ddim5", float(z_g.abs().mean()), float(z_l.abs().mean()), float(z_q.abs().mean()))


if __name__ == "__main__":
    main()
