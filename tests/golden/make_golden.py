"""Generate tests/golden/*.npz and *_keys.json from the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py

What is pinned (the reference has no golden vectors of its own):
  keys.json            state_dict key -> shape of PVCNN2Prior (+CLIP), LatentPointDecPVC,
                       PriorSEDrop, PriorSEClip            (checkpoint key contract, App. E)
  prior_fwd.npz        PVCNN2Prior.forward, B=2, two different t
  prior_clip_fwd.npz   same with clipforge.enable=1
  decoder_fwd.npz      LatentPointDecPVC.forward, B=1
  global_fwd.npz       PriorSEDrop / PriorSEClip forward
  ddpm10.npz           DiffusionDiscretized(num_steps=10).run_denoising_diffusion on both priors
                       with given_noise, then vae.sample -> BASELINE.json configs[0]
                       ("1 shape, 10 DDPM steps, 2048 pts, CPU PyTorch reference")
  schedule.npz         the fp32 schedule tables for num_steps=1000

Weights come from tests/synth.py (key-name seeded), loaded into the reference modules with
load_state_dict(strict=True).  The reference's CUDA point ops are replaced by
oracle/point_ops.py (they cannot run here); everything else is the reference's own code.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import ref_import as R  # noqa: E402

R.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.synth import synth_state_dict  # noqa: E402

torch.set_num_threads(8)


def shapes_of(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def load_synth(m, seed):
    sd = synth_state_dict(shapes_of(m), seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return sd


def gen(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def main():
    cfg = R.load_cfg()
    cfg_clip = R.load_cfg(overrides=["clipforge.enable", 1,
                                     "latent_pts.style_prior", "models.score_sde.resnet.PriorSEClip"])
    from models.latent_points_ada_localprior import PVCNN2Prior
    from models.score_sde.resnet import PriorSEDrop, PriorSEClip
    from models.latent_points_ada import LatentPointDecPVC
    from models.vae_adain import Model
    from utils.diffusion_pvd import DiffusionDiscretized

    keys = {}
    with torch.no_grad():
        # ---- local prior ------------------------------------------------------------
        prior = PVCNN2Prior(cfg.sde, 1, cfg)
        keys["prior"] = shapes_of(prior)
        load_synth(prior, 11)
        x = gen(101, 2, 8192, 1, 1)
        t = torch.tensor([981.0, 12.0])
        style = gen(102, 2, 128, 1, 1)
        eps = prior(x=x, t=t, condition_input=style, clip_feat=None)
        np.savez_compressed(os.path.join(HERE, "prior_fwd.npz"), x=x.numpy(), t=t.numpy(),
                            style=style.numpy(), eps=eps.numpy())
        print("prior_fwd", float(eps.abs().mean()), float(eps.std()))

        prior_c = PVCNN2Prior(cfg_clip.sde, 1, cfg_clip)
        keys["prior_clip"] = shapes_of(prior_c)
        load_synth(prior_c, 12)
        x = gen(111, 1, 8192, 1, 1)
        t = torch.tensor([500.0])
        style = gen(112, 1, 128, 1, 1)
        clip = gen(113, 1, 512)
        eps = prior_c(x=x, t=t, condition_input=style, clip_feat=clip)
        np.savez_compressed(os.path.join(HERE, "prior_clip_fwd.npz"), x=x.numpy(), t=t.numpy(),
                            style=style.numpy(), clip=clip.numpy(), eps=eps.numpy())
        print("prior_clip_fwd", float(eps.abs().mean()))

        # ---- decoder ----------------------------------------------------------------
        dec = LatentPointDecPVC(point_dim=3, context_dim=1, args=cfg)
        keys["decoder"] = shapes_of(dec)
        load_synth(dec, 13)
        ctx = gen(121, 1, 8192)
        style = gen(122, 1, 128)
        pts = dec(None, beta=None, context=ctx, style=style)
        np.savez_compressed(os.path.join(HERE, "decoder_fwd.npz"), context=ctx.numpy(), style=style.numpy(),
                            points=pts.numpy())
        print("decoder_fwd", float((pts - ctx.view(1, 2048, 4)[:, :, :3]).abs().mean()))

        # ---- global prior -----------------------------------------------------------
        gp = PriorSEDrop(cfg.sde, cfg.latent_pts.style_dim, cfg)
        keys["global"] = shapes_of(gp)
        load_synth(gp, 14)
        x = gen(131, 3, 128, 1, 1)
        t = torch.tensor([1000.0, 400.0, 1.0])
        out = gp(x=x, t=t, condition_input=None, clip_feat=None)
        gpc = PriorSEClip(cfg_clip.sde, cfg_clip.latent_pts.style_dim, cfg_clip)
        keys["global_clip"] = shapes_of(gpc)
        load_synth(gpc, 15)
        xc = gen(132, 2, 128, 1, 1)
        tc = torch.tensor([77.0, 640.0])
        clipc = gen(133, 2, 512)
        outc = gpc(x=xc, t=tc, condition_input=None, clip_feat=clipc)
        np.savez_compressed(os.path.join(HERE, "global_fwd.npz"), x=x.numpy(), t=t.numpy(), out=out.numpy(),
                            xc=xc.numpy(), tc=tc.numpy(), clipc=clipc.numpy(), outc=outc.numpy())
        print("global_fwd", float(out.abs().mean()), float(outc.abs().mean()))

        # ---- schedule tables (T=1000) -------------------------------------------------
        diff = DiffusionDiscretized(cfg.sde, None, cfg)
        np.savez_compressed(os.path.join(HERE, "schedule.npz"), betas=diff._betas_init.numpy(),
                            alphas=diff._alphas.numpy(), alpha_bars=diff._alpha_bars.numpy(),
                            betas_post=diff._betas_post_init.numpy())

        # ---- configs[0]: 1 shape, 10 DDPM steps -----------------------------------------
        cfg10 = R.load_cfg(overrides=["ddpm.num_steps", 10])
        diff10 = DiffusionDiscretized(cfg10.sde, None, cfg10)
        vae = Model(cfg10)
        keys["vae_decoder"] = {k: list(v.shape) for k, v in vae.state_dict().items() if k.startswith("decoder.")}
        vae.decoder.load_state_dict(dec.state_dict())
        vae.eval()
        T = 10
        noise_g = (gen(141, 1, 128, 1, 1), [gen(1410 + i, 1, 128, 1, 1) for i in range(T)])
        noise_l = (gen(142, 1, 8192, 1, 1), [gen(1420 + i, 1, 8192, 1, 1) for i in range(T)])
        z_g, lst_g = diff10.run_denoising_diffusion(gp, 1, [128, 1, 1], given_noise=noise_g)
        cond = vae.global2style(z_g)
        z_l, lst_l = diff10.run_denoising_diffusion(prior, 1, [8192, 1, 1], condition_input=cond,
                                                    given_noise=noise_l)
        eps_all = vae.compose_eps([z_g, z_l])
        img = vae.sample(num_samples=1, decomposed_eps=vae.decompose_eps(eps_all))
        np.savez_compressed(
            os.path.join(HERE, "ddpm10.npz"),
            xT_g=noise_g[0].numpy(), z_g=torch.stack(noise_g[1]).numpy(),
            xT_l=noise_l[0].numpy(), z_l=torch.stack(noise_l[1]).numpy(),
            out_g=z_g.numpy(), out_l=z_l.numpy(), image=img.numpy(),
            traj_l=torch.stack(lst_l["pred_x"]).numpy()[:, 0, :64, 0, 0],
            traj_full=torch.stack(lst_l["pred_x"]).numpy()[:, 0, :, 0, 0].astype(np.float32),
            betas=diff10._betas_init.numpy(), alpha_bars=diff10._alpha_bars.numpy())
        print("ddpm10", float(z_g.abs().mean()), float(z_l.abs().mean()), float(img.abs().mean()))

    with open(os.path.join(HERE, "keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
