"""Generate tests/golden/encoder_fwd.npz + keys_encoder.json from the UNMODIFIED reference (container only):

    python tests/golden/make_golden_encoder.py

vae_adain.Model's `style_encoder` (PointNetPlusEncoder on the non-Ada pvcnn2.py blocks) and `encoder`
(PointTransPVC) on one 2048-point cloud, synthetic key-seeded weights (tests/synth.py).  SURVEY.md 8f rank 3.
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

from tests.golden.make_golden import load_synth, shapes_of, gen  # noqa: E402

torch.set_num_threads(8)


def main():
    cfg = R.load_cfg()
    from models.vae_adain import Model
    with torch.no_grad():
        vae = Model(cfg)
        vae.eval()
        keys = {"style_encoder": shapes_of(vae.style_encoder), "point_encoder": shapes_of(vae.encoder)}
        load_synth(vae.style_encoder, 21)
        load_synth(vae.encoder, 22)
        x = gen(301, 1, 2048, 3) * 0.5
        zs = vae.style_encoder(x)
        style = gen(302, 1, 128) * 0.5
        ze = vae.encoder([x, style])
        np.savez_compressed(os.path.join(HERE, "encoder_fwd.npz"), x=x.numpy(), style=style.numpy(),
                            s_mu=zs["mu_1d"].numpy(), s_sigma=zs["sigma_1d"].numpy(),
                            e_mu=ze["mu_1d"].numpy(), e_sigma=ze["sigma_1d"].numpy(),
                            log_sigma_offset=np.float32(cfg.shapelatent.log_sigma_offset),
                            skip_weight=np.float32(cfg.latent_pts.skip_weight),
                            pts_sigma_offset=np.float32(cfg.latent_pts.pts_sigma_offset))
        print("style", float(zs["mu_1d"].abs().mean()), "enc", float(ze["mu_1d"].abs().mean()), float(ze["sigma_1d"].abs().mean()))
    with open(os.path.join(HERE, "keys_encoder.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
