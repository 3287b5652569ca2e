"""the VAE encoder path on the CUDA engine -- PointNetPlusEncoder (non-Ada PVCNN blocks, plain
GroupNorm) and PointTransPVC (the Ada U-Net with embed_dim = 0, 3-channel input, 8 outputs per point) -- against
tests/golden/encoder_fwd.npz, which was produced by the reference's OWN modules (tests/golden/make_golden_encoder.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import encoder as OE
from tests.synth import synth_state_dict
from tests.util import assert_close, gen

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = json.load(open(os.path.join(G, "keys_encoder.json")))
TOL = 5e-3      # 3x3x3 / 1x1 convolutions in TF32, as the reference's cuDNN path runs them


def _vae():
    from lion_b200.config import default_prior_cfg
    from lion_b200.models.vae_adain import Model
    vae = Model(default_prior_cfg())
    vae.style_encoder.load_state_dict(synth_state_dict(KEYS["style_encoder"], 21), strict=True)
    vae.encoder.load_state_dict(synth_state_dict(KEYS["point_encoder"], 22), strict=True)
    shp = {k: list(v.shape) for k, v in vae.decoder.state_dict().items()}
    vae.decoder.load_state_dict(synth_state_dict(shp, 13), strict=True)
    return vae.cuda().eval()


def test_style_encoder_golden():
    z = np.load(os.path.join(G, "encoder_fwd.npz"))
    out = _vae().style_encoder(torch.from_numpy(z["x"]).cuda())
    assert_close(out["mu_1d"], torch.from_numpy(z["s_mu"]), TOL, "PointNetPlusEncoder mu vs reference golden")
    assert_close(out["sigma_1d"], torch.from_numpy(z["s_sigma"]), TOL, "PointNetPlusEncoder log sigma vs reference golden")


def test_point_encoder_golden():
    z = np.load(os.path.join(G, "encoder_fwd.npz"))
    x = torch.from_numpy(z["x"])
    out = _vae().encoder([x.cuda(), torch.from_numpy(z["style"]).cuda()])
    # compare the network's contribution: mu = skip_weight * net + x, so subtract the xyz pass-through
    B = x.shape[0]
    passthrough = torch.cat([x, torch.zeros(B, x.shape[1], 1)], dim=2).reshape(B, -1)
    assert_close(out["mu_1d"].cpu() - passthrough, torch.from_numpy(z["e_mu"]) - passthrough, TOL, "PointTransPVC mu (network part)")
    assert_close(out["sigma_1d"], torch.from_numpy(z["e_sigma"]), TOL, "PointTransPVC log sigma vs reference golden")


def test_encoders_batch_vs_oracle_and_recont_shapes():
    """B = 3 against the CPU oracle (itself pinned to the golden), then encode / recont wiring: shapes, the log-sigma
    offset, determinism of the deterministic parts, and the decoder consuming the sampled latents."""
    vae = _vae()
    x = gen(401, 3, 2048, 3) * 0.5
    sd_s = synth_state_dict(KEYS["style_encoder"], 21)
    out = vae.style_encoder(x.cuda())
    with torch.no_grad():
        mu_o, sig_o = OE.style_encoder_forward(sd_s, x)
    assert_close(out["mu_1d"], mu_o, TOL, "style encoder mu, B=3")
    assert_close(out["sigma_1d"], sig_o, TOL, "style encoder log sigma, B=3")
    torch.manual_seed(3)
    all_eps, all_log_q, latent_list = vae.encode(x.cuda())
    assert all_eps.shape == (3, 128 + 8192) and len(latent_list) == 2
    zg, mu_g, ls_g = latent_list[0]
    assert torch.equal(mu_g, out["mu_1d"]), "style encoder is not bit-reproducible"
    d = vae.encode_local(x.cuda(), zg)
    assert torch.equal(d.mu, latent_list[1][1]) and torch.equal(d.log_sigma, latent_list[1][2])
    rec = vae.recont(x.cuda())
    assert rec["x_0_pred"].shape == (3, 2048, 3)          # (values: random-weight log-sigmas can overflow exp(); the decoder itself is golden-checked)
    assert torch.isfinite(latent_list[1][1]).all() and torch.isfinite(latent_list[0][1]).all()
    assert rec["all_eps"][0].shape == (3, 128, 1, 1) and rec["all_eps"][1].shape == (3, 8192, 1, 1)
    assert rec["vis/latent_pts"].shape == (3, 2048, 3)
