"""SURVEY 8f rank 3 groundwork: the encoder-path oracle against the reference's own modules
(tests/golden/encoder_fwd.npz, made by tests/golden/make_golden_encoder.py)."""
import json
import os

import numpy as np
import torch

from oracle import encoder as OE
from tests.synth import synth_state_dict

G = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, tol, name):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    e = ((a - b).abs().max() / b.abs().max()).item()
    assert e <= tol, "%s: %.3e > %.1e" % (name, e, tol)


def test_encoder_oracle_matches_reference():
    z = np.load(os.path.join(G, "encoder_fwd.npz"))
    keys = json.load(open(os.path.join(G, "keys_encoder.json")))
    sd_s = synth_state_dict(keys["style_encoder"], 21)
    sd_e = synth_state_dict(keys["point_encoder"], 22)
    x = torch.from_numpy(z["x"])
    with torch.no_grad():
        mu, sig = OE.style_encoder_forward(sd_s, x)
        _close(mu, z["s_mu"], 5e-5, "style encoder mu")
        _close(sig, z["s_sigma"], 5e-5, "style encoder log sigma")
        emu, esig = OE.point_encoder_forward(sd_e, x, torch.from_numpy(z["style"]), skip_weight=float(z["skip_weight"]),
                                             pts_sigma_offset=float(z["pts_sigma_offset"]))
        _close(emu, z["e_mu"], 5e-5, "point encoder mu")
        _close(esig, z["e_sigma"], 5e-5, "point encoder log sigma")
        # composition (recont's encoder half) with fixed draws: shapes and finiteness
        zg, zl = OE.encode(sd_s, sd_e, x, torch.zeros(1, 128), torch.zeros(1, 8192), float(z["log_sigma_offset"]))
        assert zg.shape == (1, 128) and zl.shape == (1, 8192) and torch.isfinite(zl).all()
        _close(zg, z["s_mu"], 5e-5, "z_global at eps = 0")
