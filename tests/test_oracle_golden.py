"""The CPU oracle (oracle/) against the golden vectors made from the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; the CUDA path is then compared
with the oracle (and with the same goldens) in the -m gpu tests."""
import json
import os

import numpy as np
import torch

from oracle import diffusion as OD
from oracle import net as ON
from tests.synth import synth_state_dict

G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = json.load(open(os.path.join(G, "keys.json")))


def _close(a, b, rtol, name):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= rtol * ref, "%s: max-abs err %.3e vs scale %.3e" % (name, err, ref)


def test_schedule_tables():
    z = np.load(os.path.join(G, "schedule.npz"))
    s = OD.make_schedule(1000, 1e-4, 0.02)
    assert np.array_equal(s["betas"].numpy(), z["betas"])
    assert np.array_equal(s["alphas"].numpy(), z["alphas"])
    assert np.array_equal(s["alpha_bars"].numpy(), z["alpha_bars"])


def test_prior_forward_matches_reference():
    z = np.load(os.path.join(G, "prior_fwd.npz"))
    sd = synth_state_dict(KEYS["prior"], 11)
    with torch.no_grad():
        eps = ON.prior_forward(sd, ON.prior_spec(), torch.from_numpy(z["x"]), torch.from_numpy(z["t"]),
                               torch.from_numpy(z["style"]))
    _close(eps, z["eps"], 2e-5, "prior eps")


def test_prior_clip_forward_matches_reference():
    z = np.load(os.path.join(G, "prior_clip_fwd.npz"))
    sd = synth_state_dict(KEYS["prior_clip"], 12)
    with torch.no_grad():
        eps = ON.prior_forward(sd, ON.prior_spec(clip=True), torch.from_numpy(z["x"]), torch.from_numpy(z["t"]),
                               torch.from_numpy(z["style"]), clip_feat=torch.from_numpy(z["clip"]))
    _close(eps, z["eps"], 2e-5, "prior clip eps")


def test_decoder_forward_matches_reference():
    z = np.load(os.path.join(G, "decoder_fwd.npz"))
    sd = synth_state_dict(KEYS["decoder"], 13)
    with torch.no_grad():
        pts = ON.decoder_forward(sd, ON.decoder_spec(), torch.from_numpy(z["context"]), torch.from_numpy(z["style"]))
    _close(pts, z["points"], 1e-5, "decoder points")


def test_global_prior_matches_reference():
    z = np.load(os.path.join(G, "global_fwd.npz"))
    with torch.no_grad():
        out = ON.global_prior_forward(synth_state_dict(KEYS["global"], 14), torch.from_numpy(z["x"]),
                                      torch.from_numpy(z["t"]))
        outc = ON.global_prior_forward(synth_state_dict(KEYS["global_clip"], 15), torch.from_numpy(z["xc"]),
                                       torch.from_numpy(z["tc"]), clip_feat=torch.from_numpy(z["clipc"]))
    _close(out, z["out"], 2e-5, "global")
    _close(outc, z["outc"], 2e-5, "global clip")


def test_ddpm10_config0_matches_reference():
    """BASELINE.json configs[0]: airplane prior, 1 shape, 10 DDPM steps, 2048 pts, CPU."""
    z = np.load(os.path.join(G, "ddpm10.npz"))
    sched = OD.make_schedule(10, 1e-4, 0.02)
    assert np.array_equal(sched["betas"].numpy(), z["betas"])
    sd_g = synth_state_dict(KEYS["global"], 14)
    sd_l = synth_state_dict(KEYS["prior"], 11)
    sd_d = synth_state_dict(KEYS["decoder"], 13)
    spec, dspec = ON.prior_spec(), ON.decoder_spec()
    with torch.no_grad():
        pts, z_g, z_l = OD.sample_2prior(
            lambda x, t: ON.global_prior_forward(sd_g, x, t),
            lambda x, t, style: ON.prior_forward(sd_l, spec, x, t, style),
            lambda ctx, style: ON.decoder_forward(sd_d, dspec, ctx, style),
            sched,
            (torch.from_numpy(z["xT_g"]), list(torch.from_numpy(z["z_g"]))),
            (torch.from_numpy(z["xT_l"]), list(torch.from_numpy(z["z_l"]))))
    _close(z_g, z["out_g"], 5e-5, "global latent")
    _close(z_l, z["out_l"], 2e-4, "local latent")
    _close(pts, z["image"], 2e-4, "decoded points")


def test_ddim5_matches_reference():
    """run_ddim (5 of 10 steps; uniform kappa=1 on both priors, quad kappa=0.5 on the
    global prior; noise = the reference's CPU-generator draws, stored in the golden)."""
    z = np.load(os.path.join(G, "ddim5.npz"))
    sched = OD.make_schedule(10, 1e-4, 0.02)
    assert np.array_equal(sched["alpha_bars"].numpy(), z["alpha_bars"])
    assert OD.ddim_taus(10, 5, "uniform") == [9, 6, 4, 2, 0]
    assert OD.ddim_taus(10, 5, "quad") == [8, 4, 2, 0, 0]
    sd_g = synth_state_dict(KEYS["global"], 14)
    sd_l = synth_state_dict(KEYS["prior"], 11)
    spec = ON.prior_spec()
    with torch.no_grad():
        g_out, g_traj = OD.run_ddim(lambda x, t: ON.global_prior_forward(sd_g, x, t), sched,
                                    torch.from_numpy(z["g_xT"]), list(torch.from_numpy(z["g_z"])), 5)
        q_out, _ = OD.run_ddim(lambda x, t: ON.global_prior_forward(sd_g, x, t), sched,
                               torch.from_numpy(z["q_xT"]), list(torch.from_numpy(z["q_z"])), 5, "quad", 0.5)
        style = torch.from_numpy(z["l_cond"]).reshape(1, -1)
        l_out, l_traj = OD.run_ddim(lambda x, t: ON.prior_forward(sd_l, spec, x, t, style), sched,
                                    torch.from_numpy(z["l_xT"]), list(torch.from_numpy(z["l_z"])), 5)
    _close(g_out, z["g_out"], 5e-5, "ddim global")
    _close(torch.stack(g_traj), z["g_traj"], 5e-5, "ddim global trajectory")
    _close(q_out, z["q_out"], 5e-5, "ddim global quad")
    _close(l_out, z["l_out"], 2e-4, "ddim local")
    _close(torch.stack(l_traj)[:, 0, :, 0, 0], z["l_traj"], 2e-4, "ddim local trajectory")
