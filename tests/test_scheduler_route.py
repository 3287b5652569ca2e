"""a21: the `LION.sample` route (models/lion.py) -- diffusers-style DDPM scheduler.  The scheduler is an
un-vendored third-party dependency (no reference-held vector: "parity unpinned" by the letter, see
oracle/scheduler.py).  The restatement is checked (i) against the known-answer values of the dependency's own
test-suite, (ii) for consistency with the in-tree DiffusionDiscretized posterior, which IS pinned by reference
goldens; the CUDA path is checked against the restatement."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import diffusion as OD
from oracle import net as ON
from oracle import scheduler as OS
from tests.synth import synth_state_dict
from tests.util import assert_close, rms_err, gen

G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = json.load(open(os.path.join(G, "keys.json")))


def test_scheduler_restatement_agrees_with_pinned_posterior():
    """For t > 0 the scheduler's mean c0*x0 + c1*x is algebraically the posterior mean that
    DiffusionDiscretized.get_q_posterior_mean evaluates (pinned by ddpm10.npz); the variance LION's
    'fixedlarge' string selects is the posterior variance (1-abar_{t-1})/(1-abar_t)*beta_t."""
    s = OS.make_scheduler(1000)
    sched = OD.make_schedule(1000)
    assert torch.allclose(s["alphas_cumprod"], sched["alpha_bars"], rtol=2e-5, atol=0)
    x, e = gen(71, 2, 512), gen(72, 2, 512)
    zero = torch.zeros_like(x)
    for t in (999, 640, 17, 1):
        mean_sched = OS.step(s, e, t, x, zero)
        mean_pvd = OD.ddpm_step(sched, x, e, t, zero)
        assert_close(mean_sched, mean_pvd, 2e-4, "posterior mean t=%d" % t)   # fp32 cancellation in 1-abar at small t
        ab, abp = sched["alpha_bars"][t].double(), sched["alpha_bars"][t - 1].double()
        v = (1 - abp) / (1 - ab) * sched["betas"][t].double()
        assert abs(float(OS.variance(s, t, "fixedlarge")) / float(v) - 1) < 1e-3    # fp32 1-abar at small t
        assert float(OS.variance(s, t, "fixed_large")) == float(s["betas"][t])
    # t = 0: x0 itself, no noise
    x0 = OS.step(s, e, 0, x, None)
    assert_close(x0, OD.ddpm_step(sched, x, e, 0, None), 5e-4, "t=0")   # fp32 (1 - 0.9999) in the scheduler vs float64 tables


def test_scheduler_restatement_reproduces_diffusers_known_answers():
    """The known-answer tests of the dependency's own test-suite for DDPMScheduler (diffusers 0.11 line,
    DDPMSchedulerTest.test_variance and .test_full_loop_no_noise; constants quoted from the published test file from
    memory, see oracle/scheduler.py) evaluated on the restatement, at the tolerances those tests state."""
    s = OS.make_scheduler(1000, 1e-4, 0.02)
    for t, want in ((0, 0.0), (487, 0.00979), (999, 0.02)):
        assert abs(float(OS.variance(s, t, "fixed_small")) - want) < 1e-5
    # dummy_sample_deter / dummy_model of the scheduler test base class: batch 4, 3 x 8 x 8
    n = 4 * 3 * 8 * 8
    x = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2)
    g = torch.Generator().manual_seed(0)
    for t in reversed(range(1000)):
        eps = x * t / (t + 1)
        z = torch.randn(eps.shape, generator=g, dtype=eps.dtype) if t > 0 else None
        x = OS.step(s, eps, t, x, z, variance_type="fixed_small", clip_sample=True)
    assert abs(float(x.abs().sum()) - 258.9070) < 1e-2
    assert abs(float(x.abs().mean()) - 0.3374) < 1e-3


def test_product_scheduler_tables_reproduce_diffusers_known_answers():
    """The same known answers through the PRODUCT's host side: `DDPMScheduler._get_variance` and the [T][8] table rows that
    feed `lion_scheduler_step`, evaluated with the kernel's own expression (csrc/ddpm.cu: k_sched_step) on the CPU; the
    clamp of x0 is the `clip_sample=True` of the diffusers test configuration, which LION itself switches off."""
    from lion_b200.utils.ddpm_scheduler import DDPMScheduler
    sc = DDPMScheduler(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                       variance_type="fixed_small", clip_sample=False)
    for t, want in ((0, 0.0), (487, 0.00979), (999, 0.02)):
        assert abs(float(sc._get_variance(t)) - want) < 1e-5
    tab = sc.step_tables(torch.device("cpu"))
    n = 4 * 3 * 8 * 8
    x = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2)
    g = torch.Generator().manual_seed(0)
    for t in reversed(range(1000)):
        eps = x * t / (t + 1)
        r = tab[t]
        x0 = torch.clamp((x - r[0] * eps) / r[1], -1, 1)
        prev = r[2] * x0 + r[3] * x
        x = prev + r[4] * torch.randn(eps.shape, generator=g, dtype=eps.dtype) if t > 0 else prev
    assert abs(float(x.abs().sum()) - 258.9070) < 1e-2
    assert abs(float(x.abs().mean()) - 0.3374) < 1e-3


@pytest.mark.gpu
def test_scheduler_step_kernel_matches_restatement():
    from lion_b200.utils.ddpm_scheduler import DDPMScheduler
    from lion_b200 import _lib as L
    sc = DDPMScheduler(clip_sample=False, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", num_train_timesteps=1000,
                       variance_type="fixedlarge")
    s = OS.make_scheduler(1000)
    assert torch.equal(sc.alphas_cumprod, s["alphas_cumprod"])
    assert sc.timesteps[0] == 999 and sc.timesteps[-1] == 0
    x, e, z = gen(73, 3, 8192), gen(74, 3, 8192), gen(75, 3, 8192)
    tab = sc.step_tables(torch.device("cuda"))
    for t in (999, 500, 1, 0):
        step = torch.tensor([t], dtype=torch.int32, device="cuda")
        out = torch.empty(3, 8192, device="cuda")
        xc, ec, zc = x.cuda(), e.cuda(), z.cuda()
        L.check(L.lib().lion_scheduler_step(L.ptr(xc), L.ptr(ec), L.ptr(zc), L.ptr(out), L.ptr(tab), L.ptr(step), x.numel(), L.stream()))
        ref = OS.step(s, e, t, x, z)
        assert torch.equal(out.cpu(), ref), "scheduler step differs from the restated arithmetic at t=%d" % t
    # the public step(): draws its own noise for t > 0, none at t = 0
    torch.manual_seed(5)
    o = sc.step(ec, 7, xc).prev_sample
    torch.manual_seed(5)
    zz = torch.randn(3, 8192, device="cuda")
    assert torch.equal(o.cpu(), OS.step(s, e, 7, x, zz.cpu()))
    st = torch.cuda.get_rng_state()
    o0 = sc.step(ec, 0, xc).prev_sample
    assert torch.equal(torch.cuda.get_rng_state(), st), "t = 0 must not consume the generator"
    assert torch.equal(o0.cpu(), OS.step(s, e, 0, x, None))


@pytest.mark.gpu
def test_lion_sample_route():
    from lion_b200.config import default_prior_cfg
    from lion_b200.models.lion import LION
    T = 6
    cfg = default_prior_cfg(num_steps=T)
    m = LION(cfg)
    sd_g, sd_l, sd_d = (synth_state_dict(KEYS[k], sd) for k, sd in (("global", 14), ("prior", 11), ("decoder", 13)))
    m.priors[0].load_state_dict(sd_g)
    m.priors[1].load_state_dict(sd_l)
    m.vae.decoder.load_state_dict(sd_d)
    outs = []
    for use_graph in (True, False):
        m.use_cuda_graph = use_graph
        torch.manual_seed(31)
        outs.append(m.sample(num_samples=2))
    assert outs[0]["points"].shape == (2, 2048, 3) and torch.isfinite(outs[0]["points"]).all()
    assert_close(outs[0]["z_global"], outs[1]["z_global"], 1e-6, "graph vs eager (global)")
    assert_close(outs[0]["points"], outs[1]["points"], 1e-3, "graph vs eager (points)")
    # replay the generator: x_T, T-1 step draws (none at t = 0), for the global then the local prior
    torch.manual_seed(31)
    xg = torch.randn(2, 128, 1, 1, device="cuda")
    zg = [torch.randn(2, 128, 1, 1, device="cuda") for _ in range(T - 1)]
    xl = torch.randn(2, 8192, 1, 1, device="cuda")
    zl = [torch.randn(2, 8192, 1, 1, device="cuda") for _ in range(T - 1)]
    s = OS.make_scheduler(T)
    spec = ON.prior_spec()
    with torch.no_grad():
        z_g = OS.run(lambda x, t: ON.global_prior_forward(sd_g, x, t), s, xg.cpu(), [z.cpu() for z in zg])
        assert_close(outs[0]["z_global"], z_g, 5e-3, "LION.sample global latent vs restatement")
        style = outs[0]["z_global"].cpu().reshape(2, -1)
        z_l = OS.run(lambda x, t: ON.prior_forward(sd_l, spec, x, t, style), s, xl.cpu(), [z.cpu() for z in zl])
    assert rms_err(outs[0]["z_local"], z_l) < 5e-2
