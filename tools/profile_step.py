"""One denoising step of both priors at the benchmarked batch, eager (no CUDA graph), bracketed by cudaProfilerStart/Stop
so that `ncu --profile-from-start off` sees exactly one step after a warm-up step:

    ncu --metrics <list> --clock-control none --profile-from-start off --csv --log-file gpurun_out/step.csv python tools/profile_step.py

Weights: key-seeded synthetic (tests/synth.py), x / style: seeded noise -- the same shapes bench.py times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lion_b200.config import default_prior_cfg
from lion_b200.models.latent_points_ada_localprior import PVCNN2Prior
from lion_b200.models.score_sde.resnet import PriorSEDrop
from tests.synth import synth_state_dict

B = int(os.environ.get("B", "32"))
cfg = default_prior_cfg()
shp = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
gp = PriorSEDrop(cfg.sde, 128, cfg)
gp.load_state_dict(synth_state_dict(shp(gp), 14))
lp = PVCNN2Prior(cfg.sde, 1, cfg)
lp.load_state_dict(synth_state_dict(shp(lp), 11))
gp, lp = gp.cuda().eval(), lp.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 8192, 1, 1, device="cuda", generator=g)
xg = torch.randn(B, 128, 1, 1, device="cuda", generator=g)
style = torch.randn(B, 128, 1, 1, device="cuda", generator=g)
t = torch.full((B,), 500.0, device="cuda")
for _ in range(2):                                  # warm-up: packs weights, sizes the arena, caches the style Linears
    gp(x=xg, t=t)
    lp(x=x, t=t, condition_input=style)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gp(x=xg, t=t)
lp(x=x, t=t, condition_input=style)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one global-prior + one PVCNN2Prior forward at B=%d" % B)
