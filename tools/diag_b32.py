import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.synth import synth_state_dict
from tests.util import gen
from lion_b200.config import default_prior_cfg
from lion_b200.models.latent_points_ada_localprior import PVCNN2Prior
keys = json.load(open('tests/golden/keys.json'))
cfg = default_prior_cfg()
m = PVCNN2Prior(cfg.sde, 1, cfg); m.load_state_dict(synth_state_dict(keys['prior'], 11)); m = m.cuda().eval()
B = int(os.environ.get("B", "32"))
x, style = gen(31, B, 8192, 1, 1).cuda(), gen(32, B, 128, 1, 1).cuda()
t = torch.full((B,), 500.0).cuda()
try:
    for i in range(int(os.environ.get("N", "3"))):
        a = m(x=x, t=t, condition_input=style)
        torch.cuda.synchronize()
        print("forward", i, "ok", float(a.abs().mean()))
except Exception as e:
    print("FAILED:", str(e)[:300].replace("\n", " | "))
