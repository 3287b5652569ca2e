import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.synth import synth_state_dict
from tests.util import gen
from lion_b200.config import default_prior_cfg
from lion_b200.models.pvcnn2_ada import PVConv
cfg = default_prior_cfg()
B = int(os.environ.get("B", "32"))
cin, cout, r, N = 64, 64, 32, 2048
pv = PVConv(cin, cout, 3, r, with_se=True, cfg=cfg)
pv.load_state_dict(synth_state_dict({k: list(v.shape) for k, v in pv.state_dict().items()}, 23)); pv = pv.cuda().eval()
f, c, s = gen(4, B, cin, N).cuda(), gen(5, B, 3, N, scale=0.4).cuda(), gen(6, B, 128).cuda()
try:
    for i in range(10):
        o = pv((f, c, None, s))[0]
        torch.cuda.synchronize()
    print("OK", float(o.abs().mean()))
except Exception as e:
    print("FAILED", str(e)[:120].replace("\n", " "))
