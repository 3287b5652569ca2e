"""Where does one local-prior denoising step spend its time, per block and per stream?  The image has no nsys, so the
library can drop %globaltimer stamps into its two streams at block boundaries (LION_TIMELINE=1 at context creation;
include/lion_b200.h: lion_ctx_timeline).  This tool captures one PVCNN2Prior forward at B=32 in a CUDA graph -- as the
sampling loop does --, replays it, and prints the stamps of the last replay relative to the step's start.

    LION_TIMELINE=1 python tools/timeline_step.py > gpurun_out/timeline.txt

The ~100 one-thread stamp kernels add a few microseconds each: read the table for where the time goes, not for totals."""
import ctypes as C
import os
import sys

os.environ["LION_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lion_b200 import _lib as L
from lion_b200.config import default_prior_cfg
from lion_b200.models.latent_points_ada_localprior import PVCNN2Prior
from tests.synth import synth_state_dict

B = int(os.environ.get("B", "32"))
cfg = default_prior_cfg()
lp = PVCNN2Prior(cfg.sde, 1, cfg)
lp.load_state_dict(synth_state_dict({k: list(v.shape) for k, v in lp.state_dict().items()}, 11))
lp = lp.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 8192, 1, 1, device="cuda", generator=g) * 0.5
style = torch.randn(B, 128, 1, 1, device="cuda", generator=g)
t = torch.full((B,), 500.0, device="cuda")
for _ in range(2):
    lp(x=x, t=t, condition_input=style)
with L.capture_graph() as gr:
    out = lp(x=x, t=t, condition_input=style)
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gr.replay()
e1.record()
torch.cuda.synchronize()
N = 256
ts = (C.c_ulonglong * N)()
names = C.create_string_buffer(N * 24)
n = L.lib().lion_ctx_timeline(L.ctx(), C.cast(ts, C.c_void_p), C.cast(names, C.c_void_p), N)
if n < 0:
    raise SystemExit("lion_ctx_timeline failed: %s" % L.lib().lion_last_error().decode())
rows = [(ts[i], names.raw[i * 24:(i + 1) * 24].split(b"\0")[0].decode()) for i in range(n)]
t0 = min(r[0] for r in rows if r[1] == "start")
print("# one PVCNN2Prior forward, B=%d, graph replay; %d stamps; replay period %.3f ms (includes the stamp kernels)" % (
    B, n, e0.elapsed_time(e1) / 20))
print("# %-22s %10s %10s" % ("block", "t [us]", "d [us]"))
main = [r for r in rows if not r[1].startswith("aux:")]
aux = [r for r in rows if r[1].startswith("aux:")]
prev = t0
for tt, nm in main:
    print("  %-22s %10.1f %10.1f" % (nm, (tt - t0) / 1e3, (tt - prev) / 1e3))
    prev = tt
print("# side stream")
for tt, nm in aux:
    print("  %-22s %10.1f" % (nm, (tt - t0) / 1e3))
