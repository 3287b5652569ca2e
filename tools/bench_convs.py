"""Time the convolution kernel alone on every shape of the PVCNN2 prior step (B=32):
python tools/bench_convs.py  -> table + JSON lines (CUDA events via lion_bench_conv)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lion_b200 import _lib as L

B = int(os.environ.get("B", "32"))
SHAPES = [  # (ntaps, cin, cout, r_or_rows, launches per step, label)
    (27, 4, 32, 32, 1, "sa0.0 conv1"), (27, 32, 32, 32, 3, "sa0.x conv"), (27, 128, 64, 16, 1, "sa1.0 conv1"),
    (27, 64, 64, 16, 1, "sa1.0 conv2"), (27, 192, 128, 8, 1, "sa2.0 conv1"), (27, 128, 128, 8, 13, "r=8 128->128"),
    (27, 128, 128, 16, 4, "fp2 r=16"), (27, 64, 64, 32, 4, "fp3 r=32"),
    (1, 36, 32, 32768, 1, "SA0 mlp0"), (1, 32, 64, 32768, 1, "SA0 mlp1"), (1, 68, 64, 8192, 1, "SA1 mlp0"),
    (1, 64, 128, 8192, 1, "SA1 mlp1"), (1, 196, 128, 2048, 1, "fp3 mlp0"), (1, 128, 128, 2048, 1, "fp3 mlp1"),
    (1, 64, 384, 1024, 1, "attn qkv"),
]
ONLY = os.environ.get("ONLY")  # e.g. ONLY="fp3 r=32" to time one shape (ncu captures)
if ONLY:
    SHAPES = [s for s in SHAPES if s[5] == ONLY]
if os.environ.get("TAPS"):           # TAPS=1 -> only the 1x1 shapes, TAPS=27 -> only the 3x3x3 ones
    SHAPES = [s for s in SHAPES if s[0] == int(os.environ["TAPS"])]
torch.cuda.init()
ITERS = int(os.environ.get("ITERS", "10"))      # ITERS=3000 CLOCKS=1: long enough for nvidia-smi to see the clock under load


def sample_clocks(stop, rows):
    import subprocess
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    for line in p.stdout:
        rows.append([float(c) for c in line.split(",")])
        if stop.is_set():
            break
    p.terminate()


tot = 0.0
for nt, ci, co, r, n, label in SHAPES:
    ms, fl = C.c_float(), C.c_double()
    rows, stop, th = [], None, None
    if os.environ.get("CLOCKS"):
        import threading
        stop = threading.Event()
        th = threading.Thread(target=sample_clocks, args=(stop, rows), daemon=True)
        th.start()
    L.check(L.lib().lion_bench_conv(L.ctx(), nt, ci, co, r, B, ITERS, 2, C.byref(ms), C.byref(fl), L.stream()), label)
    tf = fl.value / (ms.value * 1e-3) / 1e12
    tot += ms.value * n
    rec = {"shape": label, "ntaps": nt, "cin": ci, "cout": co, "r_or_rows": r, "B": B, "ms": round(ms.value, 4),
           "tflops_algorithmic": round(tf, 1), "launches_per_step": n}
    if th is not None:
        stop.set()
        th.join(timeout=2)
        hot = sorted(x[0] for x in rows if x[1] > 300) or [x[0] for x in rows]
        rec["sm_mhz_under_load_median"] = hot[len(hot) // 2] if hot else None
        rec["power_w_max"] = max((x[1] for x in rows), default=None)
        rec["clock_samples"] = len(hot)
    print(json.dumps(rec))
print(json.dumps({"sum_ms_per_step_listed": round(tot, 3)}))
