"""Per-shape comparison of the 3x3x3 voxel convolution at B=32 (SURVEY.md App. A shape classes):
lion_b200's tcgen05 kernel (CUDA events via lion_bench_conv, kernel only, packed layouts) against
torch.nn.functional.conv3d on the same box = cuDNN with TF32 allowed and cudnn.benchmark=True, i.e.
exactly how the reference runs nn.Conv3d (models/pvcnn2_ada.py:211-222, utils/utils.py:472).

    python tools/bench_conv_vs_cudnn.py > profiles/r02_conv_vs_cudnn.jsonl

Both sides: dense random input, bias, 10 timed iterations after 3 warm-ups, inputs (>= 134 MB at the large
shapes) larger than... no L2 flush is needed for r = 32; the r = 8 shapes (8-17 MB) are L2-resident on
BOTH sides (as they are inside the real step, where producer and consumer run back to back)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from lion_b200 import _lib as L

B = int(os.environ.get("B", "32"))
SHAPES = [  # (cin, cout, r, launches per step, label)
    (4, 32, 32, 1, "sa0.0 conv1"), (32, 32, 32, 3, "sa0.x conv"), (128, 64, 16, 1, "sa1.0 conv1"),
    (64, 64, 16, 1, "sa1.0 conv2"), (192, 128, 8, 1, "sa2.0 conv1"), (128, 128, 8, 13, "r=8 128->128"),
    (128, 128, 16, 4, "fp2 r=16"), (64, 64, 32, 4, "fp3 r=32"),
]


def time_cudnn(cin, cout, r, channels_last=False):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, cin, r, r, r, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, device=dev, generator=g) * (27 * cin) ** -0.5
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last_3d)
        w = w.contiguous(memory_format=torch.channels_last_3d)
    with torch.no_grad():
        for _ in range(3):
            y = F.conv3d(x, w, b, padding=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = F.conv3d(x, w, b, padding=1)
        e1.record()
        torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    return e0.elapsed_time(e1) / 10.0


def main():
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True          # torch default for convolutions; stated for the record
    tot_l, tot_c = 0.0, 0.0
    for cin, cout, r, n, label in SHAPES:
        ms, fl = C.c_float(), C.c_double()
        L.check(L.lib().lion_bench_conv(L.ctx(), 27, cin, cout, r, B, 10, 3, C.byref(ms), C.byref(fl), L.stream()), label)
        ms_c = time_cudnn(cin, cout, r)
        ms_cl = time_cudnn(cin, cout, r, channels_last=True)
        best = min(ms_c, ms_cl)
        tot_l += ms.value * n
        tot_c += best * n
        print(json.dumps({"shape": label, "cin": cin, "cout": cout, "r": r, "B": B, "launches_per_step": n,
                          "lion_ms": round(ms.value, 4), "lion_tflops": round(fl.value / ms.value / 1e9, 1),
                          "cudnn_tf32_ms_ncdhw": round(ms_c, 4), "cudnn_tf32_ms_channels_last": round(ms_cl, 4),
                          "cudnn_tflops_best": round(fl.value / best / 1e9, 1), "speedup_vs_cudnn_best": round(best / ms.value, 2)}))
    print(json.dumps({"sum_ms_per_step": {"lion": round(tot_l, 3), "cudnn_tf32_best_layout": round(tot_c, 3)},
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
