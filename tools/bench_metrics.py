"""Time the Chamfer metric kernels (CUDA events): the fused pairwise CD matrix and the drop-in
nearest-neighbour op driven the reference's way (one sample cloud expanded against all references).
python tools/bench_metrics.py [n_sample n_ref]   -> JSON lines"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lion_b200.utils.evaluation_metrics_fast import pairwise_CD, pairwise_EMD, distChamferCUDAnograd

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 405
N = 2048
g = torch.Generator().manual_seed(0)
s = torch.randn(ns, N, 3, generator=g).cuda()
r = torch.randn(nr, N, 3, generator=g).cuda()


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def composed():
    rows = []
    for i in range(ns):
        dl, dr = distChamferCUDAnograd(s[i:i + 1].expand(nr, -1, -1).contiguous(), r)
        rows.append(dl.mean(1) + dr.mean(1))
    return torch.stack(rows)


pair_evals = 2.0 * ns * nr * N * N          # both directions
for name, fn in (("fused lion_chamfer_pairwise", lambda: pairwise_CD(s, r)), ("drop-in op, reference-style loop", composed)):
    ms = timed(fn)
    print(json.dumps({"what": name, "n_sample": ns, "n_ref": nr, "points": N, "ms": round(ms, 3),
                      "cloud_pairs_per_s": round(ns * nr / (ms * 1e-3)), "gflops_8_per_point_pair": round(8 * pair_evals / (ms * 1e-3) / 1e9)}))
assert torch.allclose(pairwise_CD(s, r), composed(), rtol=1e-5, atol=0)
# approximate EMD: 30 exp-weighted N x M passes per pair (MUFU-bound); a smaller block of the matrix
ne = max(1, ns // 8)
ms = timed(lambda: pairwise_EMD(s[:ne], r), iters=1)
print(json.dumps({"what": "fused lion_emd_pairwise", "n_sample": ne, "n_ref": nr, "points": N, "ms": round(ms, 3),
                  "cloud_pairs_per_s": round(ne * nr / (ms * 1e-3)), "gexp_per_s": round(30.0 * ne * nr * N * N / (ms * 1e-3) / 1e9)}))
