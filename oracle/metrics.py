"""CPU restatement of the Chamfer nearest-neighbour kernel and the pairwise CD matrix.

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.

  chamfer_forward   third_party/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-143 (NmDistanceKernel,
                    both directions as launched by chamfer_cuda_forward)
  pairwise_cd       utils/evaluation_metrics_fast.py:272-340 (_pairwise_EMD_CD_ / _pairwise_EMD_CD_sub,
                    metric 'CD': dl.mean(1) + dr.mean(1) of one sample cloud against every reference)

The kernel's distance is d = x2*x2 + y2*y2 + z2*z2 on fp32 differences; nvcc contracts it to
t = y2*y2; t = fma(x2,x2,t); t = fma(z2,z2,t) (read off the SASS of the reference source built for
sm_100a), emulated here exactly: fp32 products are exact in float64, and each fma rounds once.
Ties keep the lowest index (first candidate unconditionally, later ones only when strictly
smaller, also across the kernel's 512-point chunks).

Parity status: pinned on the GPU box against oracle/_ref/chamfer_3D.so (the reference's own
extension built by oracle/build_ref.py) in tests/test_metrics_gpu.py; no golden vector exists
in the reference for this path.
"""
import numpy as np


def _nn(q, c):
    """q [n,3], c [m,3] fp32 -> (dist [n] fp32, idx [n] int32)"""
    q = np.asarray(q, np.float32)
    c = np.asarray(c, np.float32)
    dx = (c[None, :, 0] - q[:, None, 0]).astype(np.float32).astype(np.float64)
    dy = (c[None, :, 1] - q[:, None, 1]).astype(np.float32).astype(np.float64)
    dz = (c[None, :, 2] - q[:, None, 2]).astype(np.float32).astype(np.float64)
    t = (dy * dy).astype(np.float32).astype(np.float64)          # FMUL
    t = (dx * dx + t).astype(np.float32).astype(np.float64)      # FFMA (single rounding)
    d = (dz * dz + t).astype(np.float32)                         # FFMA
    idx = np.argmin(d, axis=1).astype(np.int32)                  # first (lowest-index) minimum
    return d[np.arange(q.shape[0]), idx], idx


def chamfer_forward(xyz1, xyz2):
    """xyz1 [B,N,3], xyz2 [B,M,3] -> dist1 [B,N], dist2 [B,M], idx1, idx2 (int32)"""
    B = xyz1.shape[0]
    d1, i1, d2, i2 = [], [], [], []
    for b in range(B):
        d, i = _nn(xyz1[b], xyz2[b]); d1.append(d); i1.append(i)
        d, i = _nn(xyz2[b], xyz1[b]); d2.append(d); i2.append(i)
    return np.stack(d1), np.stack(d2), np.stack(i1), np.stack(i2)


def pairwise_cd(samples, refs):
    """[Ns,N,3], [Nr,M,3] -> [Ns,Nr]; means accumulated in float64 (the reference: torch .mean(dim=1) in fp32)."""
    out = np.zeros((samples.shape[0], refs.shape[0]), np.float64)
    for i in range(samples.shape[0]):
        for j in range(refs.shape[0]):
            dl, _ = _nn(samples[i], refs[j])
            dr, _ = _nn(refs[j], samples[i])
            out[i, j] = dl.astype(np.float64).mean() + dr.astype(np.float64).mean()
    return out
