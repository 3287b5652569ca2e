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

  emd_approx        third_party/PyTorchEMD/cuda/emd_kernel.cu:23-170 (approxmatch) + :196-246 (matchcost),
                    emd_nograd.py:9-45: float64 restatement (exact exp instead of __expf, so it agrees
                    with the kernels to ~1e-4, not bit for bit); pinned on the GPU box against
                    oracle/_ref/emd_ext.so = the reference's own kernels (built with oracle/shim/).
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


def emd_approx(xyz1, xyz2):
    """xyz1 [B,N,3], xyz2 [B,M,3] -> cost [B] (sum d^2 * match; divide by N for earth_mover_distance_nograd)."""
    xyz1 = np.asarray(xyz1, np.float64)
    xyz2 = np.asarray(xyz2, np.float64)
    B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    out = np.zeros(B)
    for b in range(B):
        d2 = ((xyz2[b][None, :, :] - xyz1[b][:, None, :]) ** 2).sum(-1)          # [n, m]
        multiL, multiR = (1.0, float(n // m)) if n >= m else (float(m // n), 1.0)
        remainL = np.full(n, multiL)
        remainR = np.full(m, multiR)
        match = np.zeros((n, m))
        for j in range(7, -3, -1):
            level = -(4.0 ** j) if j != -2 else 0.0
            e = np.exp(level * d2)
            ratioL = remainL / (1e-9 + e @ remainR)
            sumr = (e.T @ ratioL) * remainR
            consumption = np.minimum(remainR / (sumr + 1e-9), 1.0)
            ratioR = consumption * remainR
            remainR = np.maximum(0.0, remainR - sumr)
            w = e * ratioL[:, None] * ratioR[None, :]
            match += w
            remainL = np.maximum(0.0, remainL - w.sum(1))
        out[b] = (d2 * match).sum()
    return out
