"""Point-operator backend that runs the REFERENCE's own CUDA kernels (oracle/_ref/_pvcnn_backend.so,
built by oracle/build_ref.py from third_party/pvcnn/functional/src) on CUDA tensors, with the same
function names as oracle/point_ops.py, so that oracle/net.py -- the restatement of the reference's
PyTorch modules -- becomes a GPU port of the reference's eager path: cuDNN / cuBLAS through torch for
the dense layers + the reference's point kernels.

TEST INFRASTRUCTURE ONLY: used by bench.py's `gpu_baseline` leg (`--impl reference-gpu`) as the thing
to compare against, never by the product path.  The wrappers follow third_party/pvcnn/functional/*.py
(voxelization.py:13-28, devoxelization.py:10-27, sampling.py:11-54, ball_query.py:8-20,
grouping.py:9-30, interpolatation.py:14-30) and Voxelization.forward (models/pvcnn2_ada.py:173-188).
"""
import torch

from . import build_ref

_mod = None


def _ref():
    global _mod
    if _mod is None:
        _mod = build_ref.load_ref()
        if _mod is None:
            raise RuntimeError("oracle/_ref/_pvcnn_backend.so is missing (python oracle/build_ref.py)")
    return _mod


def voxel_coords(coords, r, normalize=True, eps=0.0):
    nc = coords - coords.mean(2, keepdim=True)
    if normalize:
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    return nc, torch.round(nc).to(torch.int32)


def avg_voxelize(features, vox, r):
    b, c = features.shape[:2]
    out, ind, cnt = _ref().avg_voxelize_forward(features.contiguous(), vox.contiguous(), r)
    return out.view(b, c, r, r, r), ind, cnt


def trilinear_devoxelize(grid, coords, r, is_training=False):
    b, c = grid.shape[:2]
    out, _, _ = _ref().trilinear_devoxelize_forward(r, is_training, coords.contiguous(), grid.contiguous().view(b, c, -1))
    return out


def gather(features, idx):
    return _ref().gather_features_forward(features.contiguous(), idx.contiguous())


def furthest_point_sample(coords, m):
    coords = coords.contiguous()
    return gather(coords, _ref().furthest_point_sampling(coords, m))


def ball_query(centers, points, radius, k):
    return _ref().ball_query(centers.contiguous(), points.contiguous(), radius, k)


def grouping(features, idx):
    return _ref().grouping_forward(features.contiguous(), idx.contiguous())


def nearest_neighbor_interpolate(points, centers, centers_features):
    out, _, _ = _ref().three_nearest_neighbors_interpolate_forward(points.contiguous(), centers.contiguous(),
                                                                   centers_features.contiguous())
    return out
