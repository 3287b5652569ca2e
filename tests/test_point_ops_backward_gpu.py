"""backward passes of the five differentiable point/voxel operators, through torch.autograd on the
library's Functions, against the reference's OWN backward kernels (oracle/_ref/_pvcnn_backend.so: vox.cu:86-110,
trilinear_devox.cu:119-162, grouping.cu:58-77, neighbor_interpolate.cu:145-170, sampling.cu:52-66) fed with the same
saved indices.  Gather-type gradients are bit-identical; scatter-adds (fp32 atomics on both sides) agree to 1e-6."""
import pytest
import torch

from tests.util import assert_close, gen

pytestmark = pytest.mark.gpu


def _F():
    from lion_b200.third_party.pvcnn import functional as F
    return F


def _ref():
    from oracle.build_ref import load_ref
    ref = load_ref()
    assert ref is not None, "oracle/_ref/_pvcnn_backend.so is missing"
    return ref


@pytest.mark.parametrize("B,C,N,r", [(2, 16, 2048, 32), (3, 7, 300, 8), (32, 64, 2048, 32)])
def test_avg_voxelize_backward(B, C, N, r):
    F, ref = _F(), _ref()
    feats = gen(1, B, C, N).cuda().requires_grad_(True)
    vox = torch.randint(0, r, (B, 3, N), generator=torch.Generator().manual_seed(2), dtype=torch.int32).cuda()
    vox[0, :, : N // 2] = 1                                   # many points in one voxel
    out = F.avg_voxelize(feats, vox, r)
    gy = gen(3, B, C, r, r, r).cuda()
    out.backward(gy)
    _, ind, cnt = ref.avg_voxelize_forward(feats.detach(), vox, r)
    want = ref.avg_voxelize_backward(gy.view(B, C, -1).contiguous(), ind, cnt)
    assert torch.equal(feats.grad, want), "avg_voxelize backward is a pure gather: must be bit-identical"


@pytest.mark.parametrize("B,C,N,r", [(2, 32, 2048, 32), (2, 5, 100, 4), (32, 64, 1024, 16)])
def test_trilinear_devoxelize_backward(B, C, N, r):
    F, ref = _F(), _ref()
    grid = gen(4, B, C, r, r, r).cuda().requires_grad_(True)
    coords = (torch.rand(B, 3, N, generator=torch.Generator().manual_seed(5)) * (r - 1)).cuda()
    out = F.trilinear_devoxelize(grid, coords, r, True)
    gy = gen(6, B, C, N).cuda()
    out.backward(gy)
    _, inds, wgts = ref.trilinear_devoxelize_forward(r, True, coords, grid.detach().view(B, C, -1))
    want = ref.trilinear_devoxelize_backward(gy, inds, wgts, r)
    assert_close(grid.grad.view(B, C, -1), want, 1e-6, "trilinear_devoxelize backward vs reference kernel")


@pytest.mark.parametrize("B,C,N,M,U", [(2, 35, 2048, 1024, 32), (3, 9, 64, 16, 32)])
def test_grouping_and_gather_backward(B, C, N, M, U):
    F, ref = _F(), _ref()
    feats = gen(7, B, C, N).cuda().requires_grad_(True)
    idx = torch.randint(0, N, (B, M, U), generator=torch.Generator().manual_seed(8), dtype=torch.int32).cuda()
    g = F.grouping(feats, idx)
    gy = gen(9, B, C, M, U).cuda()
    g.backward(gy)
    assert_close(feats.grad, ref.grouping_backward(gy, idx, N), 1e-6, "grouping backward vs reference kernel")
    feats2 = gen(10, B, C, N).cuda().requires_grad_(True)
    idx1 = torch.randint(0, N, (B, M), generator=torch.Generator().manual_seed(11), dtype=torch.int32).cuda()
    o = F.gather(feats2, idx1)
    gy1 = gen(12, B, C, M).cuda()
    o.backward(gy1)
    assert_close(feats2.grad, ref.gather_features_backward(gy1, idx1, N), 1e-6, "gather backward vs reference kernel")


@pytest.mark.parametrize("B,C,N,M", [(2, 192, 256, 64), (2, 17, 2048, 1024)])
def test_nearest_neighbor_interpolate_backward(B, C, N, M):
    F, ref = _F(), _ref()
    pts = gen(13, B, 3, N, scale=0.5).cuda()
    ctr = pts[:, :, :M].contiguous()
    cf = gen(14, B, C, M).cuda().requires_grad_(True)
    out = F.nearest_neighbor_interpolate(pts, ctr, cf)
    gy = gen(15, B, C, N).cuda()
    out.backward(gy)
    _, idx, wgt = ref.three_nearest_neighbors_interpolate_forward(pts, ctr, cf.detach())
    want = ref.three_nearest_neighbors_interpolate_backward(gy, idx, wgt, M)
    assert_close(cf.grad, want, 1e-6, "3-NN interpolate backward vs reference kernel")


def test_no_grad_path_is_unchanged():
    """The sampling path (no_grad) keeps the plain forward calls: no autograd graph, same values."""
    F = _F()
    feats = gen(16, 2, 8, 512).cuda().requires_grad_(True)
    idx = torch.randint(0, 512, (2, 64, 32), generator=torch.Generator().manual_seed(17), dtype=torch.int32).cuda()
    with torch.no_grad():
        a = F.grouping(feats, idx)
    b = F.grouping(feats, idx)
    assert not a.requires_grad and b.requires_grad and torch.equal(a, b.detach())
