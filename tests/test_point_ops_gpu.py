"""The seven point/voxel operators through the C ABI, against (a) the CPU oracle and (b) the
reference's own CUDA kernels compiled into oracle/_ref (bit-exact for every index output)."""
import numpy as np
import pytest
import torch

from oracle import point_ops as P
from tests.util import assert_close, gen

pytestmark = pytest.mark.gpu


def _F():
    from lion_b200.third_party.pvcnn import functional as F
    return F


def _ref():
    from oracle.build_ref import load_ref
    return load_ref()


def cloud(seed, B, N, spread=0.5):
    return gen(seed, B, 3, N, scale=spread)


@pytest.mark.parametrize("B,N,r", [(2, 2048, 32), (3, 1024, 16), (2, 256, 8), (1, 64, 8)])
def test_voxel_coords_and_avg_voxelize(B, N, r):
    F = _F()
    coords = cloud(1, B, N)
    feats = gen(2, B, 7, N)
    nc_o, vox_o = P.voxel_coords(coords, r)
    nc, vox = F.voxel_coords(coords.cuda(), r)
    assert_close(nc, nc_o, 2e-6, "norm_coords")
    # bit-exact voxel index assignment given the same normalised coordinates
    assert torch.equal(vox.cpu(), P.round_to_voxel(nc.cpu()))
    mism = (vox.cpu() != vox_o).float().mean().item()
    assert mism < 2e-3, "voxel indices differ from the oracle on %.4f of the coordinates" % mism
    out = F.avg_voxelize(feats.cuda(), vox, r)
    out_o, ind_o, cnt_o = P.avg_voxelize(feats, vox.cpu(), r)
    assert_close(out, out_o, 1e-5, "avg_voxelize")
    ref = _ref()
    if ref is not None:
        o, ind, cnt = ref.avg_voxelize_forward(feats.cuda(), vox.contiguous(), r)
        assert torch.equal(ind.cpu(), ind_o) and torch.equal(cnt.cpu(), cnt_o)
        assert_close(out.view(B, 7, -1), o, 1e-5, "avg_voxelize vs reference kernel")


def test_avg_voxelize_collisions_and_single_point():
    F = _F()
    B, C, N, r = 2, 5, 33, 4
    feats = gen(3, B, C, N)
    vox = torch.zeros(B, 3, N, dtype=torch.int32)          # every point in voxel 0
    vox[1, :, 1:] = r - 1                                   # batch 1: one point alone, rest in the last voxel
    out = F.avg_voxelize(feats.cuda(), vox.cuda(), r)
    out_o, _, _ = P.avg_voxelize(feats, vox, r)
    assert_close(out, out_o, 1e-5, "avg_voxelize collisions")


@pytest.mark.parametrize("B,C,N,r", [(2, 32, 2048, 32), (2, 64, 1024, 16), (1, 130, 64, 8)])
def test_trilinear_devoxelize(B, C, N, r):
    F = _F()
    grid = gen(4, B, C, r, r, r)
    coords = torch.rand(B, 3, N, generator=torch.Generator().manual_seed(5)) * (r - 1)
    coords[:, :, :8] = torch.floor(coords[:, :, :8])       # exact lattice points: hi offset must be 0
    coords[:, :, 8] = r - 1                                 # upper corner of the grid
    coords[:, :, 9] = 0
    out = F.trilinear_devoxelize(grid.cuda(), coords.cuda(), r, False)
    assert_close(out, P.trilinear_devoxelize(grid, coords, r), 2e-6, "trilinear_devoxelize")
    ref = _ref()
    if ref is not None:
        o, inds, wgts = ref.trilinear_devoxelize_forward(r, True, coords.cuda(), grid.view(B, C, -1).cuda())
        assert_close(out, o, 1e-6, "devox vs reference kernel")
        idx_o, w_o = P.trilinear_corners(coords, r)
        assert torch.equal(inds.cpu().long(), idx_o)
        assert torch.equal(wgts.cpu(), w_o)


@pytest.mark.parametrize("B,N,M", [(3, 2048, 1024), (2, 1024, 256), (2, 256, 64), (4, 64, 16), (1, 700, 33)])
def test_furthest_point_sampling(B, N, M):
    from lion_b200.third_party.pvcnn.functional import furthest_point_sample_indices
    F = _F()
    coords = cloud(6, B, N)
    if N >= 256:
        coords[0, :, 100:140] = coords[0, :, 7:8]          # duplicates => exact ties in the distances
    idx = furthest_point_sample_indices(coords.cuda(), M)
    ref = _ref()
    if ref is not None:
        idx_r = ref.furthest_point_sampling(coords.cuda(), M)
        assert torch.equal(idx.cpu(), idx_r.cpu()), "FPS differs from the reference kernel"
    idx_o = P.furthest_point_sample_idx(coords, M)
    assert torch.equal(idx.cpu(), idx_o), "FPS differs from the oracle"
    centers = F.furthest_point_sample(coords.cuda(), M)
    assert torch.equal(centers.cpu(), P.gather(coords, idx_o))


@pytest.mark.parametrize("B,N,M,radius", [(2, 2048, 1024, 0.1), (2, 1024, 256, 0.2), (2, 256, 64, 0.4), (2, 64, 16, 0.8)])
def test_ball_query_and_grouping(B, N, M, radius):
    F = _F()
    pts = cloud(7, B, N, spread=0.3)
    ctr = pts[:, :, :M].clone()
    ctr[0, :, 0] = 50.0                                     # a centre with no neighbour: all zeros
    idx = F.ball_query(ctr.cuda(), pts.cuda(), radius, 32)
    ref = _ref()
    if ref is not None:
        idx_r = ref.ball_query(ctr.cuda(), pts.cuda(), radius, 32)
        assert torch.equal(idx.cpu(), idx_r.cpu()), "ball query differs from the reference kernel"
    idx_o = P.ball_query(ctr, pts, radius, 32)
    assert torch.equal(idx.cpu(), idx_o), "ball query differs from the oracle"
    assert (idx[0, 0] == 0).all()
    feats = gen(8, B, 9, N)
    g = F.grouping(feats.cuda(), idx)
    assert torch.equal(g.cpu(), P.grouping(feats, idx_o))


@pytest.mark.parametrize("B,C,N,M", [(2, 192, 64, 16), (2, 192, 256, 64), (2, 64, 1024, 256), (2, 17, 2048, 1024)])
def test_nearest_neighbor_interpolate(B, C, N, M):
    F = _F()
    pts = cloud(9, B, N)
    ctr = pts[:, :, :M].clone()                            # centres coincide with points: d = 0 -> clamp 1e-10
    cf = gen(10, B, C, M)
    out = F.nearest_neighbor_interpolate(pts.cuda(), ctr.cuda(), cf.cuda())
    out_o = P.nearest_neighbor_interpolate(pts, ctr, cf)
    assert_close(out, out_o, 2e-6, "3-NN interpolate")
    ref = _ref()
    if ref is not None:
        o, idx, w = ref.three_nearest_neighbors_interpolate_forward(pts.cuda(), ctr.cuda(), cf.cuda())
        idx_o, w_o = P.three_nn(pts, ctr)
        assert torch.equal(idx.cpu(), idx_o), "3-NN indices: oracle vs reference kernel"
        assert_close(out, o, 1e-6, "3-NN interpolate vs reference kernel")


def test_gather():
    F = _F()
    feats = gen(11, 3, 5, 100)
    idx = torch.randint(0, 100, (3, 40), generator=torch.Generator().manual_seed(1), dtype=torch.int32)
    assert torch.equal(F.gather(feats.cuda(), idx.cuda()).cpu(), P.gather(feats, idx))
