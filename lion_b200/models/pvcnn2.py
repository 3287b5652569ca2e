"""The NON-Ada PVCNN2 blocks -- host-side mirror of the reference's models/pvcnn2.py, which the VAE's global style
encoder (models/shapelatent_modules.py: PointNetPlusEncoder) is built from: same class names, constructor
arguments and state_dict keys (`voxel_layers.{0,1,4,5,6}`, `point_features.layers.{0,1}`, `mlps.0.layers.{3i,3i+1}`,
`attn.to_qkv/to_out`), so released VAE checkpoints load unchanged.  Normalisation is plain nn.GroupNorm(8, C)
instead of AdaGN and no style vector exists; in the library that is the same kernel path with the style Linear
fixed at (factor, bias) = (1, 0) (descriptor style_dim = 0).  No PyTorch fallback.

  SharedMLP                       pvcnn2.py:117-138
  PVConv                          pvcnn2.py:170-247
  PointNetSAModule                pvcnn2.py:288-351
  create_pointnet2_sa_components  pvcnn2.py:440-509
SE3d, LinearAttention, Swish, BallQuery and Voxelization are identical in both files and are shared."""
import torch
import torch.nn as nn

from .. import _lib as L
from .pvcnn2_ada import SE3d, LinearAttention, Swish, BallQuery, Voxelization, _run, _f32c  # noqa: F401


def _gn_params(seq):
    """[conv.w, conv.b, gn.w, gn.b] * n of a (conv, GroupNorm, Swish) * n Sequential"""
    ps = []
    for l in seq:
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.GroupNorm)):
            ps += [l.weight, l.bias]
    return ps


class SharedMLP(nn.Module):
    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        conv = nn.Conv1d if dim == 1 else nn.Conv2d
        if not isinstance(out_channels, (list, tuple)):
            out_channels = [out_channels]
        self.in_channels, self.out_channels = in_channels, list(out_channels)
        layers = []
        for oc in out_channels:
            layers += [conv(in_channels, oc, 1), nn.GroupNorm(8, oc), Swish()]
            in_channels = oc
        self.layers = nn.Sequential(*layers)

    def lion_params(self):
        return _gn_params(self.layers)

    @torch.no_grad()
    def _fwd(self, x):
        shape = x.shape
        x = _f32c(x).reshape(shape[0], shape[1], -1)
        B, _, R = x.shape
        m = L.model_for(self, L.KIND_SHARED_MLP, [self.in_channels, 0, len(self.out_channels)] + self.out_channels, self.lion_params())
        out = torch.empty(B, self.out_channels[-1], R, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _run(L.lib().lion_shared_mlp_fwd, m.h, L.ptr(x), None, L.ptr(out), B, R, L.stream())
        return out.reshape(B, self.out_channels[-1], *shape[2:])

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            return (self._fwd(inputs[0]), *inputs[1:])
        return self._fwd(inputs)


class PVConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, resolution, normalize=1, eps=0, with_se=False,
                 add_point_feat=True, attention=False, dropout=0.1, verbose=True):
        super().__init__()
        assert kernel_size == 3 and with_se and add_point_feat and normalize and eps == 0, \
            "lion_b200 implements the PVConv variant LION instantiates (3x3x3, SE, point branch, normalised coords)"
        self.in_channels, self.out_channels, self.resolution = in_channels, out_channels, resolution
        self.voxelization = Voxelization(resolution, normalize=normalize, eps=eps)
        self.voxel_layers = nn.Sequential(
            nn.Conv3d(in_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2), nn.GroupNorm(8, out_channels),
            Swish(), nn.Dropout(dropout),
            nn.Conv3d(out_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2), nn.GroupNorm(8, out_channels),
            SE3d(out_channels))
        self.attn = LinearAttention(out_channels, verbose=verbose) if attention else None
        self.point_features = SharedMLP(in_channels, out_channels)
        self.add_point_feat = add_point_feat

    def lion_params(self):
        v = self.voxel_layers
        ps = [v[0].weight, v[0].bias, v[1].weight, v[1].bias, v[4].weight, v[4].bias, v[5].weight, v[5].bias] + v[6].lion_params()
        if self.attn is not None:
            ps += self.attn.lion_params()
        return ps + self.point_features.lion_params()

    @torch.no_grad()
    def forward(self, inputs):
        features, coords_input, time_emb = inputs[0], inputs[1], inputs[2]
        coords = coords_input[:, :3] if coords_input.shape[1] > 3 else coords_input
        assert features.shape[0] == coords.shape[0] and features.shape[2] == coords.shape[2] and coords.shape[1] == 3
        features, coords = _f32c(features), _f32c(coords)
        B, _, N = features.shape
        m = L.model_for(self, L.KIND_PVCONV, [self.in_channels, self.out_channels, self.resolution, int(self.attn is not None), 0],
                        self.lion_params())
        out = torch.empty(B, self.out_channels, N, device=features.device, dtype=torch.float32)
        with torch.cuda.device(features.device):
            _run(L.lib().lion_pvconv_fwd, m.h, L.ptr(features), L.ptr(coords), None, L.ptr(out), B, N, L.stream())
        return out, coords_input, time_emb


class PointNetSAModule(nn.Module):
    def __init__(self, num_centers, radius, num_neighbors, in_channels, out_channels, include_coordinates=True):
        super().__init__()
        assert include_coordinates and not isinstance(radius, (list, tuple)), \
            "lion_b200 implements the single-scale SA module LION instantiates"
        out_channels = list(out_channels) if isinstance(out_channels, (list, tuple)) else [out_channels]
        self.num_centers, self.radius, self.num_neighbors, self.in_channels = num_centers, radius, num_neighbors, in_channels
        self.out_channels = out_channels[-1]
        self.groupers = nn.ModuleList([BallQuery(radius=radius, num_neighbors=num_neighbors, include_coordinates=True)])
        self.mlps = nn.ModuleList([SharedMLP(in_channels=in_channels + 3, out_channels=out_channels, dim=2)])

    def lion_desc(self):
        oc = self.mlps[0].out_channels
        return [self.in_channels, self.num_centers, L.float_bits(self.radius), self.num_neighbors, 0, len(oc)] + oc

    def lion_params(self):
        return self.mlps[0].lion_params()

    @torch.no_grad()
    def forward(self, inputs):
        features, coords, time_emb = inputs[0], inputs[1], inputs[2]
        if coords.shape[1] > 3:
            coords = coords[:, :3]
        features, coords = _f32c(features), _f32c(coords)
        B, _, N = features.shape
        M = self.num_centers
        m = L.model_for(self, L.KIND_SA, self.lion_desc(), self.lion_params())
        out = torch.empty(B, self.out_channels, M, device=features.device, dtype=torch.float32)
        centers = torch.empty(B, 3, M, device=features.device, dtype=torch.float32)
        with torch.cuda.device(features.device):
            _run(L.lib().lion_sa_module_fwd, m.h, L.ptr(features), L.ptr(coords), None, L.ptr(out), L.ptr(centers), B, N, L.stream())
        if time_emb is not None and type(time_emb) is not dict:
            time_emb = time_emb[:, :, :M]
        return out, centers, time_emb

    def extra_repr(self):
        return f'num_centers={self.num_centers}, out_channels={self.out_channels}'


def create_pointnet2_sa_components(sa_blocks, extra_feature_channels, input_dim=3, embed_dim=64, use_att=False, force_att=0,
                                   dropout=0.1, with_se=False, normalize=True, eps=0, has_temb=1, width_multiplier=1,
                                   voxel_resolution_multiplier=1, verbose=True):
    """Module table of the SA half (pvcnn2.py:440-509), including its quirk that levels after the first keep only their
    first PVConv.  Returns (sa_layers, sa_in_channels, channels_sa_features, num_centers)."""
    assert width_multiplier == 1 and voxel_resolution_multiplier == 1 and not force_att
    in_channels = extra_feature_channels + input_dim
    sa_layers, sa_in_channels = [], []
    num_centers = None
    for c, (conv_configs, sa_configs) in enumerate(sa_blocks):
        k = 0
        sa_in_channels.append(in_channels)
        blocks = []
        if conv_configs is not None:
            out_channels, num_blocks, voxel_resolution = conv_configs
            for p in range(num_blocks):
                attention = (c + 1) % 2 == 0 and use_att and p == 0
                if c == 0 or k == 0:
                    cin = in_channels if c == 0 else in_channels + embed_dim * has_temb
                    blocks.append(PVConv(cin, out_channels, kernel_size=3, resolution=voxel_resolution, attention=attention,
                                         dropout=dropout, with_se=with_se, normalize=normalize, eps=eps, verbose=verbose))
                in_channels = out_channels
                k += 1
            extra_feature_channels = in_channels
        num_centers, radius, num_neighbors, out_channels = sa_configs
        blocks.append(PointNetSAModule(num_centers=num_centers, radius=radius, num_neighbors=num_neighbors,
                                       in_channels=extra_feature_channels + (embed_dim * has_temb if k == 0 else 0),
                                       out_channels=list(out_channels), include_coordinates=True))
        in_channels = extra_feature_channels = blocks[-1].out_channels
        sa_layers.append(blocks[0] if len(blocks) == 1 else nn.Sequential(*blocks))
    return sa_layers, sa_in_channels, in_channels, 1 if num_centers is None else num_centers
