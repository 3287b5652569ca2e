"""The VAE's global style encoder -- host-side mirror of the reference's models/shapelatent_modules.py
(PointNetPlusEncoder :13-52): two set-abstraction levels of the NON-Ada PVCNN blocks (models/pvcnn2.py), max over the
256 remaining points, Linear(64 -> 2*zdim).  The whole forward is ONE C-ABI call (lion_style_encoder_forward,
lion_b200/csrc/net.cu: style_enc_forward); the module tree only owns the parameters under the reference's names
(`layers.{level}.{block}...`, `mlp.weight/bias`)."""
import torch
import torch.nn as nn

from .. import _lib as L
from .pvcnn2 import create_pointnet2_sa_components, PVConv, PointNetSAModule


def _blocks_of(layer):
    return list(layer) if isinstance(layer, nn.Sequential) else [layer]


class PointNetPlusEncoder(nn.Module):
    sa_blocks = [
        [[32, 2, 32], [1024, 0.1, 32, [32, 32]]],
        [[32, 1, 16], [256, 0.2, 32, [32, 64]]],
    ]
    force_att = 0

    def __init__(self, zdim, input_dim, extra_feature_channels=0, args={}):
        super().__init__()
        assert extra_feature_channels == 0 and input_dim == 3
        layers, _, channels_sa_features, _ = create_pointnet2_sa_components(
            self.sa_blocks, extra_feature_channels, input_dim=input_dim, embed_dim=0, force_att=self.force_att, use_att=True,
            with_se=True)
        self.mlp = nn.Linear(channels_sa_features, zdim * 2)
        self.zdim = zdim
        self.input_dim = input_dim
        self.layers = nn.ModuleList(layers)
        self.voxel_dim = [n[1][-1][-1] for n in self.sa_blocks]

    def lion_desc(self):
        d = [self.input_dim, self.zdim, 1, len(self.sa_blocks)]
        for (oc, nblk, res), (m, radius, k, mlp) in self.sa_blocks:
            d += [1, oc, nblk, res, m, L.float_bits(radius), k, len(mlp)] + list(mlp)
        return d

    def lion_params(self):
        ps = []
        for layer in self.layers:
            for blk in _blocks_of(layer):
                assert isinstance(blk, (PVConv, PointNetSAModule))
                ps += blk.lion_params()
        return ps + [self.mlp.weight, self.mlp.bias]

    @torch.no_grad()
    def forward(self, x):
        """x: [B,N,3] -> {'mu_1d': [B,zdim], 'sigma_1d': [B,zdim] (log sigma)}"""
        x = x.detach().to(torch.float32).contiguous()
        B, N, D = x.shape
        assert D == self.input_dim
        m = L.model_for(self, L.KIND_STYLE_ENC, self.lion_desc(), self.lion_params())
        out = torch.empty(B, 2 * self.zdim, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.check(L.lib().lion_style_encoder_forward(m.h, L.ptr(x), L.ptr(out), B, N, L.stream()), "style_encoder_forward")
        return {'mu_1d': out[:, :self.zdim], 'sigma_1d': out[:, self.zdim:]}
