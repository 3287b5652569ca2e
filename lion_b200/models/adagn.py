"""AdaGN: GroupNorm(8, C) followed by a style-conditioned affine (reference: models/adagn.py:19-65).

Parameter names match the reference (`norm.weight/bias`, `emd.weight/bias`; rows [0:C] of
`emd` are the factor, [C:2C] the bias, adagn.py:62).  On the hot path the parent block
(SharedMLP / PVConv) hands these tensors to the fused kernels, which take the GroupNorm
statistics from the producing convolution's epilogue and fold GroupNorm + style affine into a
single per-(sample, channel) scale/shift; a stand-alone call (`forward`) runs the same affine
through `lion_adagn_fwd` with a separate statistics kernel.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .dense import dense


class AdaGN(nn.Module):
    def __init__(self, ndim, cfg, n_channel):
        super().__init__()
        style_dim = cfg.latent_pts.style_dim
        init_scale = cfg.latent_pts.ada_mlp_init_scale
        self.ndim = ndim
        self.n_channel = n_channel
        self.style_dim = style_dim
        self.out_dim = n_channel * 2
        self.norm = nn.GroupNorm(8, n_channel)
        self.emd = dense(style_dim, n_channel * 2, init_scale=init_scale)
        self.emd.bias.data[:n_channel] = 1
        self.emd.bias.data[n_channel:] = 0

    def __repr__(self):
        return f"AdaGN(GN(8, {self.n_channel}), Linear({self.style_dim}, {self.out_dim}))"

    def lion_params(self):
        return [self.norm.weight, self.norm.bias, self.emd.weight, self.emd.bias]

    @torch.no_grad()
    def forward(self, image, style):
        """image [B, C, ...] (ndim trailing dims), style [B, style_dim] -> same shape as image."""
        assert style.dim() == 2, 'get {} {}'.format(style.shape, len(style.shape))
        assert image.dim() == self.ndim + 2, 'get {} {}'.format(image.shape, len(image.shape))
        shape = image.shape
        x = image.detach().to(torch.float32).contiguous().view(shape[0], shape[1], -1)
        style = style.detach().to(torch.float32).contiguous()
        m = L.model_for(self, L.KIND_ADAGN, [self.n_channel, self.style_dim], self.lion_params())
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.lib().lion_adagn_fwd(m.h, L.ptr(x), L.ptr(style), L.ptr(out), shape[0], x.shape[2], L.stream()), "adagn_fwd")
        return out.view(shape)
