"""Global-latent prior -- mirror of the reference's models/score_sde/resnet.py
(Prior :124-218, ResBlockSEDrop :60-90, ResBlockSEClip :29-56, SE :16-27,
PriorSEDrop :221-224, PriorSEClip :226-229).  The forward is one C-ABI call
(`lion_global_prior_forward`, lion_b200/csrc/global_prior.cu)."""
import functools

import torch
import torch.nn as nn

from ... import _lib as L


class SE(nn.Module):
    def __init__(self, channel, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1, 1, bias=False), nn.ReLU(inplace=True),
                                nn.Conv2d(channel // reduction, channel, 1, 1, bias=False), nn.Sigmoid())


class ResBlockSEClip(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.conv1 = nn.Conv2d(input_dim * 2, output_dim, 1, 1)
        self.conv2 = nn.Conv2d(output_dim, output_dim, 1, 1)
        self.SE = SE(output_dim)

    def __repr__(self):
        return "ResBlockSEClip(%d, %d)" % (self.input_dim, self.output_dim)


class ResBlockSEDrop(nn.Module):
    def __init__(self, input_dim, output_dim, dropout):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.conv1 = nn.Conv2d(input_dim, output_dim, 1, 1)
        self.conv2 = nn.Conv2d(output_dim, output_dim, 1, 1)
        self.SE = SE(output_dim)
        self.dropout = nn.Dropout(dropout)
        self.dropout_ratio = dropout

    def __repr__(self):
        return "ResBlockSE_withdropout(%d, %d, drop=%f)" % (self.input_dim, self.output_dim, self.dropout_ratio)


class Prior(nn.Module):
    building_block = None

    def __init__(self, args, num_input_channels, *oargs, **kwargs):
        super().__init__()
        self.condition_input = kwargs.get('condition_input', False)
        self.cfg = oargs[0]
        self.clip_forge_enable = self.cfg.clipforge.enable
        self.num_scales = args.num_scales_dae
        self.num_input_channels = num_input_channels
        self.nf = nf = args.num_channels_dae
        if self.clip_forge_enable:
            self.clip_feat_mapping = nn.Conv1d(self.cfg.clipforge.feat_dim, self.nf, 1)
        self.mixed_prediction = args.mixed_prediction
        if self.mixed_prediction:
            raise NotImplementedError("lion_b200: sde.mixed_prediction is false in every shipped prior config")
        self.mixing_logit = None
        self.is_active = None
        self.embedding_dim = args.embedding_dim
        self.embedding_scale = args.embedding_scale
        assert args.embedding_type == 'positional', "lion_b200 implements the positional time embedding"
        self.temb_layer = nn.Sequential(nn.Conv2d(self.embedding_dim, self.embedding_dim * 4, 1, 1),
                                        nn.Conv2d(self.embedding_dim * 4, nf, 1, 1))
        self.input_layer = nn.Conv2d(num_input_channels, nf, 1, 1)
        self.all_modules = nn.ModuleList([self.building_block(nf, nf) for _ in range(args.num_cell_per_scale_dae)])
        self.output_layer = nn.Conv2d(nf, num_input_channels, 1, 1)

    def lion_desc(self):
        return [self.num_input_channels, self.nf, self.embedding_dim, len(self.all_modules), int(bool(self.clip_forge_enable)),
                self.cfg.clipforge.feat_dim, L.float_bits(self.embedding_scale)]

    def lion_params(self):
        ps = []
        if self.clip_forge_enable:
            ps += [self.clip_feat_mapping.weight, self.clip_feat_mapping.bias]
        ps += [self.temb_layer[0].weight, self.temb_layer[0].bias, self.temb_layer[1].weight, self.temb_layer[1].bias,
               self.input_layer.weight, self.input_layer.bias]
        for m in self.all_modules:
            ps += [m.conv1.weight, m.conv1.bias, m.conv2.weight, m.conv2.bias, m.SE.fc[0].weight, m.SE.fc[2].weight]
        return ps + [self.output_layer.weight, self.output_layer.bias]

    @torch.no_grad()
    def forward(self, x, t, **kwargs):
        """x [B, D, 1, 1], t [B] (or 0-dim) -> [B, D, 1, 1]   (resnet.py:195-218)"""
        shape = x.shape
        B = shape[0]
        xin = x.detach().to(torch.float32).contiguous().view(B, -1)
        t = t.detach().to(torch.float32)
        if t.dim() == 0:
            t = t.expand(1)
        if t.shape[0] == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        clip = None
        if self.clip_forge_enable:
            clip = kwargs['clip_feat'].detach().to(torch.float32).contiguous()
        m = L.model_for(self, L.KIND_GLOBAL_PRIOR, self.lion_desc(), self.lion_params())
        out = torch.empty_like(xin)
        with torch.cuda.device(xin.device):
            L.check(L.lib().lion_global_prior_forward(m.h, L.ptr(xin), L.ptr(t), L.ptr(clip), L.ptr(out), B, L.stream()),
                    "global_prior_forward")
        return out.view(shape)


class PriorSEDrop(Prior):
    def __init__(self, *args, **kwargs):
        self.building_block = functools.partial(ResBlockSEDrop, dropout=args[0].dropout)
        super().__init__(*args, **kwargs)


class PriorSEClip(Prior):
    building_block = ResBlockSEClip

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
