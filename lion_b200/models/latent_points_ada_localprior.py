"""The latent-point score network -- mirror of the reference's
models/latent_points_ada_localprior.py (PVCNN2Prior :16-83)."""
import torch

from .latent_points_ada import PVCNN2Unet


class PVCNN2Prior(PVCNN2Unet):
    sa_blocks = [
        ((32, 2, 32), (1024, 0.1, 32, (32, 64))),
        ((64, 3, 16), (256, 0.2, 32, (64, 128))),
        ((128, 3, 8), (64, 0.4, 32, (128, 128))),
        (None, (16, 0.8, 32, (128, 128, 128))),
    ]
    fp_blocks = [
        ((128, 128), (128, 3, 8)),
        ((128, 128), (128, 3, 8)),
        ((128, 128), (128, 2, 16)),
        ((128, 128, 64), (64, 2, 32)),
    ]

    def __init__(self, args, num_input_channels, cfg):
        self.clip_forge_enable = cfg.clipforge.enable
        num_classes = cfg.shapelatent.latent_dim + cfg.ddpm.input_dim
        self.num_classes = num_classes
        self.num_points = cfg.data.tr_max_sample_points
        super().__init__(num_classes, cfg.ddpm.time_dim, True, dropout=cfg.ddpm.dropout, input_dim=cfg.ddpm.input_dim,
                         extra_feature_channels=cfg.shapelatent.latent_dim, time_emb_scales=cfg.sde.embedding_scale,
                         verbose=True, condition_input=False, cfg=cfg, sa_blocks=self.sa_blocks,
                         fp_blocks=self.fp_blocks, clip_forge_enable=self.clip_forge_enable,
                         clip_forge_dim=cfg.clipforge.feat_dim)
        self.mixed_prediction = cfg.sde.mixed_prediction
        if self.mixed_prediction:
            raise NotImplementedError("lion_b200: sde.mixed_prediction is false in every shipped prior config")
        self.mixing_logit = None
        self.is_active = None

    @torch.no_grad()
    def forward(self, x, t, *args, **kwargs):
        """x: [B, N*D] or [B, N*D, 1, 1]; returns the same shape (localprior.py:72-83).
        The reference's view/permute/contiguous pairs vanish: [B,N,D] is the kernels' layout."""
        assert 'condition_input' in kwargs, 'require condition_input'
        input_shape = x.shape
        xin = x.detach().to(torch.float32).contiguous().view(-1, self.num_points, self.num_classes)
        style = kwargs['condition_input']
        style = style.reshape(xin.shape[0], -1)
        out = self.forward_point_major(xin, t=t, style=style, clip_feat=kwargs.get('clip_feat', None))
        return out.view(input_shape)
