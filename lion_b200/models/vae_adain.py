"""Sampling-side subset of the reference's VAE wrapper (models/vae_adain.py): `sample`
(:301-333), `latent_shape` (:335-339), `compose_eps` / `decompose_eps` (:97-103) and
`global2style` (:120-127).  The encoders (`encode`, `get_loss`, ...) are training /
reconstruction code and are out of scope; their parameters are therefore absent and a full
reference checkpoint is loaded with `load_state_dict(..., strict=False)` (only `decoder.*`
is consumed)."""
import torch
import torch.nn as nn

from .latent_points_ada import LatentPointDecPVC


class Model(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.input_dim = args.ddpm.input_dim
        self.latent_dim = args.shapelatent.latent_dim
        self.num_points = args.data.tr_max_sample_points
        assert len(args.latent_pts.style_mlp) == 0, "lion_b200: latent_pts.style_mlp is '' in every shipped config"
        self.style_mlp = None
        assert 'LatentPointDecPVC' in args.shapelatent.decoder_type
        self.decoder = LatentPointDecPVC(context_dim=self.latent_dim, point_dim=args.ddpm.input_dim, args=args)

    def compose_eps(self, all_eps):
        return torch.cat(all_eps, dim=1)

    def decompose_eps(self, all_eps):
        sd = self.args.latent_pts.style_dim
        return [all_eps[:, :sd], all_eps[:, sd:]]

    def global2style(self, style):
        return style            # style_mlp is None

    @torch.no_grad()
    def sample(self, num_samples=10, temp=None, decomposed_eps=[], enable_autocast=False, device_str='cuda', cls_emb=None):
        latent_shape = (num_samples, self.num_points * (self.latent_dim + self.input_dim))
        style_latent_shape = (num_samples, self.args.latent_pts.style_dim)
        if len(decomposed_eps) == 0:
            z_local = torch.zeros(*latent_shape).to(torch.device(device_str)).normal_()
            z_global = torch.zeros(*style_latent_shape).to(torch.device(device_str)).normal_()
        else:
            z_global = decomposed_eps[0].reshape(style_latent_shape)
            z_local = decomposed_eps[1].reshape(latent_shape)
        return self.decoder(None, beta=None, context=z_local, style=z_global)

    def latent_shape(self):
        return [[self.args.latent_pts.style_dim, 1, 1], [self.num_points * (self.latent_dim + self.input_dim), 1, 1]]
