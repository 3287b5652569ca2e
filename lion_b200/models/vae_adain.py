"""The hierarchical VAE wrapper -- host-side mirror of the inference half of the reference's models/vae_adain.py:
`sample` (:301-333), `latent_shape` (:335-339), `compose_eps` / `decompose_eps` (:97-103), `global2style` (:120-127),
and the encoder path `encode` (:55-95), `encode_global` / `encode_local` (:105-135), `recont` (:137-207) used by the
reconstruction / interpolation apps.  `get_loss` (training) is out of scope.

Module names (`style_encoder`, `encoder`, `decoder`) and their state_dict keys are the reference's, so a released
`vae_state_dict` loads with strict=True.  Every network forward is a C-ABI call into liblion_b200.so; the reshapes,
the reparameterisation `z = mu + exp(log_sigma) * rho` and the log-density are a handful of torch tensor ops exactly
as in the reference (models/distributions.py)."""
import torch
import torch.nn as nn

from .distributions import Normal
from .latent_points_ada import LatentPointDecPVC, PointTransPVC
from .shapelatent_modules import PointNetPlusEncoder


class Model(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_total_iter = 0
        self.args = args
        self.input_dim = args.ddpm.input_dim
        self.latent_dim = args.shapelatent.latent_dim
        self.kl_weight = args.shapelatent.kl_weight
        self.num_points = args.data.tr_max_sample_points
        assert 'PointNetPlusEncoder' in args.latent_pts.style_encoder, "lion_b200: latent_pts.style_encoder must be PointNetPlusEncoder"
        self.style_encoder = PointNetPlusEncoder(zdim=args.latent_pts.style_dim, input_dim=self.input_dim, args=args)
        assert len(args.latent_pts.style_mlp) == 0, "lion_b200: latent_pts.style_mlp is '' in every shipped config"
        self.style_mlp = None
        assert 'PointTransPVC' in args.shapelatent.encoder_type
        self.encoder = PointTransPVC(zdim=self.latent_dim, input_dim=self.input_dim, args=args)
        assert 'LatentPointDecPVC' in args.shapelatent.decoder_type
        self.decoder = LatentPointDecPVC(context_dim=self.latent_dim, point_dim=args.ddpm.input_dim, args=args)
        assert not getattr(args.data, 'cond_on_cat', 0), "lion_b200: class-conditional VAEs (data.cond_on_cat) are not provided"

    # ---- encoder path (reconstruction / interpolation apps) ------------------------------------------------
    def _encode(self, x):
        assert x.shape[2] == self.input_dim, f'expect input in [B,Npoint,PointDim={self.input_dim}], get: {x.shape}'
        latent_list, all_eps, all_log_q = [], [], []
        z = self.style_encoder(x)
        z_mu, z_sigma = z['mu_1d'], z['sigma_1d']          # log sigma
        dist = Normal(mu=z_mu, log_sigma=z_sigma)
        z_global = dist.sample()[0]
        all_eps.append(z_global)
        all_log_q.append(dist.log_p(z_global))
        latent_list.append([z_global, z_mu, z_sigma])
        style = self.global2style(z_global)
        z = self.encoder([x, style])
        z_mu, z_sigma = z['mu_1d'], z['sigma_1d'] - self.args.shapelatent.log_sigma_offset
        dist = Normal(mu=z_mu, log_sigma=z_sigma)
        z_local = dist.sample()[0]
        all_eps.append(z_local)
        all_log_q.append(dist.log_p(z_local))
        latent_list.append([z_local, z_mu, z_sigma])
        return all_eps, all_log_q, latent_list, style

    @torch.no_grad()
    def encode(self, x, class_label=None):
        all_eps, all_log_q, latent_list, _ = self._encode(x)
        return self.compose_eps(all_eps), all_log_q, latent_list

    @torch.no_grad()
    def encode_global(self, x, class_label=None):
        z = self.style_encoder(x)
        return Normal(mu=z['mu_1d'], log_sigma=z['sigma_1d'])

    @torch.no_grad()
    def encode_local(self, x, style):
        z = self.encoder([x, style])
        return Normal(mu=z['mu_1d'], log_sigma=z['sigma_1d'] - self.args.shapelatent.log_sigma_offset)

    @torch.no_grad()
    def recont(self, x, target=None, class_label=None, cls_emb=None):
        batch_size = x.shape[0]
        x_0_target = x if target is None else target
        all_eps, all_log_q, latent_list, style = self._encode(x)
        z_local = latent_list[1][0]
        x_0_pred = self.decoder(None, beta=None, context=z_local, style=style)
        make_4d = lambda v: v.unsqueeze(-1).unsqueeze(-1) if len(v.shape) == 2 else v.unsqueeze(-1)
        output = {'all_eps': [make_4d(e) for e in all_eps], 'all_log_q': [make_4d(e) for e in all_log_q],
                  'latent_list': latent_list, 'x_0_pred': x_0_pred, 'x_0_target': x_0_target,
                  'x_t': torch.zeros_like(x_0_target), 't': torch.zeros(batch_size), 'x_0': x_0_target}
        output['hist/global_var'] = latent_list[0][2].exp()
        latent_pts = z_local.view(batch_size, -1, self.latent_dim + self.input_dim)[:, :, :self.input_dim].contiguous().clone()
        output['vis/latent_pts'] = latent_pts.detach().cpu().view(batch_size, -1, self.input_dim)
        output['final_pred'] = output['x_0_pred']
        return output

    # ---- sampling path ---------------------------------------------------------------------------------------
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
