"""CPU restatement of the VAE ENCODER path (reconstruction / interpolation apps).

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.  No product code exists for
this row yet; the oracle and its golden vectors are laid down first (round-2 groundwork).

  style_encoder_forward   models/shapelatent_modules.py:13-52 (PointNetPlusEncoder) on the NON-Ada blocks of
                          models/pvcnn2.py: PVConv :170-247 (Conv3d, GroupNorm(8), Swish, Conv3d, GroupNorm(8),
                          SE3d; + SharedMLP point branch; LinearAttention), PointNetSAModule :288-351,
                          table builder create_pointnet2_sa_components :440-509 (same quirk as the Ada builder:
                          levels > 0 keep only their first PVConv)
  point_encoder_forward   models/latent_points_ada.py:175-220 (PointTransPVC): the Ada U-Net of oracle/net.py with
                          embed_dim = 0, no extra features, num_classes = 2*zdim + 2*input_dim
  encode                  models/vae_adain.py:137-175 (the encoder half of `recont`): z = mu + exp(log_sigma) * eps

Parity status: pinned by tests/golden/encoder_fwd.npz (the reference's own modules run on CPU,
tests/golden/make_golden_encoder.py).
"""
import torch
import torch.nn.functional as TF

from . import net as ON
from . import point_ops as P

STYLE_SA_BLOCKS = [  # models/shapelatent_modules.py:14-17
    ((32, 2, 32), (1024, 0.1, 32, (32, 32))),
    ((32, 1, 16), (256, 0.2, 32, (32, 64))),
]


def _gn(sd, p, x):
    return TF.group_norm(x, 8, sd[p + "weight"], sd[p + "bias"], eps=1e-5)


def shared_mlp_plain(sd, p, x, n):
    """pvcnn2.py SharedMLP :117-138: n x (1x1 conv, GroupNorm(8), Swish); keys layers.{3i}, layers.{3i+1}."""
    for i in range(n):
        w = sd[p + "layers.%d.weight" % (3 * i)]
        x = torch.einsum("oc,bc...->bo...", w.reshape(w.shape[0], -1), x) + sd[p + "layers.%d.bias" % (3 * i)].reshape(
            (1, -1) + (1,) * (x.dim() - 2))
        x = ON.swish(_gn(sd, p + "layers.%d." % (3 * i + 1), x))
    return x


def pvconv_plain(sd, p, blk, features, coords):
    """pvcnn2.py PVConv.forward :207-247 with with_se=True."""
    r = blk["r"]
    norm_coords, vox = P.voxel_coords(coords, r)
    g, _, _ = P.avg_voxelize(features, vox, r)
    g = TF.conv3d(g, sd[p + "voxel_layers.0.weight"], sd[p + "voxel_layers.0.bias"], padding=1)
    g = ON.swish(_gn(sd, p + "voxel_layers.1.", g))
    g = TF.conv3d(g, sd[p + "voxel_layers.4.weight"], sd[p + "voxel_layers.4.bias"], padding=1)
    g = _gn(sd, p + "voxel_layers.5.", g)
    se = g.mean(-1).mean(-1).mean(-1)
    se = torch.sigmoid(TF.linear(torch.relu(TF.linear(se, sd[p + "voxel_layers.6.fc.0.weight"])), sd[p + "voxel_layers.6.fc.2.weight"]))
    g = g * se[:, :, None, None, None]
    out = P.trilinear_devoxelize(g, norm_coords, r)
    out = out + shared_mlp_plain(sd, p + "point_features.", features, 1)
    if blk["attn"]:
        out = ON.linear_attention(sd, p + "attn.", out, 4)
    return out


def sa_module_plain(sd, p, blk, features, coords):
    """pvcnn2.py PointNetSAModule.forward :323-351 (one grouper) + BallQuery.forward :94-112."""
    centers = P.furthest_point_sample(coords, blk["m"])
    idx = P.ball_query(centers, coords, blk["radius"], blk["k"])
    ncoords = P.grouping(coords, idx) - centers.unsqueeze(-1)
    nfeat = torch.cat([ncoords, P.grouping(features, idx)], dim=1)
    out = shared_mlp_plain(sd, p + "mlps.0.", nfeat, len(blk["mlp"])).max(dim=-1).values
    return out, centers


def style_plan(sa_blocks=STYLE_SA_BLOCKS, input_dim=3, use_att=True):
    """create_pointnet2_sa_components(sa_blocks, 0, input_dim, embed_dim=0, use_att=True, with_se=True)."""
    in_ch = input_dim
    levels = []
    for c, (conv_cfg, sa_cfg) in enumerate(sa_blocks):
        blocks, k = [], 0
        if conv_cfg is not None:
            oc, nblk, res = conv_cfg
            for pidx in range(nblk):
                att = ((c + 1) % 2 == 0 and use_att and pidx == 0)
                if c == 0 or k == 0:
                    blocks.append(dict(kind="pvconv", cin=in_ch, cout=oc, r=res, attn=att))
                in_ch = oc
                k += 1
        m, radius, nn, mlp = sa_cfg
        blocks.append(dict(kind="sa", m=m, radius=radius, k=nn, cin=in_ch + 3, mlp=list(mlp)))
        in_ch = mlp[-1]
        levels.append(blocks)
    return levels, in_ch


def style_encoder_forward(sd, x, zdim=128, prefix=""):
    """x [B,N,3] -> (mu_1d [B,zdim], sigma_1d [B,zdim] = log sigma)."""
    x = torch.as_tensor(x, dtype=torch.float32).transpose(1, 2).contiguous()
    xyz, feat = x, x
    levels, _ = style_plan()
    for li, blocks in enumerate(levels):
        for j, blk in enumerate(blocks):
            p = prefix + ("layers.%d." % li if len(blocks) == 1 else "layers.%d.%d." % (li, j))
            if blk["kind"] == "pvconv":
                feat = pvconv_plain(sd, p, blk, feat, xyz)
            else:
                feat, xyz = sa_module_plain(sd, p, blk, feat, xyz)
    feat = feat.max(-1)[0]
    out = TF.linear(feat, sd[prefix + "mlp.weight"], sd[prefix + "mlp.bias"])
    return out[:, :zdim], out[:, zdim:]


def point_encoder_spec(zdim=1, input_dim=3):
    return ON.UnetSpec(2 * zdim + 2 * input_dim, 0, 0, ON.DEC_SA_BLOCKS, ON.FP_BLOCKS, input_dim=input_dim)


def point_encoder_forward(sd, x, style, zdim=1, input_dim=3, skip_weight=0.01, pts_sigma_offset=0.0, prefix="layers."):
    """PointTransPVC.forward: x [B,N,3], style [B,S] -> (mu_1d [B, N*(3+zdim)], sigma_1d same shape)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    B = x.shape[0]
    out = ON.unet_forward(sd, point_encoder_spec(zdim, input_dim), x.permute(0, 2, 1).contiguous(), style=style, prefix=prefix)
    out = out.permute(0, 2, 1).contiguous()
    pt_mu = skip_weight * out[:, :, :input_dim] + x
    pt_sigma = out[:, :, input_dim:2 * input_dim] - pts_sigma_offset
    ft_mu = out[:, :, 2 * input_dim:-zdim]
    ft_sigma = out[:, :, -zdim:]
    mu = torch.cat([pt_mu, ft_mu], dim=2).reshape(B, -1)
    sigma = torch.cat([pt_sigma, ft_sigma], dim=2).reshape(B, -1)
    return mu, sigma


def encode(sd_style, sd_enc, x, eps_global, eps_local, log_sigma_offset=6.0):
    """vae_adain.Model.recont up to the decoder: (z_global, z_local); eps_* are the N(0,1) draws."""
    mu_g, ls_g = style_encoder_forward(sd_style, x)
    z_g = mu_g + torch.exp(ls_g) * eps_global
    mu_l, ls_l = point_encoder_forward(sd_enc, x, z_g)
    z_l = mu_l + torch.exp(ls_l - log_sigma_offset) * eps_local
    return z_g, z_l
