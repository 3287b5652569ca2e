"""CPU restatement (plain torch, fp32) of the networks on LION's sampling hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/point_ops.py for the import rules.

The networks are restated *functionally*: a forward is a function of a flat state_dict that
uses the reference's parameter names, an architecture table and the
inputs.  File:line citations point into /root/reference.

  unet_forward        models/latent_points_ada.py:117-173   (PVCNN2Unet.forward)
  prior_forward       models/latent_points_ada_localprior.py:72-83
  decoder_forward     models/latent_points_ada.py:255-272    (LatentPointDecPVC.forward)
  global_prior_forward models/score_sde/resnet.py:195-218    (Prior.forward)

Parity status: pinned against the reference's own Python modules imported in the build
container (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py).  The reference has no golden vectors of its own for this path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as TF

from . import point_ops as P


def set_point_ops(ops):
    """Swap the point-operator backend (default: oracle/point_ops.py on the CPU).  oracle/ref_cuda_ops.py plugs
    in the reference's own CUDA kernels for the GPU-side baseline of bench.py."""
    global P
    P = ops

# ----------------------------------------------------------------------------------------
# architecture tables
# ----------------------------------------------------------------------------------------
PRIOR_SA_BLOCKS = [  # models/latent_points_ada_localprior.py:17-22
    ((32, 2, 32), (1024, 0.1, 32, (32, 64))),
    ((64, 3, 16), (256, 0.2, 32, (64, 128))),
    ((128, 3, 8), (64, 0.4, 32, (128, 128))),
    (None, (16, 0.8, 32, (128, 128, 128))),
]
DEC_SA_BLOCKS = [  # models/latent_points_ada.py:225-230
    ((32, 2, 32), (1024, 0.1, 32, (32, 64))),
    ((64, 3, 16), (256, 0.2, 32, (64, 128))),
    ((128, 3, 8), (64, 0.4, 32, (128, 256))),
    (None, (16, 0.8, 32, (128, 128, 128))),
]
FP_BLOCKS = [  # identical in both (localprior.py:23-28, latent_points_ada.py:231-236)
    ((128, 128), (128, 3, 8)),
    ((128, 128), (128, 3, 8)),
    ((128, 128), (128, 2, 16)),
    ((128, 128, 64), (64, 2, 32)),
]


class UnetSpec:
    """What PVCNN2Unet.__init__ (latent_points_ada.py:23-99) is told."""

    def __init__(self, num_classes, embed_dim, extra_feature_channels, sa_blocks, fp_blocks,
                 input_dim=3, use_att=True, clip=False, time_emb_scales=1.0):
        self.num_classes = num_classes
        self.embed_dim = embed_dim
        self.extra_feature_channels = extra_feature_channels
        self.sa_blocks = sa_blocks
        self.fp_blocks = fp_blocks
        self.input_dim = input_dim
        self.use_att = use_att
        self.clip = clip
        self.time_emb_scales = time_emb_scales


def prior_spec(latent_dim=1, input_dim=3, time_dim=64, clip=False):
    """PVCNN2Prior (localprior.py:31-57): num_classes = latent_dim + input_dim."""
    return UnetSpec(latent_dim + input_dim, time_dim, latent_dim, PRIOR_SA_BLOCKS, FP_BLOCKS,
                    input_dim=input_dim, clip=clip)


def decoder_spec(latent_dim=1, input_dim=3):
    """LatentPointDecPVC (latent_points_ada.py:238-253): embed_dim=0, num_classes=point_dim."""
    return UnetSpec(input_dim, 0, latent_dim, DEC_SA_BLOCKS, FP_BLOCKS, input_dim=input_dim)


def build_plan(spec):
    """Layer list of the U-Net, restating create_pointnet2_sa_components
    (models/pvcnn2_ada.py:448-517) and create_pointnet2_fp_modules (:520-567), including the
    two table-builder quirks: levels > 0 keep only their first PVConv,
    and no FP PVConv ever gets attention."""
    E = spec.embed_dim
    in_ch = spec.extra_feature_channels + spec.input_dim
    sa, sa_in = [], []
    for c, (conv_cfg, sa_cfg) in enumerate(spec.sa_blocks):
        blocks = []
        k = 0
        sa_in.append(in_ch)
        if conv_cfg is not None:
            oc, nblk, res = conv_cfg
            for p in range(nblk):
                att = ((c + 1) % 2 == 0 and spec.use_att and p == 0)
                if c == 0:
                    blocks.append(dict(kind="pvconv", cin=in_ch, cout=oc, r=res, attn=att))
                elif k == 0:
                    blocks.append(dict(kind="pvconv", cin=in_ch + E, cout=oc, r=res, attn=att))
                in_ch = oc
                k += 1
        extra = in_ch
        m, radius, nn, mlp = sa_cfg
        cin = extra + (E if k == 0 else 0) + 3
        blocks.append(dict(kind="sa", m=m, radius=radius, k=nn, cin=cin, mlp=list(mlp)))
        in_ch = mlp[-1]
        sa.append(blocks)
    ch_sa = in_ch
    sa_in[0] = spec.extra_feature_channels + spec.input_dim - 3      # latent_points_ada.py:81
    fp = []
    for i, (fp_cfg, conv_cfg) in enumerate(spec.fp_blocks):
        blocks = [dict(kind="fp", cin=in_ch + sa_in[-1 - i] + E, mlp=list(fp_cfg))]
        in_ch = fp_cfg[-1]
        if conv_cfg is not None:
            oc, nblk, res = conv_cfg
            for _ in range(nblk):
                blocks.append(dict(kind="pvconv", cin=in_ch, cout=oc, r=res, attn=False))
                in_ch = oc
        fp.append(blocks)
    return dict(sa=sa, fp=fp, ch_sa=ch_sa, ch_fp=in_ch)


def _prefix(kind, level, j, nblocks):
    """State-dict prefix: a level with one block is the block itself, otherwise an
    nn.Sequential (pvcnn2_ada.py:512-515, :561-564)."""
    base = "%s_layers.%d" % (kind, level)
    return base + "." if nblocks == 1 else "%s.%d." % (base, j)


# ----------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------
def swish(x):
    return x * torch.sigmoid(x)


def adagn(sd, p, x, style):
    """AdaGN.forward (models/adagn.py:45-65): GroupNorm(8,C,eps=1e-5,affine) then
    *factor + bias with [factor|bias] = Linear(style)."""
    C = x.shape[1]
    fb = TF.linear(style, sd[p + "emd.weight"], sd[p + "emd.bias"])
    shape = [x.shape[0], C] + [1] * (x.dim() - 2)
    y = TF.group_norm(x, 8, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return y * fb[:, :C].reshape(shape) + fb[:, C:].reshape(shape)


def shared_mlp(sd, p, x, style, n):
    """SharedMLP.forward (pvcnn2_ada.py:140-164): n x (1x1 conv -> AdaGN -> Swish); layers are
    numbered conv 3i, AdaGN 3i+1, Swish 3i+2 (:131-137).  Works for [B,C,N] and [B,C,M,U]."""
    for i in range(n):
        w = sd[p + "layers.%d.weight" % (3 * i)]
        w2 = w.reshape(w.shape[0], w.shape[1])
        x = torch.einsum("oc,bc...->bo...", w2, x) + sd[p + "layers.%d.bias" % (3 * i)].reshape(
            [1, -1] + [1] * (x.dim() - 2))
        x = swish(adagn(sd, p + "layers.%d." % (3 * i + 1), x, style))
    return x


def linear_attention(sd, p, x, heads):
    """LinearAttention.forward (pvcnn2_ada.py:54-71): softmax over points on k only, q
    unscaled, no residual.  Channel split of to_qkv is (qkv, heads, c) (:64)."""
    B, C, N = x.shape
    wq = sd[p + "to_qkv.weight"].reshape(-1, C)
    qkv = torch.einsum("oc,bcn->bon", wq, x).view(B, 3, heads, -1, N)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, -1, N)
    wo = sd[p + "to_out.weight"]
    return torch.einsum("oc,bcn->bon", wo.reshape(wo.shape[0], -1), out) + sd[p + "to_out.bias"][None, :, None]


def pvconv(sd, p, blk, features, coords, style):
    """PVConv.forward (pvcnn2_ada.py:235-280)."""
    r = blk["r"]
    norm_coords, vox = P.voxel_coords(coords, r)                     # Voxelization (:173-188)
    g, _, _ = P.avg_voxelize(features, vox, r)
    g = TF.conv3d(g, sd[p + "voxel_layers.0.weight"], sd[p + "voxel_layers.0.bias"], padding=1)
    g = swish(adagn(sd, p + "voxel_layers.1.", g, style))            # .2 Swish, .3 Dropout (eval)
    g = TF.conv3d(g, sd[p + "voxel_layers.4.weight"], sd[p + "voxel_layers.4.bias"], padding=1)
    g = adagn(sd, p + "voxel_layers.5.", g, style)
    # SE3d (:27-41): reduction 8, no bias
    se = g.mean(-1).mean(-1).mean(-1)
    se = torch.sigmoid(TF.linear(torch.relu(TF.linear(se, sd[p + "voxel_layers.6.fc.0.weight"])),
                                 sd[p + "voxel_layers.6.fc.2.weight"]))
    g = g * se[:, :, None, None, None]
    out = P.trilinear_devoxelize(g, norm_coords, r)
    out = out + shared_mlp(sd, p + "point_features.", features, style, 1)
    if blk["attn"]:
        out = linear_attention(sd, p + "attn.", out, 4)
    return out


def sa_module(sd, p, blk, features, coords, temb, style):
    """PointNetSAModule.forward (pvcnn2_ada.py:354-382) + BallQuery.forward (:98-114)."""
    centers = P.furthest_point_sample(coords, blk["m"])
    S = centers.shape[-1]
    if temb is not None:
        temb = temb[:, :, :S]
    idx = P.ball_query(centers, coords, blk["radius"], blk["k"])
    ncoords = P.grouping(coords, idx) - centers.unsqueeze(-1)
    nfeat = torch.cat([ncoords, P.grouping(features, idx)], dim=1)
    out = shared_mlp(sd, p + "mlps.0.", nfeat, style, len(blk["mlp"])).max(dim=-1).values
    return out, centers, temb


def fp_module(sd, p, blk, points_coords, centers_coords, centers_features, points_features, temb, style):
    """PointNetFPModule.forward (pvcnn2_ada.py:393-411)."""
    x = P.nearest_neighbor_interpolate(points_coords, centers_coords, centers_features)
    if points_features is not None:
        x = torch.cat([x, points_features], dim=1)
    if temb is not None:
        temb = temb[:, :, 0:1].expand(-1, -1, points_coords.shape[-1])
    return shared_mlp(sd, p + "mlp.", x, style, len(blk["mlp"])), temb


def timestep_embedding(t, dim, scale=1.0):
    """PVCNN2Unet.get_timestep_embedding (latent_points_ada.py:101-115): frequencies in
    float64 numpy, cast to fp32, sin | cos."""
    t = torch.as_tensor(t, dtype=torch.float32) * scale
    half = dim // 2
    f = torch.from_numpy(np.exp(np.arange(0, half) * -(np.log(10000) / (half - 1)))).float().to(t.device)
    e = t[:, None] * f[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


# ----------------------------------------------------------------------------------------
# whole networks
# ----------------------------------------------------------------------------------------
def unet_forward(sd, spec, inputs, t=None, style=None, clip_feat=None, prefix="", plan=None, tap=None):
    """PVCNN2Unet.forward (latent_points_ada.py:117-173).  inputs [B, 3+extra, N]."""
    pre = prefix
    plan = plan or build_plan(spec)
    inputs = torch.as_tensor(inputs, dtype=torch.float32)
    B, _, N = inputs.shape
    coords = inputs[:, :spec.input_dim].contiguous()
    features = inputs
    temb = None
    if spec.embed_dim > 0 and t is not None:
        e = timestep_embedding(t, spec.embed_dim, spec.time_emb_scales)
        e = TF.linear(e, sd[pre + "embedf.0.weight"], sd[pre + "embedf.0.bias"])
        e = TF.leaky_relu(e, 0.1)
        e = TF.linear(e, sd[pre + "embedf.2.weight"], sd[pre + "embedf.2.bias"])
        temb = e[:, :, None].expand(-1, -1, N)
    if spec.clip:                                                     # :132-137
        cf = TF.linear(clip_feat, sd[pre + "clip_forge_mapping.weight"], sd[pre + "clip_forge_mapping.bias"])
        style = TF.linear(torch.cat([style, cf], dim=1), sd[pre + "style_clip.weight"], sd[pre + "style_clip.bias"])

    coords_list, feats_list = [], []
    for i, blocks in enumerate(plan["sa"]):
        feats_list.append(features)
        coords_list.append(coords)
        if i > 0 and temb is not None:
            features = torch.cat([features, temb], dim=1)
        for j, blk in enumerate(blocks):
            p = pre + _prefix("sa", i, j, len(blocks))
            if blk["kind"] == "pvconv":
                features = pvconv(sd, p, blk, features, coords, style)
            else:
                features, coords, temb = sa_module(sd, p, blk, features, coords, temb, style)
            if tap is not None:
                tap[p] = features
    feats_list[0] = inputs[:, 3:].contiguous()                        # :153
    if spec.use_att:
        features = linear_attention(sd, pre + "global_att.", features, 8)
        if tap is not None:
            tap[pre + "global_att."] = features
    for i, blocks in enumerate(plan["fp"]):
        for j, blk in enumerate(blocks):
            p = pre + _prefix("fp", i, j, len(blocks))
            if blk["kind"] == "fp":
                cf = torch.cat([features, temb], dim=1) if temb is not None else features
                features, temb = fp_module(sd, p, blk, coords_list[-1 - i], coords, cf,
                                           feats_list[-1 - i], temb, style)
                coords = coords_list[-1 - i]
            else:
                features = pvconv(sd, p, blk, features, coords, style)
            if tap is not None:
                tap[p] = features
    # classifier: SharedMLP(ch_fp -> 128), Dropout, Conv1d(128 -> num_classes) (:94-99, :168-172)
    features = shared_mlp(sd, pre + "classifier.0.", features, style, 1)
    w = sd[pre + "classifier.2.weight"]
    return torch.einsum("oc,bcn->bon", w.reshape(w.shape[0], -1), features) + sd[pre + "classifier.2.bias"][None, :, None]


def prior_forward(sd, spec, x, t, condition_input, clip_feat=None, num_points=2048, prefix=""):
    """PVCNN2Prior.forward (localprior.py:72-83): x [B, N*D(,1,1)] -> same shape."""
    shape = x.shape
    xin = torch.as_tensor(x, dtype=torch.float32).reshape(-1, num_points, spec.num_classes).permute(0, 2, 1).contiguous()
    style = torch.as_tensor(condition_input, dtype=torch.float32).reshape(xin.shape[0], -1)
    out = unet_forward(sd, spec, xin, t=t, style=style, clip_feat=clip_feat, prefix=prefix)
    return out.permute(0, 2, 1).contiguous().view(shape)


def decoder_forward(sd, spec, context, style, num_points=2048, skip_weight=0.01, prefix="layers."):
    """LatentPointDecPVC.forward (latent_points_ada.py:255-272): context [B, N*(3+D)] ->
    points [B,N,3] = out*skip_weight + xyz."""
    ctx = torch.as_tensor(context, dtype=torch.float32).view(-1, num_points, spec.input_dim + spec.extra_feature_channels)
    x = ctx[:, :, :spec.input_dim]
    out = unet_forward(sd, spec, ctx.permute(0, 2, 1).contiguous(), style=style, prefix=prefix)
    return out.permute(0, 2, 1).contiguous() * skip_weight + x


def positional_embedding(t, dim, scale=1.0):
    """models/utils.py:16-31 (fp32 frequencies, unlike the U-Net's)."""
    t = torch.as_tensor(t, dtype=torch.float32) * scale
    half = dim // 2
    f = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(t.device)
    e = t[:, None] * f[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


def _c11(sd, name, x, bias=True):
    w = sd[name + ".weight"]
    y = TF.linear(x, w.reshape(w.shape[0], -1))
    return y + sd[name + ".bias"] if bias else y


def global_prior_forward(sd, x, t, clip_feat=None, embedding_dim=128, embedding_scale=1.0, prefix=""):
    """Prior.forward with ResBlockSEDrop / ResBlockSEClip cells (resnet.py:195-218, :60-90,
    :29-56).  x [B,128(,1,1)] -> same shape.  temb_layer is two 1x1 convs with no activation
    (:181-184); clip feature is mapped by a Conv1d and concatenated to temb (:203-208)."""
    p = prefix
    shape = x.shape
    h = torch.as_tensor(x, dtype=torch.float32).reshape(shape[0], -1)
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    temb = positional_embedding(t, embedding_dim, embedding_scale)
    temb = _c11(sd, p + "temb_layer.1", _c11(sd, p + "temb_layer.0", temb))
    clip = None
    if clip_feat is not None:
        clip = _c11(sd, p + "clip_feat_mapping", torch.as_tensor(clip_feat, dtype=torch.float32))
        if temb.shape[0] == 1 and clip.shape[0] > 1:
            temb = temb.expand(clip.shape[0], -1)
    h = _c11(sd, p + "input_layer", h)
    k = 0
    while (p + "all_modules.%d.conv1.weight" % k) in sd:
        m = p + "all_modules.%d." % k
        o = h + temb
        if clip is not None:
            o = torch.cat([o, clip], dim=1)
        o = torch.relu(_c11(sd, m + "conv1", o))
        o = torch.relu(_c11(sd, m + "conv2", o))
        s = torch.sigmoid(_c11(sd, m + "SE.fc.2", torch.relu(_c11(sd, m + "SE.fc.0", o, bias=False)), bias=False))
        h = h + o * s
        k += 1
    return _c11(sd, p + "output_layer", h).view(shape)
