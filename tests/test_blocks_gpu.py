"""Module-level parity: each block of pvcnn2_ada through the C ABI against the CPU oracle,
on the reference's channel-major layouts, with key-seeded synthetic weights."""
import pytest
import torch

from oracle import net as ON
from tests.synth import synth_state_dict
from tests.util import assert_close, gen

pytestmark = pytest.mark.gpu

TOL = 2e-3   # convolutions run in TF32 on the tensor cores (the reference's cuDNN path does too)


def _cfg():
    from lion_b200.config import default_prior_cfg
    return default_prior_cfg()


def _load(mod, seed):
    sd = synth_state_dict({k: list(v.shape) for k, v in mod.state_dict().items()}, seed)
    mod.load_state_dict(sd)
    return mod.cuda().eval(), sd


@pytest.mark.parametrize("cin,outs,R", [(35, [32, 64], 4096), (320, [128, 128], 300), (193, [128, 128, 64], 2048), (64, [128], 777)])
def test_shared_mlp(cin, outs, R):
    from lion_b200.models.pvcnn2_ada import SharedMLP
    m, sd = _load(SharedMLP(cin, outs, dim=1, cfg=_cfg()), 21)
    x, style = gen(1, 2, cin, R), gen(2, 2, 128)
    out = m(x.cuda(), style.cuda())
    assert_close(out, ON.shared_mlp(sd, "", x, style, len(outs)), TOL, "SharedMLP")


@pytest.mark.parametrize("C,heads,N", [(64, 4, 1024), (128, 8, 16), (64, 4, 100)])
def test_linear_attention(C, heads, N):
    from lion_b200.models.pvcnn2_ada import LinearAttention
    m, sd = _load(LinearAttention(C, heads), 22)
    x = gen(3, 2, C, N)
    assert_close(m(x.cuda()), ON.linear_attention(sd, "", x, heads), TOL, "LinearAttention")


@pytest.mark.parametrize("cin,cout,r,N,attn", [(4, 32, 32, 2048, False), (128, 64, 16, 1024, True), (192, 128, 8, 256, False),
                                               (128, 128, 8, 64, False), (64, 64, 32, 2048, False), (32, 32, 32, 2048, False),
                                               (64, 64, 32, 700, False)])
def test_pvconv(cin, cout, r, N, attn):
    from lion_b200.models.pvcnn2_ada import PVConv
    m, sd = _load(PVConv(cin, cout, 3, r, with_se=True, attention=attn, cfg=_cfg()), 23)
    B = 2
    feats, coords, style = gen(4, B, cin, N), gen(5, B, 3, N, scale=0.4), gen(6, B, 128)
    out, *_ = m((feats.cuda(), coords.cuda(), None, style.cuda()))
    blk = dict(kind="pvconv", cin=cin, cout=cout, r=r, attn=attn)
    assert_close(out, ON.pvconv(sd, "", blk, feats, coords, style), TOL, "PVConv")


@pytest.mark.parametrize("cfeat,m,radius,outs,N", [(32, 1024, 0.1, [32, 64], 2048), (64, 256, 0.2, [64, 128], 1024),
                                                   (192, 16, 0.8, [128, 128, 128], 64)])
def test_sa_module(cfeat, m, radius, outs, N):
    from lion_b200.models.pvcnn2_ada import PointNetSAModule
    mod, sd = _load(PointNetSAModule(m, radius, 32, cfeat, outs, cfg=_cfg()), 24)
    B = 2
    feats, coords, style = gen(7, B, cfeat, N), gen(8, B, 3, N, scale=0.3), gen(9, B, 128)
    out, centers, _, _ = mod((feats.cuda(), coords.cuda(), None, style.cuda()))
    blk = dict(kind="sa", m=m, radius=radius, k=32, cin=cfeat + 3, mlp=outs)
    o_feat, o_centers, _ = ON.sa_module(sd, "", blk, feats, coords, None, style)
    assert torch.equal(centers.cpu(), o_centers)
    assert_close(out, o_feat, TOL, "SA module")


def test_sa_module_is_bit_reproducible():
    """The fused level-0 SA kernel (csrc/sa_fused.cu) and the pooled convolution epilogue reduce their GroupNorm
    statistics in a fixed order: the same input gives the same bits, eagerly and run after run (the sampling loop relies
    on it: graph replay == eager loop)."""
    from lion_b200.models.pvcnn2_ada import PointNetSAModule
    for cfeat, m, radius, outs, N in [(32, 1024, 0.1, [32, 64], 2048), (64, 256, 0.2, [64, 128], 1024)]:
        mod, sd = _load(PointNetSAModule(m, radius, 32, cfeat, outs, cfg=_cfg()), 24)
        B = 4
        feats, coords, style = gen(7, B, cfeat, N).cuda(), gen(8, B, 3, N, scale=0.3).cuda(), gen(9, B, 128).cuda()
        ref = mod((feats, coords, None, style))[0].clone()
        for _ in range(4):
            assert torch.equal(mod((feats, coords, None, style))[0], ref)


@pytest.mark.parametrize("cc,cp,outs,N,M", [(192, 128, [128, 128], 64, 16), (192, 1, [128, 128, 64], 2048, 1024), (128, 0, [64], 256, 64)])
def test_fp_module(cc, cp, outs, N, M):
    from lion_b200.models.pvcnn2_ada import PointNetFPModule
    mod, sd = _load(PointNetFPModule(cc + cp, outs, cfg=_cfg()), 25)
    B = 2
    pc = gen(10, B, 3, N, scale=0.3)
    cctr = pc[:, :, :M].contiguous()
    cf, style = gen(11, B, cc, M), gen(12, B, 128)
    pf = gen(13, B, cp, N) if cp else None
    if pf is not None:
        out, *_ = mod((pc.cuda(), cctr.cuda(), cf.cuda(), pf.cuda(), None, style.cuda()))
    else:
        out, *_ = mod((pc.cuda(), cctr.cuda(), cf.cuda(), None, style.cuda()))
    blk = dict(kind="fp", cin=cc + cp, mlp=outs)
    o, _ = ON.fp_module(sd, "", blk, pc, cctr, cf, pf, None, style)
    assert_close(out, o, TOL, "FP module")


@pytest.mark.parametrize("C,shape", [(64, (5, 7, 3)), (32, (300,)), (128, (40, 32))])
def test_adagn_standalone(C, shape):
    from lion_b200.models.adagn import AdaGN
    m, sd = _load(AdaGN(len(shape), _cfg(), C), 26)
    x, style = gen(14, 2, C, *shape), gen(15, 2, 128)
    assert_close(m(x.cuda(), style.cuda()), ON.adagn(sd, "", x, style), 1e-5, "AdaGN")


def test_se3d_and_swish_standalone():
    from lion_b200.models.pvcnn2_ada import SE3d, Swish
    m, sd = _load(SE3d(64), 27)
    x = gen(16, 2, 64, 8, 8, 8)
    se = x.mean(-1).mean(-1).mean(-1)
    ref = x * torch.sigmoid(torch.relu(se @ sd["fc.0.weight"].T) @ sd["fc.2.weight"].T)[:, :, None, None, None]
    assert_close(m(x.cuda()), ref, 1e-5, "SE3d")
    assert_close(Swish()(x.cuda()), x * torch.sigmoid(x), 1e-5, "Swish")


def _tf32_rna(x):
    """cvt.rna.tf32.f32 on the CPU: round the magnitude to 10 mantissa bits, ties away from zero."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


@pytest.mark.parametrize("cin,cout,r,B", [
    (4, 32, 32, 1),      # sa0 first conv (K padded to one chunk)
    (32, 32, 32, 2),     # N = 32 tiles
    (64, 64, 32, 1),     # the dominant shape of the step (fp3)
    (128, 64, 16, 2),    # N = 64, four 32-channel chunks
    (192, 128, 8, 3),    # N = 128, 16-channel chunks, ragged row-tile groups
    (128, 128, 16, 1),
    (64, 256, 8, 2),     # two N tiles
    (20, 40, 5, 2),      # cout not a multiple of 32: SIMT kernel
    (3, 8, 4, 1),
])
def test_conv3d_standalone(cin, cout, r, B):
    """a7: the 3x3x3 convolution alone (tcgen05 kernel, lion_conv3d_gn_fwd) against torch's fp32
    conv3d on the CPU -- TF32 tolerance -- and against the same convolution evaluated on
    TF32-rounded operands, which isolates the kernel's indexing / accumulation (fp32 order only)."""
    from lion_b200.models.pvcnn2_ada import Conv3d
    m = Conv3d(cin, cout, 3, stride=1, padding=1)
    w, b = gen(61, cout, cin, 3, 3, 3, scale=(27 * cin) ** -0.5), gen(62, cout, scale=0.1)
    m.load_state_dict({"weight": w, "bias": b})
    m = m.cuda().eval()
    x = gen(63, B, cin, r, r, r)
    out, ssum, ssq = m(x.cuda(), return_gn_stats=True)
    ref = torch.nn.functional.conv3d(x, w, b, padding=1)
    assert_close(out, ref, TOL, "conv3d vs fp32")
    # (the SIMT kernel that serves output widths that are not a multiple of 32 keeps fp32 weights)
    w_t = _tf32_rna(w) if cout % 32 == 0 else w
    ref_t = torch.nn.functional.conv3d(_tf32_rna(x), w_t, b, padding=1)
    assert_close(out, ref_t, 2e-5, "conv3d vs TF32-operand emulation")
    o64 = out.double().cpu().view(B, cout, -1)
    assert_close(ssum, o64.sum(-1), 1e-5, "fused GroupNorm sum")
    assert_close(ssq, (o64 * o64).sum(-1), 1e-5, "fused GroupNorm sum of squares")
    assert torch.equal(m(x.cuda()), out), "conv3d is not bit-reproducible"


# the eight (r, C_in, C_out) classes of one PVCNN2Prior step at the BENCHMARKED batch
B32_CONV_SHAPES = [(4, 32, 32), (32, 32, 32), (64, 64, 32), (128, 64, 16), (64, 64, 16), (128, 128, 16), (192, 128, 8), (128, 128, 8)]


@pytest.mark.parametrize("cin,cout,r", B32_CONV_SHAPES)
def test_conv3d_b32_vs_cudnn(cin, cout, r):
    """a7 at B = 32 (BASELINE configs[1]): the tcgen05 convolution against torch's conv3d evaluated ON THE GPU
    (the CPU would need minutes): (i) cuDNN with TF32 allowed = the reference's own path under default torch
    flags (models/pvcnn2_ada.py:211-222) -- both sides round operands to TF32, tolerance 2e-3; (ii) full-fp32
    cuDNN on TF32(rna)-rounded operands, which isolates indexing / accumulation order: 2e-5; (iii) the fused
    GroupNorm statistics against float64 sums of the kernel's own output."""
    from lion_b200.models.pvcnn2_ada import Conv3d
    B = 32
    m = Conv3d(cin, cout, 3, stride=1, padding=1)
    w, b = gen(61, cout, cin, 3, 3, 3, scale=(27 * cin) ** -0.5), gen(62, cout, scale=0.1)
    m.load_state_dict({"weight": w, "bias": b})
    m = m.cuda().eval()
    x = gen(63, B, cin, r, r, r).cuda()
    out, ssum, ssq = m(x, return_gn_stats=True)
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        ref_tf32 = torch.nn.functional.conv3d(x, w.cuda(), b.cuda(), padding=1)
        torch.backends.cudnn.allow_tf32 = False
        ref_t = torch.nn.functional.conv3d(_tf32_rna(x.cpu()).cuda(), _tf32_rna(w).cuda(), b.cuda(), padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert_close(out, ref_tf32, TOL, "conv3d B=32 vs cuDNN TF32")
    assert_close(out, ref_t, 2e-5, "conv3d B=32 vs fp32 cuDNN on TF32-rounded operands")
    o64 = out.double().view(B, cout, -1)
    assert_close(ssum, o64.sum(-1), 1e-5, "fused GroupNorm sum, B=32")
    assert_close(ssq, (o64 * o64).sum(-1), 1e-5, "fused GroupNorm sum of squares, B=32")
