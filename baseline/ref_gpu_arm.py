"""The UNMODIFIED reference on the GPU, timed beside lion_b200 (north_star: "side-by-side with the reference's own
third_party/pvcnn CUDA path").  Runs the reference's own `generate_samples_vada_2prior`
(trainers/train_2prior.py:49-127) -> `DiffusionDiscretized.run_denoising_diffusion` (utils/diffusion_pvd.py:223-303)
-> `PriorSEDrop` / `PVCNN2Prior` / `Model.sample` from the copy under baseline/_ref/LION (baseline/make_ref_copy.py),
with its JIT-built `_pvcnn_backend` kernels and torch's cuDNN / cuBLAS, cudnn.benchmark on as utils/utils.py:472 sets
it, default TF32 flags.  Same weights (tests/synth.py, strict load_state_dict) and batch as bench.py's own arm.

    python baseline/ref_gpu_arm.py --batch 32 --steps 1000 [--clip]      -> one JSON line

Import-time stubs only for modules that are absent from this image and arithmetic-free on the sampling path (SURVEY.md
8c): comet_ml, matplotlib, clip, calmsize, diffusers, open3d, and third_party.PyTorchEMD (its CUDA source includes the
removed THC/THC.h; metrics/loss only).  None of lion_b200's code is imported here."""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref", "LION")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Anything(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Anything("call")


def install():
    assert os.path.isdir(os.path.join(REF, "models")), "baseline/_ref/LION is missing (python baseline/make_ref_copy.py)"
    os.environ["TORCH_EXTENSIONS_DIR"] = os.path.join(HERE, "_ref", "torch_extensions")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("quiet", "1")
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)                       # tests.synth (weights) only
    for n in ["comet_ml", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "clip", "calmsize", "diffusers", "mpl_toolkits",
              "mpl_toolkits.mplot3d", "open3d", "third_party.PyTorchEMD", "third_party.PyTorchEMD.emd",
              "third_party.PyTorchEMD.emd_nograd"]:
        if n not in sys.modules:
            m = _Anything(n)
            m.__path__ = []
            sys.modules[n] = m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=1000, help="DDPM steps per prior (the metric is defined at 1000)")
    ap.add_argument("--clip", action="store_true", help="BASELINE configs[3]: PriorSEClip + CLIP-conditioned AdaGN, clip_feat = randn")
    ap.add_argument("--passes", type=int, default=1)
    args = ap.parse_args()
    install()
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    from default_config import cfg as base
    cfg = base.clone()
    cfg.merge_from_file(os.path.join(REF, "config", "airplane_prior_cfg.yml"))
    ov = ["ddpm.num_steps", args.steps]
    if args.clip:
        ov += ["clipforge.enable", 1, "latent_pts.style_prior", "models.score_sde.resnet.PriorSEClip"]
    cfg.merge_from_list(ov)
    t_imp = time.perf_counter()
    from trainers.train_2prior import generate_samples_vada_2prior          # JIT-loads _pvcnn_backend and chamfer_3D
    from utils.diffusion_pvd import DiffusionDiscretized
    from models.latent_points_ada_localprior import PVCNN2Prior
    from models.score_sde.resnet import PriorSEDrop, PriorSEClip
    from models.vae_adain import Model
    t_imp = time.perf_counter() - t_imp
    from tests.synth import synth_state_dict
    torch.backends.cudnn.benchmark = True                                    # utils/utils.py:472 (common_init)
    shp = lambda m: {k: list(v.shape) for k, v in m.state_dict().items()}
    gp = (PriorSEClip if args.clip else PriorSEDrop)(cfg.sde, cfg.latent_pts.style_dim, cfg)
    gp.load_state_dict(synth_state_dict(shp(gp), 14), strict=True)
    lp = PVCNN2Prior(cfg.sde, 1, cfg)
    lp.load_state_dict(synth_state_dict(shp(lp), 11), strict=True)
    vae = Model(cfg)
    vae.decoder.load_state_dict(synth_state_dict(shp(vae.decoder), 13), strict=True)
    dae = torch.nn.ModuleList([gp, lp]).cuda().eval()
    vae = vae.cuda().eval()
    diffusion = DiffusionDiscretized(cfg.sde, None, cfg)
    shape = vae.latent_shape()
    B = args.batch
    clip_feat = None
    if args.clip:
        clip_feat = torch.randn(B, 512, generator=torch.Generator().manual_seed(7)).cuda()

    def one(seed):
        torch.manual_seed(seed)
        img, *_ = generate_samples_vada_2prior(shape, dae, diffusion, vae, B, False, clip_feat=clip_feat)
        return img

    # warm-up: a 3-step run of the same modules (cuDNN algorithm search, allocator)
    warm_cfg = cfg.clone()
    warm_cfg.merge_from_list(["ddpm.num_steps", 3])
    warm = DiffusionDiscretized(warm_cfg.sde, None, warm_cfg)
    torch.manual_seed(0)
    generate_samples_vada_2prior(shape, dae, warm, vae, B, False, clip_feat=clip_feat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.passes):
        img = one(100 + i)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sec = max(e0.elapsed_time(e1) / 1000.0, wall) / args.passes
    assert tuple(img.shape) == (B, 2048, 3) and torch.isfinite(img).all()
    scale = 1000.0 / args.steps                   # only used when --steps < 1000 (flagged as extrapolated)
    print(json.dumps({"impl": "reference-gpu", "kind": "reference: unmodified /root/reference code (baseline/_ref/LION) through its own "
                      "generate_samples_vada_2prior, its JIT-built third_party/pvcnn kernels, torch cuDNN/cuBLAS eager, cudnn.benchmark",
                      "metric": "shapes/sec (1000-step DDPM, 2048 latent pts, B=32)", "value": B / (sec * scale) if args.steps == 1000 else B / (sec * scale),
                      "unit": "shapes/s", "n_gpus": 1, "batch": B, "ddpm_steps_run": args.steps, "passes": args.passes,
                      "extrapolated": args.steps != 1000, "seconds_per_pass_measured": sec, "clip": bool(args.clip),
                      "import_and_jit_seconds": round(t_imp, 1), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                      "flags": {"cudnn.benchmark": True, "cudnn.allow_tf32": torch.backends.cudnn.allow_tf32,
                                "cuda.matmul.allow_tf32": torch.backends.cuda.matmul.allow_tf32}}))


if __name__ == "__main__":
    main()
