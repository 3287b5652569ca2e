"""Import the *unmodified* reference (read-only at /root/reference) on CPU.

Only usable in the build container (the GPU box has no /root/reference); used by
make_golden.py and by the container-only cross-checks in tests/test_oracle_vs_reference.py.

Stubs (all arithmetic-free for the sampling path):
  comet_ml, matplotlib, clip, calmsize, diffusers  -- not installed
  third_party.PyTorchEMD                          -- present but does not build on torch>=1.11
  third_party.pvcnn.functional                    -- the CUDA extension; replaced by the CPU
                                                     restatement in oracle/point_ops.py
"""
import os
import sys
import types

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def available():
    return os.path.isdir(os.path.join(REF, "models"))


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


def _stub(name):
    m = _Anything(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def install():
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    if REF not in sys.path:
        sys.path.insert(1, REF)
    os.environ.setdefault("quiet", "1")
    for n in ["comet_ml", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "clip", "calmsize", "diffusers",
              "mpl_toolkits", "mpl_toolkits.mplot3d", "open3d",
              "third_party.PyTorchEMD", "third_party.PyTorchEMD.emd", "third_party.PyTorchEMD.emd_nograd"]:
        if n not in sys.modules:
            _stub(n)
    import torch
    from oracle import point_ops as P

    F = types.ModuleType("third_party.pvcnn.functional")
    F.avg_voxelize = lambda feats, coords, r: P.avg_voxelize(feats, coords[:, :3], r)[0]
    F.trilinear_devoxelize = lambda feats, coords, r, training=True: P.trilinear_devoxelize(feats, coords[:, :3], r)
    F.furthest_point_sample = lambda coords, m, normals=None: P.furthest_point_sample(coords, m)
    F.gather = P.gather
    F.ball_query = lambda c, p, radius, k: P.ball_query(c[:, :3], p[:, :3], radius, k)
    F.grouping = P.grouping
    F.nearest_neighbor_interpolate = lambda p, c, cf: P.nearest_neighbor_interpolate(p[:, :3], c[:, :3], cf)
    import third_party  # namespace package under /root/reference
    pv = types.ModuleType("third_party.pvcnn")
    pv.__path__ = []
    pv.functional = F
    sys.modules["third_party.pvcnn"] = pv
    sys.modules["third_party.pvcnn.functional"] = F
    third_party.pvcnn = pv

    # the reference hard-codes device='cuda' in its sampling loop (diffusion_pvd.py:136-140,
    # 237,257,286); on this CPU-only container map those to CPU.
    def _cpu_kw(fn):
        def w(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return w
    for name in ["randn", "ones", "zeros", "tensor", "rand"]:
        setattr(torch, name, _cpu_kw(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def load_cfg(name="airplane_prior_cfg.yml", overrides=()):
    install()
    from default_config import cfg as base
    cfg = base.clone()
    cfg.merge_from_file(os.path.join(REF, "config", name))
    if overrides:
        cfg.merge_from_list(list(overrides))
    return cfg
