"""Copy the UNMODIFIED reference tree into baseline/_ref/LION (git-ignored; it travels to the GPU box with the gpurun
snapshot) and pre-build its two JIT CUDA extensions in place, so that the reference's own code path can be timed on
the B200 next to lion_b200 (bench.py `gpu_baseline`, baseline/ref_gpu_arm.py).

    python baseline/make_ref_copy.py          # container only (needs /root/reference); no GPU needed, ~3 min

Nothing is edited: the copy exists because the reference JIT-builds next to its sources
(third_party/pvcnn/functional/backend.py:6-27, third_party/ChamferDistancePytorch/chamfer3D/dist_chamfer_3D.py:7-16)
and /root/reference is read-only.  `assets/` (images) is left out.  Called by __graft_entry__.build()."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "LION")


def copy_tree():
    if not os.path.isdir(os.path.join(SRC, "models")):
        return None
    if os.path.isdir(os.path.join(DST, "models")):
        return DST
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "assets", "__pycache__", "*.pyc"), symlinks=False)
    for dp, _, fs in os.walk(DST):                       # the source tree is read-only; the JIT builds need to write
        os.chmod(dp, 0o755)
        for f in fs:
            os.chmod(os.path.join(dp, f), 0o644)
    return DST


def prebuild(verbose=False):
    """Run the reference's own two load() recipes (same names, sources, flags, build directories) without importing
    its packages, so that their ninja builds are warm on the GPU box."""
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    src = os.path.join(DST, "third_party", "pvcnn", "functional")
    os.makedirs(os.path.join(src, "build"), exist_ok=True)
    pv = ["ball_query/ball_query.cpp", "ball_query/ball_query.cu", "grouping/grouping.cpp", "grouping/grouping.cu",
          "interpolate/neighbor_interpolate.cpp", "interpolate/neighbor_interpolate.cu", "interpolate/trilinear_devox.cpp",
          "interpolate/trilinear_devox.cu", "sampling/sampling.cpp", "sampling/sampling.cu", "voxelization/vox.cpp",
          "voxelization/vox.cu", "bindings.cpp"]
    # backend.py:8-27 passes no build_directory: torch's default is TORCH_EXTENSIONS_DIR/<name>; point it into the copy
    os.environ["TORCH_EXTENSIONS_DIR"] = os.path.join(HERE, "_ref", "torch_extensions")
    load(name="_pvcnn_backend", extra_cflags=["-O3", "-std=c++17"], verbose=verbose,
         sources=[os.path.join(src, "src", f) for f in pv], is_python_module=False)
    ch = os.path.join(DST, "third_party", "ChamferDistancePytorch", "chamfer3D")
    build_path = ch.replace("chamfer3D", "tmp")
    os.makedirs(build_path, exist_ok=True)
    load(name="chamfer_3D", sources=[os.path.join(ch, "chamfer_cuda.cpp"), os.path.join(ch, "chamfer3D.cu")],
         build_directory=build_path, verbose=verbose, is_python_module=False)


def main():
    d = copy_tree()
    if d is None:
        print("no /root/reference here: nothing to copy (the GPU box uses the shipped baseline/_ref)")
        return None
    prebuild(verbose="-v" in sys.argv)
    return d


if __name__ == "__main__":
    print(main())
