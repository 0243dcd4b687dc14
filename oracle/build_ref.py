"""TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference extension for sm_100a.

Compiles the reference's own three native source files
(/root/reference/models/PointUtils/src/{point_utils_api.cpp,furthest_point_sampling.cpp,
furthest_point_sampling_gpu.cu}) where they lie, into oracle/_ref/point_utils_cuda.so.
No reference source is copied into this repository; oracle/_ref/ is git-ignored but travels to the
GPU box with the gpurun snapshot.  The resulting module can only *execute* on a GPU (CUDA-only
kernels); tests/test_gpu_ref_ext.py uses it as the bit-exact GPU oracle for FPS / weighted FPS /
gather, and bench.py can time it as "the reference kernel recompiled for the same box".

Usage:  python oracle/build_ref.py        (takes ~3 min: torch headers)
"""
import os
import sys

REF_SRC = "/root/reference/models/PointUtils/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> str:
    so = os.path.join(OUT, "point_utils_cuda.so")
    if os.path.exists(so):
        return so
    if not os.path.isdir(REF_SRC):
        raise FileNotFoundError(f"{REF_SRC} not present (only exists in the build container)")
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    load(
        name="point_utils_cuda",
        sources=[os.path.join(REF_SRC, f) for f in (
            "point_utils_api.cpp", "furthest_point_sampling.cpp", "furthest_point_sampling_gpu.cu")],
        extra_cflags=["-g"],               # reference setup.py:14
        extra_cuda_cflags=["-O2"],         # reference setup.py:15
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,            # do not import (no GPU needed to build)
    )
    return so


STAGE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
# what the reference arm of bench.py needs of the reference, relative to /root/reference:
STAGED = ("models/utils.py", "models/HRegNet/layers.py", "models/HRegNet/models.py", "models/model_v2/layers.py",
          "models/model_v2/models.py", "models/model_v4/layers.py", "models/model_v4/models.py",
          "ckpt/pretrained/nusc_feats.pth")


def stage_reference(ref_root: str = "/root/reference") -> str:
    """Stages the UNMODIFIED reference files of the path under baseline/_ref/ (git-ignored: never part of this
    repository's history, but it travels to the GPU box with the gpurun snapshot, where /root/reference does not exist).
    `bench.py --impl reference` and the `gpu_reference` leg import the reference's own model graph from there, byte for
    byte what the reference ships; nothing under pcd_reg_hregnet_b200/ reads it."""
    import shutil
    if not os.path.isdir(ref_root):
        raise FileNotFoundError(ref_root)
    for rel in STAGED:
        dst = os.path.join(STAGE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, rel), dst)
    return STAGE


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(stage_reference())
