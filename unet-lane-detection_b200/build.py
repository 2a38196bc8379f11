"""Build libunet_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python unet-lane-detection_b200/build.py [--force]
nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunet_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["ptx.cuh", "conv_umma.cuh", "conv_halo.cuh", "stem_umma.cuh", "stem_halo.cuh", "epilogue.cuh", "aux_kernels.cuh", "train_kernels.cuh",
           "wgrad_umma.cuh", "wgrad_halo.cuh", "stem_wgrad_umma.cuh", "train_capi.cuh", os.path.join("..", "..", "include", "unet_b200.h")]
NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
