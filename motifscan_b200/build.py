"""Build libmsb200.so (the CUDA library) in-tree with nvcc for sm_100a.

    python -m motifscan_b200.build [--force]

The built library sits next to this file (git-ignored, but shipped to the GPU box with the
repo snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "msb200.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "common.cuh", "prefilter_tc.cuh")] + \
    [os.path.join(ROOT, "include", "msb200.h")]
OUT = os.path.join(HERE, "libmsb200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"),
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines`: experiment builds (bench_micro/tc_ablate.sh), e.g. defines=("MSB_TC_EXP=2",)."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
